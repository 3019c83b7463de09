"""ctypes binding of the C-ABI library (include/stpde.h) and the in-tree nvcc build.

The shared object lives next to this file (``libstpde.so``) so that it travels with the
repository snapshot; it is never pip-installed.  Loading fails loudly: there is no CPU
fallback for the hot path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("STPDE_LIB_PATH") or os.path.join(HERE, "libstpde.so")   # (override: kernel experiments)
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

# the tensor-core kernel families are instantiated for K = 1..10 jet components: one translation unit per family and
# half of the K range (heaviest first, so that the parallel build ends together)
SOURCES = [f"{fam}_{half}.cu" for fam in ("tc_bwd_a_pair", "tc_bwd_a_single", "tc_layers_a", "tc_layers_b", "tc_bwd_b_pair",
                                           "tc_bwd_b_single", "tc_bwd_c_pair", "tc_bwd_c_single") for half in ("hi", "lo")] + \
          ["api.cu", "simt_kernels.cu", "bwd_kernels.cu", "tc_path.cu", "tc_bwd.cu", "profile.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

MAX_DIM, MAX_LAYERS, MAX_FIRST, MAX_SECOND, MAX_COMPONENTS, MAX_OUT = 4, 8, 4, 10, 10, 8

ERRORS = {0: "OK", -1: "EINVAL", -2: "ENOMEM", -3: "ECUDA", -4: "EINDEX", -5: "EUNSUPPORTED", -6: "ERANGE"}

ACT_CODES = {"tanh": 0, "relu": 1, "softplus": 2, "elu": 3, "swish": 4, "leakyrelu": 5}
PRECISIONS = {"fp32": 0, "fp16x3": 1, "fp16": 2}


class StpdeDesc(ctypes.Structure):
    """Mirror of ``stpde_desc_t`` (include/stpde.h)."""
    _fields_ = [
        ("batch", ctypes.c_int32), ("npts", ctypes.c_int32), ("dim", ctypes.c_int32),
        ("grid_size", ctypes.c_int32 * MAX_DIM), ("channels", ctypes.c_int32),
        ("n_layers", ctypes.c_int32), ("widths", ctypes.c_int32 * MAX_LAYERS),
        ("act_kind", ctypes.c_int32), ("act_param", ctypes.c_float),
        ("n_first", ctypes.c_int32), ("first_dirs", ctypes.c_int32 * MAX_FIRST),
        ("n_second", ctypes.c_int32), ("second_pairs", (ctypes.c_int32 * 2) * MAX_SECOND),
        ("precision", ctypes.c_int32),
        ("xmin", ctypes.c_float * MAX_DIM), ("xmax", ctypes.c_float * MAX_DIM),
        ("reserved", ctypes.c_int32 * 8),
    ]


EXPORTS = ["stpde_version", "stpde_last_error", "stpde_desc_size", "stpde_device_sm_count", "stpde_workspace_bytes",
           "stpde_interp_coefficients", "stpde_interp", "stpde_jet_forward", "stpde_jet_forward_host",
           "stpde_backward_workspace_bytes", "stpde_backward_chunk_points", "stpde_jet_backward",
           "stpde_jet_forward_train",
           "stpde_residuals", "stpde_residuals_backward", "stpde_residual_loss_blocks", "stpde_residual_loss",
           "stpde_residual_loss_backward", "stpde_profile_enable", "stpde_profile_read", "stpde_profile_slot_name"]


class StpdeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"stpde {ERRORS.get(code, code)}: {message}")
        self.code = code


def sources_newer_than_lib() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "stpde.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libstpde.so (nvcc cross-compiles without a GPU).

    Translation units are compiled in parallel (one nvcc per .cu into build/*.o) and linked once."""
    if not force and not sources_newer_than_lib():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(INCLUDE, "stpde.h")]
    t_headers = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        # incremental: an object newer than its source and every header is kept
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(t_headers, os.path.getmtime(os.path.join(CSRC, src))):
            return obj
        cmd = ["nvcc"] + flags + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, cwd=CSRC)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), (os.cpu_count() or 8) + 2)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    link = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-lcuda"]
    if verbose:
        print(" ".join(link), flush=True)
    subprocess.run(link, check=True, cwd=CSRC)
    return LIB_PATH


_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()


def _declare(lib: ctypes.CDLL) -> None:
    c_void_p, c_int, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    i32, i32p, i64p, f32p = ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float)
    descp = ctypes.POINTER(StpdeDesc)
    lib.stpde_version.restype = c_int
    lib.stpde_last_error.restype = ctypes.c_char_p
    lib.stpde_device_sm_count.restype = c_int
    lib.stpde_desc_size.restype = c_size_t
    lib.stpde_workspace_bytes.restype = c_size_t
    lib.stpde_workspace_bytes.argtypes = [descp]
    lib.stpde_interp_coefficients.restype = c_int
    lib.stpde_interp_coefficients.argtypes = [i32, i32, i32, i32p, i32, c_void_p, i64p, c_void_p, i64p, f32p, f32p,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.stpde_interp.restype = c_int
    lib.stpde_interp.argtypes = [i32, i32, i32, i32p, i32, c_void_p, i64p, c_void_p, i64p, f32p, f32p, c_void_p,
                                 c_void_p, c_void_p]
    lib.stpde_jet_forward.restype = c_int
    lib.stpde_jet_forward.argtypes = [descp, c_void_p, i64p, c_void_p, i64p, ctypes.POINTER(c_void_p),
                                      ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
                                      c_void_p]
    lib.stpde_backward_workspace_bytes.restype = c_size_t
    lib.stpde_backward_workspace_bytes.argtypes = [descp]
    lib.stpde_jet_backward.restype = c_int
    lib.stpde_jet_backward.argtypes = [descp, c_void_p, i64p, c_void_p, i64p, ctypes.POINTER(c_void_p),
                                       ctypes.POINTER(c_void_p), c_void_p, c_void_p, ctypes.POINTER(c_void_p),
                                       ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_size_t, i32, c_void_p,
                                       c_void_p]
    lib.stpde_backward_chunk_points.restype = ctypes.c_int64
    lib.stpde_backward_chunk_points.argtypes = [descp, c_size_t]
    lib.stpde_jet_forward_train.restype = c_int
    lib.stpde_jet_forward_train.argtypes = lib.stpde_jet_forward.argtypes
    lib.stpde_jet_forward_host.restype = c_int
    lib.stpde_jet_forward_host.argtypes = [descp, c_void_p, c_void_p, ctypes.POINTER(c_void_p),
                                           ctypes.POINTER(c_void_p), c_void_p, c_void_p]
    lib.stpde_profile_enable.restype = c_int
    lib.stpde_profile_enable.argtypes = [c_int]
    lib.stpde_profile_read.restype = c_int
    lib.stpde_profile_read.argtypes = [ctypes.POINTER(ctypes.c_double), i64p, c_int]
    lib.stpde_profile_slot_name.restype = ctypes.c_char_p
    lib.stpde_profile_slot_name.argtypes = [c_int]
    lib.stpde_residuals_backward.restype = c_int
    lib.stpde_residuals_backward.argtypes = [i32, i32, i32, i32, i32, c_void_p, i64p, c_void_p, c_void_p, i32p, i32, f32p,
                                             i32, i32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.stpde_residual_loss_blocks.restype = i32
    lib.stpde_residual_loss_blocks.argtypes = [ctypes.c_int64]
    lib.stpde_residual_loss.restype = c_int
    lib.stpde_residual_loss.argtypes = [i32, i32, i32, i32, i32, c_void_p, i64p, c_void_p, c_void_p, c_void_p, i32p, i32, f32p,
                                        i32, i32, i32, c_void_p, c_void_p]
    lib.stpde_residual_loss_backward.restype = c_int
    lib.stpde_residual_loss_backward.argtypes = [i32, i32, i32, i32, i32, c_void_p, i64p, c_void_p, c_void_p, c_void_p, i32p, i32,
                                                 f32p, i32, i32p, i32, f32p, i32, i32, i32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.stpde_residuals.restype = c_int
    lib.stpde_residuals.argtypes = [i32, i32, i32, i32, i32, c_void_p, i64p, c_void_p, c_void_p, i32p, i32, f32p,
                                    i32, i32, c_void_p, c_void_p]


def load() -> ctypes.CDLL:
    """Load libstpde.so; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a) first; "
                                   "the hot path has no CPU fallback")
            lib = ctypes.CDLL(LIB_PATH)
            missing = [s for s in EXPORTS if not hasattr(lib, s)]
            if missing:
                raise RuntimeError(f"{LIB_PATH} does not export {missing}")
            _declare(lib)
            if lib.stpde_desc_size() != ctypes.sizeof(StpdeDesc):
                raise RuntimeError("stpde_desc_t layout mismatch between libstpde.so and the ctypes binding")
            _lib = lib
    return _lib


def profile_read():
    """{slot name: (milliseconds, launches)} since the previous read (synchronises the device)."""
    lib = load()
    n = 32
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    lib.stpde_profile_read(ms, cnt, n)
    return {lib.stpde_profile_slot_name(i).decode(): (ms[i], cnt[i]) for i in range(n)
            if lib.stpde_profile_slot_name(i)}


def check(rc: int) -> None:
    if rc != 0:
        raise StpdeError(rc, load().stpde_last_error().decode())
