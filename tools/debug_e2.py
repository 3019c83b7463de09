"""Fused final layer (STPDE_FUSE_FINAL=1) against the separate final_blend kernel (=0) over several shapes; each
setting runs in its own process (the switch is read once per process)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

CASES = [(8, 16, (3, 4, 5), 1, 1000), (16, 16, (4, 6, 5), 2, 1500), (32, 32, (4, 16, 16), 1, 5000), (32, 128, (32, 32, 32), 1, 65536),
         (64, 32, (4, 16, 16), 1, 20000), (128, 32, (4, 16, 16), 1, 70000)]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import space_time_pde_b200 as sp
    from space_time_pde_b200 import _lib
    from space_time_pde_b200.equations import JetSpec
    dev = torch.device("cuda:0")
    lib = _lib.load()
    out = {}
    for (nf, c, gs, b, p) in CASES:
        torch.manual_seed(nf)
        model = sp.ImNet(dim=3, in_features=c, out_features=4, nf=nf, activation=sp.NONLINEARITIES["softplus"]).to(dev)
        grid = (torch.randn(b, *gs, c) * 0.5).to(dev)
        q = torch.rand(b, p, 3, device=dev)
        spec = JetSpec((0, 1, 2), ((1, 1), (2, 2)))
        lib.stpde_profile_enable(1); _lib.profile_read()
        with torch.no_grad():
            y, jt = sp.fused_query(grid, q, 0., 1., list(model.fc), "softplus", None, spec=spec)
            y2, jt2 = sp.fused_query(grid, q[:, ::3], 0., 1., list(model.fc), "softplus", None, spec=spec)
        prof = _lib.profile_read(); lib.stpde_profile_enable(0)
        out[str((nf, c, b, p))] = {"y": y.cpu(), "jt": jt.cpu(), "fb_launches": prof.get("final_blend", (0, 0))[1],
                                  "subset_equal": bool(torch.equal(y2, y[:, ::3]) and torch.equal(jt2, jt[:, :, ::3]))}
    torch.save(out, sys.argv[2])
    sys.exit(0)

outs = {}
for fuse in ("1", "0"):
    path = f"/tmp/dbg_{fuse}.pt"
    subprocess.run([sys.executable, __file__, "child", path], env=dict(os.environ, STPDE_FUSE_FINAL=fuse), check=True)
    outs[fuse] = torch.load(path)
for k in outs["1"]:
    a, b = outs["1"][k], outs["0"][k]
    dy = float((a["y"] - b["y"]).abs().max() / b["y"].abs().max())
    dj = float((a["jt"] - b["jt"]).abs().max() / b["jt"].abs().max())
    print(k, "final_blend launches fused/unfused:", a["fb_launches"], b["fb_launches"], "rel diff y %.2e jets %.2e" % (dy, dj),
          "subset bitwise equal fused/unfused:", a["subset_equal"], b["subset_equal"])
