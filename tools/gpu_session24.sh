#!/bin/bash
# A/B on one box: main library (z planes loaded straight into registers) vs variant t (z chunks staged by cp.async)
O=gpurun_out/s24; mkdir -p $O
T=$PWD/space_time_pde_b200/libstpde_t.so
STPDE_LIB_PATH=$T timeout 900 python -m pytest tests/test_gpu_backward.py -q -x -k "rb2_spec_smooth or stash_is_reused or fused_vs_torch or kinked or tiny_cotangents or multi_chunk or fused_loss or swish_beta or encoder_gradients or chunked_training or cuda_graph or golden_gradients and rb2" 2>&1 | tail -5 | tee $O/pytest.log
for lib in main t; do
  if [ $lib = t ]; then export STPDE_LIB_PATH=$T; else unset STPDE_LIB_PATH; fi
  echo "== lib $lib"
  for prec in fp16x3 fp16; do
    STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py $prec 65536 3 2>&1 | tail -2
    NF=32 STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py $prec 262144 3 2>&1 | tail -2
  done
  timeout 300 python tools/profile_small_step.py 2>&1 | head -2 | cut -c1-900
done 2>&1 | tee $O/ab.log
