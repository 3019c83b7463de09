#!/bin/bash
O=gpurun_out/s20; mkdir -p $O
export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_t.so
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x -k "rb2_spec_smooth or stash_is_reused or fused_vs_torch or kinked or tiny_cotangents or multi_chunk or fused_loss or swish_beta or encoder_gradients or chunked_training or golden_gradients and rb2" 2>&1 | tail -5 | tee $O/pytest.log
for p in 1280 8192; do python tools/train_chunk_probe.py $p 40960 2>&1 | tail -2; done | tee $O/chunk_probe.log
timeout 300 python tools/profile_small_step.py 2>&1 | head -3 | tee $O/small_step.log
