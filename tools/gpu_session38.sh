#!/bin/bash
# are the vertex-adjoint atomics what bounds the nf=32 dgrad kernels?  (a) timing with the atomics dropped (debug switch,
# results wrong), (b) microbenchmark of scalar / v2 / v4 float reductions into a [nvert][992] table
O=gpurun_out/s38; mkdir -p $O
for ng in 0 1; do
  if [ $ng = 1 ]; then export STPDE_DEBUG_NO_GVB=1; else unset STPDE_DEBUG_NO_GVB; fi
  echo "== NO_GVB=$ng"
  timeout 300 python tools/train_chunk_probe.py 8192 40960 2>&1 | tail -2
  STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py fp16x3 65536 3 2>&1 | grep -E "^\{"
done 2>&1 | tee $O/no_gvb.log
unset STPDE_DEBUG_NO_GVB
./tools/redbw.bin 2>&1 | tee $O/redbw.log
