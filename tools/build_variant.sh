#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. -DFLAG2=.." : experimental build of the epilogue-dependent translation units into
# space_time_pde_b200/libstpde_NAME.so (all other objects are reused from the main build).  STPDE_LIB_PATH selects it.
set -e
NAME=$1; FLAGS=$2
cd /root/repo/space_time_pde_b200
mkdir -p build/var_$NAME
for f in tc_layers_a_hi tc_layers_b_hi tc_bwd_a_pair_hi tc_bwd_a_single_hi tc_bwd_b_pair_hi tc_bwd_b_single_hi tc_bwd_c_pair_hi tc_bwd_c_single_hi tc_path tc_bwd api; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I ../include -DSTPDE_ONLY_RB2 $FLAGS -c csrc/$f.cu -o build/var_$NAME/$f.o 2> build/var_$NAME/$f.log &
done
wait
OBJS=""
for f in simt_kernels bwd_kernels tc_layers_a_lo tc_layers_b_lo tc_bwd_a_pair_lo tc_bwd_a_single_lo tc_bwd_b_pair_lo tc_bwd_b_single_lo tc_bwd_c_pair_lo tc_bwd_c_single_lo profile; do OBJS="$OBJS build/$f.o"; done
for f in tc_layers_a_hi tc_layers_b_hi tc_bwd_a_pair_hi tc_bwd_a_single_hi tc_bwd_b_pair_hi tc_bwd_b_single_hi tc_bwd_c_pair_hi tc_bwd_c_single_hi tc_path tc_bwd api; do OBJS="$OBJS build/var_$NAME/$f.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libstpde_$NAME.so $OBJS -lcuda
ls -la libstpde_$NAME.so
