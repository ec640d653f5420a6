#!/bin/bash
# Same-box A/B builds of the field kernels with other compile-time switches (RSN_STASH_LAG_FWD / _BWD, RSN_FWD_TRAIN_SPLIT,
# RSN_BWD_SPLIT, see csrc/field_fwd.cu and csrc/field_bwd_body.cuh):
#   scripts/mkvariant.sh <name> "<-D flags>"   ->  scratch/lib_<name>.so   (the product build's other objects are re-used)
#   RSN_B200_LIB=scratch/lib_<name>.so python scripts/bench_field_train.py 16384 128 30
# (scratch/ is git-ignored and travels to the GPU box with gpurun; run python -m reflect_sampling_nerf_b200.build first.)
set -e
cd "$(dirname "$0")/../reflect_sampling_nerf_b200"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p ../scratch /tmp/var_$1
nvcc $F $2 -c csrc/field_fwd.cu -o /tmp/var_$1/field_fwd.o &
nvcc $F $2 -c csrc/field_bwd.cu -o /tmp/var_$1/field_bwd.o &
wait
OBJS=$(ls build/*.o | grep -v "field_fwd.o\|field_bwd.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../scratch/lib_$1.so $OBJS /tmp/var_$1/field_fwd.o /tmp/var_$1/field_bwd.o
echo built scratch/lib_$1.so
