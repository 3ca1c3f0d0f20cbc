#!/bin/bash
# A/B experiments: build a second copy of the product library with extra nvcc flags (e.g. -DPRL_TC_TIMING_WAITS) next to the regular one.
#   tools/ab_build.sh <name> [nvcc flags ...]   ->  parallel-reinforcement-learning_b200/prl_b200/libprl_b200_<name>.so  (select with PRL_B200_LIB=<path>)
set -e
name=$1; shift
cd "$(dirname "$0")/../parallel-reinforcement-learning_b200/csrc"
mkdir -p _obj/ab_$name
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr --expt-extended-lambda "$@" -c $f -o _obj/ab_$name/${f%.cu}.o &
done
wait
nvcc -shared -o ../prl_b200/libprl_b200_$name.so _obj/ab_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built libprl_b200_$name.so
