#!/bin/bash
# A/B of the flat GAE kernels on the GPU box: parity tests, then tools/prof_hbm.py with and without the prefetching kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gae or full_size" > gpurun_out/gae_t.log 2>&1; echo "rc=$?" >> gpurun_out/gae_t.log
tail -n 4 gpurun_out/gae_t.log
for pf in 0 1; do echo "PRL_GAE_PF=$pf"; PRL_GAE_PF=$pf timeout 120 python tools/prof_hbm.py 20 2>&1 | grep -i gae; done | tee gpurun_out/gae_ab.log
