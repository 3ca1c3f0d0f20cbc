#!/bin/bash
# Round-end evidence on one B200: bench line, ncu launch list of the same command, ncu --set full of the flat GAE kernel
# and of the seeded-reset kernel.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "short bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gae_flat_pf -c 1 -o gpurun_out/gae_pf_final python tools/prof_hbm.py 1 > gpurun_out/ncu_gae.log 2>&1; echo "ncu gae rc=$?"
