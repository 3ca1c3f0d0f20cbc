#!/bin/bash
# Round-2 evidence on one B200: the bench line, then (separately - numbers printed under ncu are never bench values) the ncu launch
# list of the same command and ncu --set full captures of the three kernels that carry the step.
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 5 --no-extra-configs --no-cpu-baseline > gpurun_out/r02_prof_bench.json 2> gpurun_out/r02_prof_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extra-configs --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_ppo_grad_tc -s 300 -c 1 -o gpurun_out/r02_tc_full \
    python bench.py --steps 1 --warmup 1 --no-extra-configs --no-cpu-baseline > gpurun_out/r02_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 1 -c 1 -o gpurun_out/r02_rollout_full \
    python bench.py --steps 1 --warmup 1 --no-extra-configs --no-cpu-baseline > gpurun_out/r02_ncu_rollout.log 2>&1; echo "ncu rollout rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gae_columns_ring -s 1 -c 1 -o gpurun_out/r02_gaecol_full \
    python bench.py --steps 1 --warmup 1 --no-extra-configs --no-cpu-baseline > gpurun_out/r02_ncu_gae.log 2>&1; echo "ncu gae rc=$?"
ls -la gpurun_out/*.ncu-rep
