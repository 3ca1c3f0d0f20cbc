python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_1gpu.json 2> gpurun_out/r02_final_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_ref.json 2>/dev/null; echo "ref rc=$?"
