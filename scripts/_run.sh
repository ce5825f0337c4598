cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err; cat gpurun_out/r01_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>/dev/null; cat gpurun_out/r01_bench_reference.json
python bench.py --steps 20 --warmup 3 --workload synth10k --accel 1 > gpurun_out/r01_bench_synth10k_tables.json 2>/dev/null; cat gpurun_out/r01_bench_synth10k_tables.json
python bench.py --steps 20 --warmup 3 --workload synth10k > gpurun_out/r01_bench_synth10k_bvh.json 2>/dev/null; cat gpurun_out/r01_bench_synth10k_bvh.json
python bench.py --steps 5 --warmup 3 --workload synth100k > gpurun_out/r01_bench_synth100k.json 2>/dev/null; cat gpurun_out/r01_bench_synth100k.json
python bench.py --steps 50 --warmup 3 --workload medium > gpurun_out/r01_bench_medium.json 2>/dev/null
python bench.py --steps 50 --warmup 3 --workload simple > gpurun_out/r01_bench_simple.json 2>/dev/null
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:"^k_" --launch-skip 28 --launch-count 7 -o gpurun_out/r01_prof_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
