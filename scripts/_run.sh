cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r01_bench.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frame'], d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r01_bench_reference.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python scripts/benchmark_csv.py --iterations 2 --threads 1,4,16 --out gpurun_out/benchmark_results 2>&1 | tail -12
