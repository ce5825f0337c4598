cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --workload synth10k > gpurun_out/r01_bench_synth10k_bvh.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_synth10k_bvh.json')); print(d['ms_per_step'], d['value'])"
python bench.py --steps 8 --warmup 3 --workload synth100k > gpurun_out/r01_bench_synth100k.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_synth100k.json')); print(d['ms_per_step'], d['value'])"
python bench.py --no-cpu-baseline --steps 200 > gpurun_out/bench_c.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
