cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q -k "large or full_size or synthetic" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --workload synth10k --no-cpu-baseline > gpurun_out/x10k.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/x10k.json')); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"
python bench.py --steps 8 --warmup 3 --workload synth100k --no-cpu-baseline > gpurun_out/x100k.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/x100k.json')); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"
