cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 300 > gpurun_out/bench_e.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_e.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 300 python scripts/probe_scene.py complex 1920 1080 5 30 | cut -c1-150
