cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for e in 0 1; do RT_NO_PDL=$e python bench.py --no-cpu-baseline --steps 300 > gpurun_out/bench_pdl$e.json 2>gpurun_out/bench_pdl$e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$e.json')); print('RT_NO_PDL=$e', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"; done
RT_NO_PDL=0 timeout 300 python scripts/probe_rank.py | grep "n=8"
RT_NO_PDL=1 timeout 300 python scripts/probe_rank.py | grep "n=8"
