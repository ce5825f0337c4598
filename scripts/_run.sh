cd $GRAFT_REPO_ROOT
N=$1
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r01_bench_n${N}_$tag.json 2> gpurun_out/r01_bench_n${N}_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r01_bench_n${N}_$tag.json").read().strip().splitlines()[-1]); print("N=$N $tag", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r01_bench_n${N}_$tag.err").read()[-1500:])
PY
}
run peer --steps 200 --warmup 10 --gather peer
run synth10k_peer --steps 20 --warmup 3 --workload synth10k
run synth100k_peer --steps 8 --warmup 3 --workload synth100k
