cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err; cut -c1-400 gpurun_out/r01_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>/dev/null
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:"^k_" --launch-skip 28 --launch-count 7 -o gpurun_out/r01_prof_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
RT_ACCEL=2 ncu --set full --import-source on --clock-control none --kernel-name regex:"^k_" --launch-skip 15 --launch-count 15 -o gpurun_out/r01_prof_bvh10k -f python scripts/probe_scene.py synth:10000:420 1920 1080 5 2 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
