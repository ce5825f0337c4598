cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q -k "bvh or large or full_size" 2>&1 | tail -2
RT_ACCEL=2 timeout 300 python scripts/probe_scene.py synth:10000:420 3840 2160 5 4 | cut -c1-150;  timeout 300 python scripts/probe_scene.py synth:100000:421 7680 4320 8 4 | cut -c1-150
