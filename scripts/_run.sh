cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q -k "bvh or large" 2>&1 | tail -5
RT_ACCEL=2 timeout 300 python scripts/probe_scene.py complex 1920 1080 5 5
RT_ACCEL=2 timeout 300 python scripts/probe_scene.py synth:10000:420 3840 2160 5 3
timeout 600 python scripts/probe_scene.py synth:100000:421 1920 1080 8 3
timeout 600 python scripts/probe_scene.py synth:100000:421 7680 4320 8 2
