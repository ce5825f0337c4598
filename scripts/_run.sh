cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/probe_scene.py complex 1920 1080 5 30
timeout 300 python scripts/probe_wave.py complex 1920 1080 5
timeout 300 python scripts/probe_scene.py medium 1920 1080 5 30
RT_ACCEL=2 timeout 300 python scripts/probe_scene.py synth:10000:420 3840 2160 5 3
