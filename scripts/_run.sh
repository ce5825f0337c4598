cd $GRAFT_REPO_ROOT
RT_ACCEL=2 timeout 300 python scripts/probe_scene.py synth:10000:420 3840 2160 5 4 | cut -c1-400;  timeout 300 python scripts/probe_scene.py synth:100000:421 7680 4320 8 4 | cut -c1-400
