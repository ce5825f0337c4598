cd $GRAFT_REPO_ROOT
timeout 1500 python scripts/fuzz_parity.py 1000 6000 2>&1 | tail -12
