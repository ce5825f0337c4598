cd $GRAFT_REPO_ROOT
timeout 300 python scripts/probe_rank.py
