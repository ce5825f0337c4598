cd $GRAFT_REPO_ROOT
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize.py 2>&1 | tail -40
