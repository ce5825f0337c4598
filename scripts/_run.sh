cd $GRAFT_REPO_ROOT
for v in "" sh3 sh5 sh6; do echo "== $v"; RTB200_LIB=${v:+$PWD/build_tools/librt_$v.so} timeout 300 python scripts/probe_wave2.py complex 1920 1080 5 1 2>/dev/null | sed -n 2p; done
