cd $GRAFT_REPO_ROOT
timeout 300 python scripts/probe_scene.py complex 1920 1080 5 30
for v in w128x3 w128x4 w256x1; do echo $v; RTB200_LIB=$PWD/build_tools/librt_$v.so timeout 300 python scripts/probe_scene.py complex 1920 1080 5 30; done
