"""In-tree build of librt_b200.so / ray_cuda (explicit nvcc, sm_100a)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _run(cmd, cwd=None):
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    p = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), p.stdout[-4000:]))
    return p.stdout


def build_library():
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see Makefile)."""
    return _run(["make", "-C", HERE, "all"])


def build_all():
    return build_library()
