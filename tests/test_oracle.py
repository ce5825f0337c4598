"""Pins the oracle (oracle/rt_oracle.c) to the reference: md5-identical PPMs and identical
hit-index / shadow maps versus outputs of the UNMODIFIED reference sources recorded in
tests/golden/golden.json (tests/golden/make_golden.py), plus the reference's only known-answer
vectors (populi-files/demo1_intersection.cpp:36-40).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, scene_path

# md5s the survey recorded from the unmodified binaries
BASELINE_MD5 = {
    ("simple", 1280, 720, 10): "060e7c2d396190303950487c401d197c",
    ("medium", 1280, 720, 10): "0c77297b412d5a9b06b812b021e7aded",
    ("complex", 1280, 720, 10): "69dd7731ca643e0a54df13941c444b1b",
    ("medium", 1920, 1080, 5): "d7b3e12ab0fa370068f83360aa976cba",
    ("complex", 1920, 1080, 5): "736538713f2a8bb95d5711a8e71c9c60",
}


def test_golden_file_agrees_with_baseline_md(golden):
    seen = {(g["scene"], g["W"], g["H"], g["depth"]): g["md5"] for g in golden["images"]}
    for key, h in BASELINE_MD5.items():
        assert seen[key] == h


@pytest.mark.parametrize("name,W,H,D", [
    ("simple", 1280, 720, 10), ("medium", 1280, 720, 10), ("complex", 1280, 720, 10),
    ("simple", 160, 90, 5), ("medium", 160, 90, 5), ("complex", 160, 90, 5),
    ("simple", 97, 61, 3), ("medium", 97, 61, 3), ("complex", 97, 61, 3),
    ("medium", 1920, 1080, 5), ("complex", 1920, 1080, 5),
])
def test_oracle_ppm_md5_matches_reference(rt, oracle, scenes, golden, name, W, H, D):
    want = [g["md5"] for g in golden["images"] if (g["scene"], g["W"], g["H"], g["depth"]) == (name, W, H, D)]
    assert want and len(set(want)) == 1
    r = oracle.render(scenes[name], W, H, D)
    got = hashlib.md5(rt.ppm_text(r["rgb"]).encode()).hexdigest()
    assert got == want[0]


@pytest.mark.parametrize("name", ["simple", "medium", "complex"])
def test_oracle_hit_index_and_shadow_maps_match_reference(oracle, scenes, name):
    z = np.load(os.path.join(GOLDEN, "small_%s.npz" % name))
    r = oracle.render(scenes[name], 160, 90, 5, want_idx=True)
    assert np.array_equal(r["hit_idx"], z["hit_idx"])
    assert np.array_equal(r["shadow_mask"], z["shadow_mask"])
    assert np.array_equal(r["rgb"][::-1], z["rgb_top_first"])
    cn = json.loads(str(z["counters"]))
    for k in ("closest_queries", "hits", "shadow_queries", "occluded"):
        assert r["counters"][k] == cn[k]


@pytest.mark.parametrize("name,W,H,D", [("simple", 1280, 720, 10), ("complex", 1920, 1080, 5)])
def test_oracle_full_size_maps_and_counters(oracle, scenes, golden, name, W, H, D):
    g = [c for c in golden["counters"] if (c["scene"], c["W"], c["H"], c["depth"]) == (name, W, H, D)][0]
    r = oracle.render(scenes[name], W, H, D, want_idx=True)
    for k in ("closest_queries", "hits", "shadow_queries", "occluded"):
        assert r["counters"][k] == g[k]
    assert r["counters"]["alive"][:D] == g["alive"]
    assert hashlib.sha256(r["hit_idx"].tobytes()).hexdigest() == g["hit_idx_sha256"]
    assert hashlib.sha256(r["shadow_mask"].tobytes()).hexdigest() == g["shadow_mask_sha256"]


def test_survey_ray_counts(golden):
    """R_c / hits / R_s of the BASELINE configs."""
    want = {("simple", 1280, 720, 10): (943122, 515264, 1030528),
            ("medium", 1920, 1080, 5): (2574504, 1654420, 4963260),
            ("complex", 1920, 1080, 5): (2453481, 1373442, 6867210)}
    for c in golden["counters"]:
        key = (c["scene"], c["W"], c["H"], c["depth"])
        if key in want:
            assert (c["closest_queries"], c["hits"], c["shadow_queries"]) == want[key]


def test_known_answer_intersections(oracle):
    """populi-files/demo1_intersection.cpp:36-40: unit sphere at the origin."""
    c, r = (0, 0, 0), 1.0
    assert oracle.intersect((5, 0, 0), (-1, 0, 0), c, r) == (True, 4.0)
    assert oracle.intersect((5, 5, 0), (-1, 0, 0), c, r)[0] is False
    assert oracle.intersect((5, 1, 0), (-1, 0, 0), c, r) == (True, 5.0)      # tangent: disc == 0 branch
    assert oracle.intersect((0, 0, 5), (0, 0, -1), c, r) == (True, 4.0)


def test_intersect_quirks(oracle):
    """include/sphere.h:34-58: no t-epsilon; inside -> far root; behind -> miss;
    tangent behind the origin is still a hit with negative t."""
    c, r = (0, 0, 0), 1.0
    assert oracle.intersect((0, 0, 0), (1, 0, 0), c, r) == (True, 1.0)        # origin inside
    assert oracle.intersect((5, 0, 0), (1, 0, 0), c, r)[0] is False           # both roots negative
    hit, t = oracle.intersect((5, 1, 0), (1, 0, 0), c, r)                     # tangent, behind
    assert hit and t == -5.0


def test_oracle_is_deterministic_and_thread_independent(oracle, scenes):
    a = oracle.render(scenes["medium"], 120, 67, 4, want_idx=True, nthreads=1)
    b = oracle.render(scenes["medium"], 120, 67, 4, want_idx=True, nthreads=0)
    assert np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["hit_idx"], b["hit_idx"])
    assert a["counters"] == b["counters"]


def test_synthetic_scene_generator_is_reproducible(rt, oracle, tmp_path):
    import gen_scene
    sph, lights, amb, cam = gen_scene.generate(300, 420)
    text = gen_scene.to_text(sph, lights, amb, cam)
    p = tmp_path / "syn.txt"
    p.write_text(text)
    sc = rt.load_scene(str(p))
    assert sc.nspheres == 300 and sc.nlights == 4
    assert np.array_equal(sc.spheres, sph) and np.array_equal(sc.camera, cam)
    assert hashlib.sha256(text.encode()).hexdigest() == hashlib.sha256(
        gen_scene.to_text(*gen_scene.generate(300, 420)).encode()).hexdigest()
    r = oracle.render(sc, 64, 36, 3)
    assert r["counters"]["hits"] > 0
