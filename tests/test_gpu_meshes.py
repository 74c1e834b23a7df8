"""GPU parity tests (-m gpu) of triangle meshes: mesh triangles are primitives of the flat
BVH next to spheres and cubes (csrc/ptb_kernels.cuh: exactCandidate, rayTriangle in ptb_math.cuh), traced by the
same kernels.  Everything goes through the C ABI and is compared with the CPU oracle on the same inputs.
Bar: hit geom index, t, normal and the inside / outside flag of every ray BIT-EXACT (integer + IEEE arithmetic
only); rendered images like the later bounces of the base path (1e-4 relative per pixel, sinf / cosf ulps)."""
import os

import numpy as np
import pytest

import orclib
import scenes
from common import pkg, same_bits, rel_err_lit, to_oracle_scene

pytestmark = pytest.mark.gpu

RADIANCE_RTOL = 1e-4
RES = (200, 160)


@pytest.fixture(scope="module")
def m():
    return pkg()


@pytest.fixture(scope="module")
def port():
    return orclib.load("port")


@pytest.fixture(scope="module")
def mesh_scene(m, tmp_path_factory):
    d = tmp_path_factory.mktemp("mesh_gpu")
    s = m.Scene.load(scenes.mesh_scene_files(str(d), res=RES, subdiv=2))
    return s, to_oracle_scene(s)


def _rays(osc, n, seed):
    rs = np.random.RandomState(seed)
    o = rs.uniform(-5, 5, (n, 3)).astype(np.float32) + np.float32([0, 5, 0])
    d = rs.normal(size=(n, 3)).astype(np.float32)
    d[::40, 0] = 0.0
    d[::55] = [0, -1, 0]
    for k, g in enumerate((6, 7, 8)):                   # origins inside each mesh: back faces, outside = 0
        sel = slice(k + 1, n, 7)
        o[sel] = osc.geoms["translation"][g] + rs.normal(size=(len(o[sel]), 3)).astype(np.float32) * 0.15
    return o, d


def _same_hits(got, want):
    return (np.array_equal(got[0], want[0]) and same_bits(got[1], want[1]) and same_bits(got[2], want[2])
            and np.array_equal(got[3], want[3]))


def test_closest_hit_bit_exact(m, port, mesh_scene):
    s, osc = mesh_scene
    assert len(s.triangles) == 320 + 12 + 20
    with m.Renderer(s, device=0) as r:
        assert r.stats()["bvh_nodes"] > 0                                   # meshes always go through the BVH
        o, d = r.stage_raygen()
        want = port.closest_hit(osc.geoms, o, d, osc)
        assert _same_hits(r.stage_intersect(o, d, m.ACCEL_BVH), want)
        assert _same_hits(r.stage_intersect(o, d, m.ACCEL_AUTO), want)
        assert set(np.unique(want[0])) >= {6, 7, 8, 9}
        o, d = _rays(osc, 300000, 31)
        want = port.closest_hit(osc.geoms, o, d, osc)
        got = r.stage_intersect(o, d, m.ACCEL_BVH)
        assert _same_hits(got, want)
        assert ((want[0] >= 6) & (want[0] <= 8) & (want[3] == 0)).sum() > 5000  # plenty of back-face hits
        with pytest.raises(m.PtbError, match="brute"):
            r.stage_intersect(o[:8], d[:8], m.ACCEL_BRUTE)
    with pytest.raises(m.PtbError, match="BVH"):
        m.Renderer(s, device=0, accel=m.ACCEL_BRUTE)


def test_ties_between_coincident_triangles_and_geoms(m, port, tmp_path):
    """the same triangles twice inside one mesh, the same mesh twice, and a cube face coplanar with a mesh face:
    the lowest triangle index wins inside a mesh, the lowest geom index across geoms (src/pathtrace.cu:192)"""
    v, f = scenes.box_quads()
    (tmp_path / "box2.obj").write_text(scenes.obj_text(v, f + f))              # every face twice
    objs = [("a", "mesh box2.obj", 1, "0 5 0", "0 0 0", "2 2 2"), ("b", "mesh box2.obj", 2, "0 5 0", "0 0 0", "2 2 2"),
            ("c", "cube", 3, "0 5 0", "0 0 0", "2 2 2"), ("d", "cube", 1, "0 0 0", "0 0 0", "10 .01 10")]
    (tmp_path / "ties.txt").write_text(scenes.emit(scenes._CORNELL_MATERIALS, objs, res=(64, 64), name="ties"))
    s = m.Scene.load(str(tmp_path / "ties.txt"))
    osc = to_oracle_scene(s)
    rs = np.random.RandomState(2)
    n = 100000
    o = rs.uniform(-4, 4, (n, 3)).astype(np.float32) + np.float32([0, 5, 0])
    d = (np.float32([0, 5, 0]) + rs.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)) - o
    o[::5] = np.float32([0, 5, 0]) + rs.uniform(-0.5, 0.5, (len(o[::5]), 3)).astype(np.float32)     # from inside
    want = port.closest_hit(osc.geoms, o, d, osc)
    with m.Renderer(s, device=0) as r:
        got = r.stage_intersect(o, d, m.ACCEL_BVH)
    assert _same_hits(got, want)
    assert (want[0] == 0).sum() > 20000 and (want[0] == 1).sum() < (want[0] == 0).sum() // 20


CONFIGS = [("fused_b1", dict(batch_iters=1)), ("fused_b4", dict(batch_iters=4)), ("fused_nograph", dict(batch_iters=2, use_graphs=0)),
           ("fused_cache", dict(batch_iters=2, first_bounce_cache=1)), ("staged", dict(pipeline=1, compaction=1)),
           ("staged_sort", dict(pipeline=1, compaction=1, sort_materials=1)), ("fused_aa_dof", dict(batch_iters=2, antialias=1, lens_radius=0.2, focal_distance=10.0))]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_render_against_oracle(m, port, mesh_scene, cfg):
    s, osc = mesh_scene
    okw = {k: v for k, v in cfg[1].items() if k in ("antialias", "lens_radius", "focal_distance")}
    oimg, oalive, oseg = port.render(osc, orclib.MODE_FULL, 1, 4, **okw)
    with m.Renderer(s, device=0, bsdf=1, **cfg[1]) as r:
        r.render(1, 4)
        img, st = r.download(), r.stats()
    alive = np.array(st["alive_per_depth"])
    exact = 1 if okw else 2
    assert list(alive[:exact]) == list(oalive[:exact]), (alive, oalive)
    assert np.all(np.abs(alive - oalive) <= np.maximum(8, oalive // 2000)), (alive, oalive)
    rel, lit = rel_err_lit(img, oimg)
    frac = float((rel <= RADIANCE_RTOL).mean())
    assert frac >= 0.998, "%s: only %.5f of lit pixels within 1e-4" % (cfg[0], frac)
    assert np.allclose(img.astype(np.float64).sum(0), oimg.astype(np.float64).sum(0), rtol=3e-3)


def test_large_mesh_bvh_against_oracle(m, port, tmp_path):
    """81 920-triangle icosphere plus analytic primitives: BVH traversal against the oracle's exhaustive loop on a
    ray sample the oracle finishes in seconds, and the rendered image against the analytic sphere it approximates."""
    v, f = scenes.icosphere(6)
    assert len(f) == 81920
    (tmp_path / "big.obj").write_text(scenes.obj_text(v, f))
    objs = list(scenes._CORNELL_OBJECTS)
    objs[6] = ("Fine icosphere", "mesh big.obj", 2, "-1 4 -1", "10 20 30", "3 3 3")
    (tmp_path / "big.txt").write_text(scenes.emit(scenes._CORNELL_MATERIALS, objs, res=(160, 160), name="big"))
    s = m.Scene.load(str(tmp_path / "big.txt"))
    osc = to_oracle_scene(s)
    with m.Renderer(s, device=0, bsdf=0) as r:
        o, d = r.stage_raygen()
        # every 7th camera ray plus (up to 400 of) the rays the BVH reports on the mesh, so that the sample holds
        # a few hundred triangle hits whatever share of the frame the mesh covers
        on_mesh = np.flatnonzero(r.stage_intersect(o, d, m.ACCEL_BVH)[0] == 6)
        sel = np.union1d(np.arange(0, len(o), 7), on_mesh[:: max(1, len(on_mesh) // 400)])
        want = port.closest_hit(osc.geoms, o[sel], d[sel], osc)
        got = r.stage_intersect(o[sel], d[sel], m.ACCEL_BVH)
        assert _same_hits(got, want) and (want[0] == 6).sum() > 100
        r.render(1, 64)
        img = r.download() / 64
    # the same scene with the analytic sphere (diffuse red): the two images agree up to Monte-Carlo noise
    objs[6] = ("Sphere", "sphere", 2, "-1 4 -1", "10 20 30", "3 3 3")
    (tmp_path / "ana.txt").write_text(scenes.emit(scenes._CORNELL_MATERIALS, objs, res=(160, 160), name="ana"))
    with m.Renderer(m.Scene.load(str(tmp_path / "ana.txt")), device=0, bsdf=0) as r:
        r.render(1, 64)
        ana = r.download() / 64
    a, b = img.reshape(160, 160, 3), ana.reshape(160, 160, 3)
    blocks = lambda x: x.reshape(20, 8, 20, 8, 3).mean((1, 3))
    err = np.sqrt(((blocks(a) - blocks(b)) ** 2).mean()) / np.sqrt((blocks(b) ** 2).mean())
    assert err < 0.05, err


def test_update_scene_with_meshes(m, port, mesh_scene):
    s, osc = mesh_scene
    moved = m.Scene(s.geoms.copy(), s.materials, s.camera, s.iterations, s.traceDepth, s.imageName, s.triangles, s.mesh_ranges)
    moved.triangles = (s.triangles * np.float32(0.8)).astype(np.float32)        # shrink every mesh in object space
    osc2 = to_oracle_scene(moved)
    with m.Renderer(s, device=0, bsdf=1, batch_iters=1) as r:
        r.render(1, 1)
        r.update_scene(moved)
        o, d = r.stage_raygen()
        assert _same_hits(r.stage_intersect(o, d, m.ACCEL_BVH), port.closest_hit(osc2.geoms, o, d, osc2))
        fewer = m.Scene(s.geoms, s.materials, s.camera, s.iterations, s.traceDepth, s.imageName, s.triangles[:-1], s.mesh_ranges)
        with pytest.raises(m.PtbError):
            r.update_scene(fewer)
