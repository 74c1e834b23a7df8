"""GPU parity tests (-m gpu) of the camera / material extensions: stochastic
antialiasing, thin-lens depth of field, imperfect specular.  The CUDA path (EXT kernel instantiations,
csrc/ptb_kernels.cuh) is called through the C ABI and compared with the CPU oracle on the same seeds.

Tolerances: the jitter is integer RNG + IEEE adds, so antialiased camera rays are BIT-EXACT; the lens and
the Phong lobe go through sincosf / powf (libdevice vs glibc: ulp-level), so those rays are compared within
DIR_ATOL and images within 1e-4 relative per pixel on >= 99.8 % of the lit pixels, as for the later
bounces of the base path (BASELINE.json north_star).
"""
import numpy as np
import pytest

import orclib
import scenes
from common import pkg, same_bits, rel_err_lit

pytestmark = pytest.mark.gpu

RADIANCE_RTOL = 1e-4
DIR_ATOL = 4e-7
RES = (200, 160)
LENS_CASES = {"aa": dict(antialias=1), "dof": dict(lens_radius=0.35, focal_distance=9.5),
              "aadof": dict(antialias=1, lens_radius=0.2, focal_distance=11.0)}


@pytest.fixture(scope="module")
def m():
    return pkg()


@pytest.fixture(scope="module")
def port():
    return orclib.load("port")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("ext_gpu")
    out = {}
    for name, text in (("cornell", scenes.cornell_text(res=RES)), ("glossy", scenes.glossy_text(res=RES)),
                       ("glass", scenes.glass_text(res=RES))):
        p = d / (name + ".txt")
        p.write_text(text)
        out[name] = str(p)
    return out


@pytest.mark.parametrize("tag", sorted(LENS_CASES))
def test_camera_rays(m, port, files, tag):
    kw = LENS_CASES[tag]
    s = m.Scene.load(files["cornell"])
    osc = port.load_scene(files["cornell"])
    with m.Renderer(s, device=0, **kw) as r:
        for it in (1, 2, 4095):
            o, d = r.stage_raygen(it)
            wo, wd = port.generate_rays_lens(osc.camera, it, **kw)
            if tag == "aa":
                assert same_bits(o, wo) and same_bits(d, wd)
            else:
                assert np.abs(o - wo).max() <= 1e-6 and np.abs(d - wd).max() <= DIR_ATOL      # one ulp of a lens-point coordinate (sincosf)
                assert (o.view(np.uint32) == wo.view(np.uint32)).all(1).mean() > 0.5
    # extensions off: stage_raygen(iter) is the pinhole ray for every iteration
    with m.Renderer(s, device=0) as r:
        a, b = r.stage_raygen(1), r.stage_raygen(9)
        w = port.generate_rays(osc.camera)
        assert same_bits(a[0], w[0]) and same_bits(a[1], w[1]) and same_bits(b[1], w[1])


CONFIGS = [("fused_b1", dict(batch_iters=1)), ("fused_b4", dict(batch_iters=4)), ("fused_bvh", dict(batch_iters=2, accel=2)),
           ("fused_nograph", dict(batch_iters=2, use_graphs=0)), ("staged", dict(pipeline=1, compaction=1)),
           ("staged_sort", dict(pipeline=1, compaction=1, sort_materials=1))]


def _compare(img, st, oimg, oalive, exact_depths, frac_min=0.998):
    alive = np.array(st["alive_per_depth"])
    assert list(alive[:exact_depths]) == list(oalive[:exact_depths]), (alive, oalive)
    assert np.all(np.abs(alive - oalive) <= np.maximum(8, oalive // 2000)), (alive, oalive)
    rel, lit = rel_err_lit(img, oimg)
    frac = float((rel <= RADIANCE_RTOL).mean())
    assert frac >= frac_min, "only %.5f of lit pixels within 1e-4" % frac
    assert np.allclose(img.astype(np.float64).sum(0), oimg.astype(np.float64).sum(0), rtol=3e-3)


@pytest.mark.parametrize("tag", sorted(LENS_CASES))
@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_lens_render_against_oracle(m, port, files, tag, cfg):
    kw = LENS_CASES[tag]
    s = m.Scene.load(files["glass"])
    osc = port.load_scene(files["glass"])
    oimg, oalive, oseg = port.render(osc, orclib.MODE_FULL, 1, 4, **kw)
    with m.Renderer(s, device=0, bsdf=1, **cfg[1], **kw) as r:
        r.render(1, 4)
        img, st = r.download(), r.stats()
    # antialiased camera rays are bit-exact, so the bounce-0 and bounce-1 path counts are too; with the lens
    # (sincosf) a handful of camera rays may graze differently
    _compare(img, st, oimg, oalive, 2 if tag == "aa" else 1)


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_glossy_render_against_oracle(m, port, files, cfg):
    s = m.Scene.load(files["glossy"])
    osc = port.load_scene(files["glossy"])
    oimg, oalive, oseg = port.render(osc, orclib.MODE_FULL, 1, 4)
    with m.Renderer(s, device=0, bsdf=1, **cfg[1]) as r:
        r.render(1, 4)
        img, st = r.download(), r.stats()
    _compare(img, st, oimg, oalive, 2)
    # golden-diffuse mode ignores SPECEX: the non-EXT kernels run and match the oracle's diffuse loop
    dimg, dalive, _ = port.render(osc, orclib.MODE_GOLDEN_DIFFUSE, 1, 2)
    with m.Renderer(s, device=0, bsdf=0, **cfg[1]) as r:
        r.render(1, 2)
        _compare(r.download(), r.stats(), dimg, dalive, 2, 0.9995)


def test_glossy_shade_step(m, port, files):
    s = m.Scene.load(files["glossy"])
    osc = port.load_scene(files["glossy"])
    o, d = port.generate_rays(osc.camera)
    hit, t, n, outside = port.closest_hit(osc.geoms, o, d)
    P = osc.pixels
    mid = np.where(hit >= 0, osc.geoms["materialid"][np.maximum(hit, 0)], 0).astype(np.int32)
    rs = np.random.RandomState(6)
    it = rs.randint(1, 5000, P).astype(np.int32)
    idx = np.arange(P, dtype=np.int32)
    dep = rs.randint(0, 8, P).astype(np.int32)
    col = rs.uniform(0.1, 1, (P, 3)).astype(np.float32)
    rem = rs.randint(1, 9, P).astype(np.int32)
    wo, wd, wc, wrem = port.shade(osc.materials, 1, hit, mid, t, n, outside, it, idx, dep, o, d, col, rem)
    with m.Renderer(s, device=0, bsdf=1) as r:
        go, gd, gc, grem = r.stage_shade(hit, t, n, outside, it, idx, dep, o, d, col, rem)
    assert np.array_equal(grem, wrem) and same_bits(gc, wc)
    alive = wrem > 0
    assert same_bits(go[alive], wo[alive])
    glossy = alive & (mid >= 5)
    assert glossy.sum() > 1000
    # powf / sincosf ulps, amplified for the tight lobe by 1/sin(theta) in sqrt(1 - up^2)
    assert np.abs(gd[glossy] - wd[glossy]).max() <= 2e-5
    assert np.median(np.abs(gd[glossy] - wd[glossy]).max(1)) <= 1e-7
    assert np.abs(gd[alive & ~glossy] - wd[alive & ~glossy]).max() <= DIR_ATOL


def test_first_bounce_cache_is_refused_with_a_jittered_camera(m, files):
    s = m.Scene.load(files["cornell"])
    for kw in (dict(antialias=1), dict(lens_radius=0.1, focal_distance=5.0)):
        with pytest.raises(m.PtbError, match="firstBounceCache"):
            m.Renderer(s, device=0, first_bounce_cache=1, **kw)
    with pytest.raises(m.PtbError, match="focalDistance"):
        m.Renderer(s, device=0, lens_radius=0.1)


def test_update_scene_switches_kernel_variant(m, port, files):
    """a material edit that introduces SPECEX > 0 must re-capture the CUDA graph with the EXT kernels"""
    plain, glossy = m.Scene.load(files["cornell"]), m.Scene.load(files["cornell"])
    glossy.materials["specularExponent"][4] = 30.0
    osc = port.load_scene(files["cornell"])
    want_plain = port.render(osc, 1, 1, 2)
    osc.materials["specExponent"][4] = 30.0
    want_glossy = port.render(osc, 1, 1, 2)
    assert not same_bits(want_plain[0], want_glossy[0])
    with m.Renderer(plain, device=0, bsdf=1, batch_iters=1) as r:
        for scene, want in ((plain, want_plain), (glossy, want_glossy), (plain, want_plain)):
            r.update_scene(scene)
            r.render(1, 2)
            _compare(r.download(), r.stats(), want[0], want[1], 2)


def test_antialiasing_softens_edges_and_keeps_energy(m, files):
    """stochastic AA is a box filter over the pixel: the image mean stays, silhouettes get intermediate values"""
    s = m.Scene.load(files["cornell"])
    n = 256
    out = []
    for aa in (0, 1):
        with m.Renderer(s, device=0, bsdf=0, antialias=aa) as r:
            r.render(1, n)
            out.append(r.download().reshape(RES[1], RES[0], 3) / n)
    a, b = out
    assert abs(a.mean() - b.mean()) / a.mean() < 0.02
    # bounce-0 coverage of the light (emissive hit => radiance 5 every iteration, no noise): binary without AA
    la, lb = a[..., 0] >= 4.99, (b[..., 0] > 0.25 * 5) & (b[..., 0] < 0.75 * 5)
    assert la.any() and lb.any()
