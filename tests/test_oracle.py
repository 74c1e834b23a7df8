"""CPU tests (-m "not gpu"): pin the plain-C oracle (oracle/pt_oracle.c) to the committed fixtures
that the reference-linked build produced (tests/golden/make_golden.py) and to the known answers of
SURVEY.md 8(c) / 9."""
import numpy as np
import pytest

import orclib
from common import golden, crc, same_bits, densify


@pytest.fixture(scope="module")
def port():
    o = orclib.load("port")
    assert o is not None, "oracle/libptoracle.so missing: run __graft_entry__.build()"
    assert o.kind == "port"
    return o


def test_scene_files_parse_to_reference_pod(port, scene_dir):
    g = golden()
    for name in ("cornell", "sphere"):
        raw = port.load_scene("%s/%s.txt" % (scene_dir, name), fixup=False)
        fixed = port.load_scene("%s/%s.txt" % (scene_dir, name), fixup=True)
        assert raw.geoms.tobytes() == g["scene_%s_geoms" % name].tobytes()
        assert raw.materials.tobytes() == g["scene_%s_materials" % name].tobytes()
        assert raw.camera.tobytes() == g["scene_%s_camera_raw" % name].tobytes()
        assert fixed.camera.tobytes() == g["scene_%s_camera_fixed" % name].tobytes()
        assert [raw.iterations, raw.depth] == list(g["scene_%s_meta" % name])
    gl = port.load_scene(scene_dir + "/glass.txt")
    assert gl.geoms.tobytes() == g["scene_glass_geoms"].tobytes()
    assert gl.camera.tobytes() == g["scene_glass_camera_fixed"].tobytes()


def test_camera_fixup_known_values(port, scene_dir):
    # SURVEY.md 3.1 [probe] values after the fix-up
    cam = port.load_scene(scene_dir + "/cornell.txt").camera[0]
    assert np.allclose(cam["position"], [0, 4.99999952, 10.5], rtol=0, atol=1e-7)
    assert cam["view"][2] == -1.0 and abs(cam["view"][1] - 4.37113883e-08) < 1e-14
    assert list(cam["right"]) == [1.0, 0.0, -0.0]
    assert np.allclose(cam["pixelLength"], [0.0025, 0.0025], atol=1e-9)
    raw = port.load_scene(scene_dir + "/cornell.txt", fixup=False).camera[0]
    assert np.isnan(raw["right"]).all()          # src/scene.cpp:138 uses `view` before it is set


def test_rng_known_answers(port):
    g = golden()
    # SURVEY.md 8(a) a8 check value
    u = port.rng_u01(1, 0, 0, 2)
    assert abs(float(u[0]) - 0.00288156001) < 1e-10 and abs(float(u[1]) - 0.0958057344) < 1e-9
    for (a, b, c), want in zip(g["rng_triples"], g["rng_u01"]):
        assert same_bits(port.rng_u01(int(a), int(b), int(c), 4), want)


def test_transforms_bit_exact(port):
    g = golden()
    trs, mats = g["xf_trs"], g["xf_matrices"]
    for i in range(len(trs)):
        got = port.build_transform(trs[i, 0], trs[i, 1], trs[i, 2])
        for k in range(3):
            assert same_bits(got[k], mats[i, k]), (i, k)


def test_first_bounce_pins(port, scene_dir):
    g = golden()
    s = port.load_scene(scene_dir + "/cornell.txt")
    o, d = port.generate_rays(s.camera)
    hit, t, n, outside = port.closest_hit(s.geoms, o, d)
    # no-FMA pins
    assert int((hit < 0).sum()) == 109612
    assert [int((hit == k).sum()) for k in range(7)] == [7229, 116044, 108545, 57274, 116227, 116294, 8775]
    assert abs(t[hit >= 0].astype(np.float64).sum() - 5667910.279906) < 1e-5
    assert np.array_equal(hit.astype(np.int8), g["fb_cornell_hit"])
    assert [crc(o), crc(d), crc(t), crc(n), crc(outside)] == list(g["fb_cornell_crc"])
    # centre ray: sphere g6 at t = 10.9997, n ~ (0.666622, 0.666622, 0.33351);
    # with the sphere removed the back wall g3 answers at t = 15.495 with n ~ (4.37e-8, 0, 1)
    cam = s.camera[0]
    ch, ct, cn, _ = port.closest_hit(s.geoms, cam["position"][None, :], cam["view"][None, :])
    assert ch[0] == 6 and abs(ct[0] - 10.9997) < 1e-3 and np.allclose(cn[0], [0.666622, 0.666622, 0.33351], atol=1e-5)
    ch, ct, cn, _ = port.closest_hit(s.geoms[:6], cam["position"][None, :], cam["view"][None, :])
    assert ch[0] == 3 and abs(ct[0] - 15.495) < 1e-3 and abs(cn[0, 2] - 1.0) < 1e-6
    for name in ("sphere", "glass"):
        s = port.load_scene("%s/%s.txt" % (scene_dir, name))
        o, d = port.generate_rays(s.camera)
        hit, t, n, outside = port.closest_hit(s.geoms, o, d)
        assert np.array_equal(hit.astype(np.int8), g["fb_%s_hit" % name])
        assert [crc(o), crc(d), crc(t), crc(n), crc(outside)] == list(g["fb_%s_crc" % name])
    s = port.load_scene(scene_dir + "/sphere.txt")
    o, d = port.generate_rays(s.camera)
    hit = port.closest_hit(s.geoms, o, d)[0]
    assert int((hit < 0).sum()) == 628376 and int((hit == 0).sum()) == 11624


def test_sincos_restatement_exhaustive(port):
    """The double-precision restatement of glibc's sinf / cosf that the CUDA path mirrors
    (csrc/ptb_math.cuh::sincosExact, oracle/pt_oracle.c::sincosRestated) against THIS machine's libm on every
    float in [0, 2 pi] -- the whole range `around = u * TWO_PI`, u in [0, 1] (src/interactions.h:19) and the lens
    angle can take: 1 086 918 620 arguments, zero disagreements (about 2 s on 8 cores)."""
    two_pi = np.float32(6.2831853071795864769252867665590057683943)
    assert port.sincos_mismatches(0, int(two_pi.view(np.uint32))) == 0
    # and a little beyond the sampler range (reduce_fast is valid below 120)
    a = np.float32([7.0, 31.4, 100.0, 119.9])
    assert same_bits(np.stack(port.sincos(a, True)), np.stack(port.sincos(a, False)))


def test_hemisphere_vectors(port):
    g = golden()
    seed = g["hemi_seed"]
    got = port.hemisphere(g["hemi_normals"], seed[:, 0], seed[:, 1], seed[:, 2])
    # same libm on the same machine => bit-exact; elsewhere sinf/cosf may differ by an ulp
    assert np.allclose(got, g["hemi_dirs"], rtol=0, atol=3e-7)
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)


@pytest.mark.parametrize("tag,scene,mode", [("cornell_diffuse", "cornell", 0), ("cornell_full", "cornell", 1), ("glass_full", "glass", 1)])
def test_normative_loop_iter1(port, scene_dir, tag, scene, mode):
    g = golden()
    s = port.load_scene("%s/%s.txt" % (scene_dir, scene))
    img, alive, seg = port.render(s, mode, 1, 1)
    want_alive = g["loop_%s_alive" % tag]
    # depth 0/1 are exact everywhere; deeper counts depend on libm's sinf/cosf ulps
    assert list(alive[:2]) == list(want_alive[:2])
    assert np.all(np.abs(alive - want_alive) <= np.maximum(3, want_alive // 2000))
    assert seg == alive.sum()
    want = densify(g["loop_%s_idx" % tag], g["loop_%s_rgb" % tag], s.pixels)
    same = (img.view(np.uint32) == want.view(np.uint32)).all(1).mean()
    assert same > 0.9995, same
    assert np.allclose(img.astype(np.float64).sum(0), g["loop_%s_sum" % tag], rtol=2e-3)


def test_known_answers_survey_section9(port, scene_dir):
    # parity-build counts for iter = 1; exact on the machine that made them
    s = port.load_scene(scene_dir + "/cornell.txt")
    _, alive0, seg0 = port.render(s, 0, 1, 1)
    _, alive1, seg1 = port.render(s, 1, 1, 1)
    assert list(alive0[:2]) == [640000, 523159] and list(alive1[:2]) == [640000, 523159]
    want0 = np.array([640000, 523159, 360703, 278408, 221680, 179196, 145906, 119309])
    want1 = np.array([640000, 523159, 362446, 280469, 223266, 180236, 146412, 119279])
    assert np.all(np.abs(alive0 - want0) <= 60) and np.all(np.abs(alive1 - want1) <= 60)
    assert abs(seg0 - 2468361) < 300 and abs(seg1 - 2475267) < 300


def test_small_image_8_iterations(port, scene_dir, tmp_path):
    import scenes
    g = golden()
    p = tmp_path / "small.txt"
    p.write_text(scenes.cornell_text(res=(96, 64)))
    s = port.load_scene(str(p))
    for tag, mode in (("cornell_diffuse", 0), ("cornell_full", 1)):
        img, alive, seg = port.render(s, mode, 1, 8)
        want = g["small_%s_img8" % tag]
        assert (img.view(np.uint32) == want.view(np.uint32)).all(1).mean() > 0.995
        assert list(alive[:2]) == list(g["small_%s_alive8" % tag][:2])


def test_accumulation_is_additive_over_iteration_ranges(port, scene_dir, tmp_path):
    import scenes
    p = tmp_path / "tiny.txt"
    p.write_text(scenes.cornell_text(res=(48, 48)))
    s = port.load_scene(str(p))
    whole, alive, seg = port.render(s, 1, 1, 6)
    part = np.zeros_like(whole)
    port.render(s, 1, 1, 3, image=part)
    port.render(s, 1, 4, 3, image=part)
    assert same_bits(whole, part)          # iterations accumulate in ascending order


def test_edge_cases(port, tmp_path):
    import scenes
    # empty scene: everything misses; ragged resolution; depth 1
    txt = scenes.emit(scenes._SPHERE_MATERIALS, [], res=(7, 5), depth=1, name="empty")
    p = tmp_path / "empty.txt"
    p.write_text(txt)
    s = port.load_scene(str(p))
    assert len(s.geoms) == 0 and s.width == 7 and s.height == 5
    img, alive, seg = port.render(s, 1, 1, 2)
    assert seg == 70 and not img.any()
    assert port.closest_hit(s.geoms, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))[0].shape == (0,)
    with pytest.raises(IOError):
        port.load_scene(str(tmp_path / "does_not_exist.txt"))
    # CRLF line endings and a file without a trailing newline parse identically (safeGetline)
    base = scenes.cornell_text()
    (tmp_path / "crlf.txt").write_text(base.replace("\n", "\r\n"))
    (tmp_path / "notrail.txt").write_text(base.rstrip("\n"))
    a = port.load_scene(str(tmp_path / "crlf.txt"))
    b = port.load_scene(str(tmp_path / "notrail.txt"))
    c = port.load_scene(str(tmp_path / "crlf.txt"))
    assert a.geoms.tobytes() == b.geoms.tobytes() == c.geoms.tobytes() and len(a.geoms) == 7
