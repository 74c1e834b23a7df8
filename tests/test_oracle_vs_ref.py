"""CPU tests: the plain-C restatement must agree BIT FOR BIT with the reference's own functions
compiled here (oracle/_ref/libptref.so).  Skipped where that library was not built (it needs
/root/reference at build time; the prebuilt .so travels to the GPU box, where these tests also run)."""
import os

import numpy as np
import pytest

import orclib
from common import same_bits, REF_ROOT

ref = orclib.load("reference")
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref/libptref.so not built")


@pytest.fixture(scope="module")
def port():
    return orclib.load("port")


def test_kinds(port):
    assert ref.kind == "reference" and port.kind == "port"


@pytest.mark.skipif(not os.path.isdir(REF_ROOT), reason="/root/reference not mounted")
def test_emitted_scenes_equal_the_reference_files(scene_dir):
    for name in ("cornell", "sphere"):
        a = ref.load_scene("%s/scenes/%s.txt" % (REF_ROOT, name), fixup=False)
        b = ref.load_scene("%s/%s.txt" % (scene_dir, name), fixup=False)
        assert a.geoms.tobytes() == b.geoms.tobytes() and a.materials.tobytes() == b.materials.tobytes()
        assert a.camera.tobytes() == b.camera.tobytes()
        assert (a.iterations, a.depth, a.name) == (b.iterations, b.depth, b.name)


@pytest.mark.parametrize("name", ["cornell", "sphere", "glass"])
def test_loader_and_fixup(port, scene_dir, name):
    for fix in (False, True):
        a = port.load_scene("%s/%s.txt" % (scene_dir, name), fixup=fix)
        b = ref.load_scene("%s/%s.txt" % (scene_dir, name), fixup=fix)
        assert a.geoms.tobytes() == b.geoms.tobytes()
        assert a.materials.tobytes() == b.materials.tobytes()
        assert a.camera.tobytes() == b.camera.tobytes()


def test_transforms_random(port):
    rs = np.random.RandomState(7)
    for i in range(3000):
        t, r, s = rs.uniform(-50, 50, 3), rs.uniform(-360, 360, 3), rs.uniform(0.01, 10, 3)
        if i % 5 == 0:
            r = np.round(r / 90) * 90
        for a, b in zip(port.build_transform(t, r, s), ref.build_transform(t, r, s)):
            assert a.tobytes() == b.tobytes()


def test_rng_random(port):
    rs = np.random.RandomState(8)
    for _ in range(2000):
        a, b, c = int(rs.randint(0, 1 << 22)), int(rs.randint(0, 1 << 24)), int(rs.randint(0, 512))
        assert same_bits(port.rng_u01(a, b, c, 3), ref.rng_u01(a, b, c, 3))


@pytest.mark.parametrize("name,mode", [("cornell", 0), ("cornell", 1), ("sphere", 1), ("glass", 1)])
def test_hot_path_bit_for_bit(port, scene_dir, name, mode):
    a = port.load_scene("%s/%s.txt" % (scene_dir, name))
    b = ref.load_scene("%s/%s.txt" % (scene_dir, name))
    o1, d1 = port.generate_rays(a.camera)
    o2, d2 = ref.generate_rays(b.camera)
    assert same_bits(o1, o2) and same_bits(d1, d2)
    h1 = port.closest_hit(a.geoms, o1, d1)
    h2 = ref.closest_hit(b.geoms, o2, d2)
    assert np.array_equal(h1[0], h2[0]) and same_bits(h1[1], h2[1]) and same_bits(h1[2], h2[2]) and np.array_equal(h1[3], h2[3])
    i1, al1, s1 = port.render(a, mode, 1, 2)
    i2, al2, s2 = ref.render(b, mode, 1, 2)
    assert same_bits(i1, i2) and np.array_equal(al1, al2) and s1 == s2


def test_random_rays_against_rotated_geoms(port, scene_dir):
    # rays from inside and outside, including axis-parallel directions (division by +-0 in the slab test)
    s = port.load_scene(scene_dir + "/glass.txt")
    rs = np.random.RandomState(9)
    n = 200000
    o = rs.uniform(-6, 6, (n, 3)).astype(np.float32) + np.float32([0, 5, 0])
    d = rs.normal(size=(n, 3)).astype(np.float32)
    d[::50, 0] = 0.0
    d[::75, 1] = 0.0
    d[::90] = [0, 0, -1]
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    h1 = port.closest_hit(s.geoms, o, d)
    h2 = ref.closest_hit(s.geoms, o, d)
    assert np.array_equal(h1[0], h2[0]) and same_bits(h1[1], h2[1]) and same_bits(h1[2], h2[2]) and np.array_equal(h1[3], h2[3])


def test_shade_step_bit_for_bit(port, scene_dir):
    s = port.load_scene(scene_dir + "/glass.txt")
    o, d = port.generate_rays(s.camera)
    hit, t, n, outside = port.closest_hit(s.geoms, o, d)
    P = s.pixels
    mid = np.where(hit >= 0, s.geoms["materialid"][np.maximum(hit, 0)], 0).astype(np.int32)
    it = np.full(P, 3, np.int32)
    idx = np.arange(P, dtype=np.int32)
    dep = np.zeros(P, np.int32)
    col = np.ones((P, 3), np.float32)
    rem = np.full(P, 8, np.int32)
    for mode in (0, 1):
        a = port.shade(s.materials, mode, hit, mid, t, n, outside, it, idx, dep, o, d, col, rem)
        b = ref.shade(s.materials, mode, hit, mid, t, n, outside, it, idx, dep, o, d, col, rem)
        for x, y in zip(a[:3], b[:3]):
            assert same_bits(x, y)
        assert np.array_equal(a[3], b[3])
