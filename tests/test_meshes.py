"""CPU tests (-m "not gpu") of triangle meshes: the `mesh <file.obj>` object type the scene
format reserves (INSTRUCTION.md:291-293) and the per-triangle test the assignment names, glm::intersectRayTriangle
(INSTRUCTION.md:123-127).  The reference implements neither; the mesh test is authored in the style of its two
tests.  Pins: the plain-C oracle against fixtures made by the reference-linked build (which
calls the vendored glm function itself), the authored test against the reference's OWN boxIntersectionTest on
a cube given as 12 triangles, and the product's host loader (OBJ reader) against an independent parser.
The GPU side is in tests/test_gpu_meshes.py."""
import os

import numpy as np
import pytest

import orclib
import scenes
from common import HERE, pkg, same_bits, to_oracle_scene

RES = (96, 64)


@pytest.fixture(scope="module")
def port():
    return orclib.load("port")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "extension_vectors.npz"))


@pytest.fixture(scope="module")
def mesh_scene(tmp_path_factory):
    d = tmp_path_factory.mktemp("mesh")
    path = scenes.mesh_scene_files(str(d), res=RES)
    return str(d), path, pkg().Scene.load(path)


def test_host_loader_reads_mesh_objects(mesh_scene):
    d, path, s = mesh_scene
    assert list(s.geoms["type"]) == [1] * 6 + [2, 2, 2, 0]
    assert s.mesh_ranges.tolist()[6:9] == [[0, 80], [80, 12], [92, 20]]
    assert all(c == 0 for f, c in s.mesh_ranges.tolist()[:6] + s.mesh_ranges.tolist()[9:])
    # vertex data equals an independent reader's, incl. fanned quads, v/vt/vn references and negative indices
    want = np.concatenate([scenes.parse_obj(open(os.path.join(d, n)).read()) for n in ("ico.obj", "box.obj", "ico0.obj")])
    assert same_bits(s.triangles, want) and len(want) == 112
    # transforms of a mesh geom are built like every other geom's (src/scene.cpp:82-85)
    port = orclib.load("port")
    g = s.geoms[7]
    T, inv, invT = port.build_transform(g["translation"], g["rotation"], g["scale"])
    assert same_bits(g["transform"], T) and same_bits(g["inverseTransform"], inv) and same_bits(g["invTranspose"], invT)


def test_loader_edge_cases(tmp_path):
    m = pkg()
    base = scenes.emit(scenes._CORNELL_MATERIALS, [("m", "mesh missing.obj", 1, "0 0 0", "0 0 0", "1 1 1")], res=(8, 8))
    (tmp_path / "missing.txt").write_text(base)
    with pytest.raises(m.PtbError):                     # unreadable OBJ: the scene does not load
        m.Scene.load(str(tmp_path / "missing.txt"))
    (tmp_path / "empty.obj").write_text("# nothing\nv 0 0 0\nf 1 5 9\n")     # out-of-range face is dropped
    (tmp_path / "e.txt").write_text(base.replace("missing.obj", "empty.obj"))
    s = m.Scene.load(str(tmp_path / "e.txt"))
    assert s.geoms["type"][0] == 2 and len(s.triangles) == 0 and s.mesh_ranges.tolist() == [[0, 0]]
    # a scene without meshes carries empty triangle arrays
    (tmp_path / "c.txt").write_text(scenes.cornell_text(res=(8, 8)))
    s = m.Scene.load(str(tmp_path / "c.txt"))
    assert len(s.triangles) == 0 and s.mesh_ranges.shape == (7, 2) and not s.mesh_ranges.any()


def test_oracle_matches_the_reference_linked_fixtures(port, gold, mesh_scene):
    osc = to_oracle_scene(mesh_scene[2])
    assert same_bits(osc.triangles, gold["mesh_triangles"]) and np.array_equal(osc.mesh_ranges, gold["mesh_ranges"])
    o, d = port.generate_rays(osc.camera)
    hit, t, n, outside = port.closest_hit(osc.geoms, o, d, osc)
    assert np.array_equal(hit, gold["mesh_hit"]) and same_bits(t, gold["mesh_t"]) and same_bits(n, gold["mesh_n"])
    assert np.array_equal(outside, gold["mesh_outside"])
    assert set(np.unique(hit)) >= {6, 7, 8}                                     # every mesh is seen by the camera
    img, alive, seg = port.render(osc, orclib.MODE_FULL, 1, 2)
    assert same_bits(img, gold["img_mesh"]) and np.array_equal(alive, gold["alive_mesh"])


def test_port_equals_reference_linked_build_on_meshes(port, mesh_scene):
    ref = orclib.load("reference")
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    osc = to_oracle_scene(mesh_scene[2])
    rs = np.random.RandomState(12)
    n = 60000
    o = rs.uniform(-5, 5, (n, 3)).astype(np.float32) + np.float32([0, 5, 0])
    d = rs.normal(size=(n, 3)).astype(np.float32)
    d[::40, 0] = 0.0
    d[::55] = [0, -1, 0]
    o[::3] = osc.geoms["translation"][8] + rs.normal(size=(len(o[::3]), 3)).astype(np.float32) * 0.2   # inside the glass mesh
    a, b = port.closest_hit(osc.geoms, o, d, osc), ref.closest_hit(osc.geoms, o, d, osc)
    assert np.array_equal(a[0], b[0]) and same_bits(a[1], b[1]) and same_bits(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert (a[3][a[0] == 8] == 0).sum() > 1000                                  # back faces: outside = 0
    ia, ib = port.render(osc, 1, 3, 2), ref.render(osc, 1, 3, 2)
    assert same_bits(ia[0], ib[0]) and np.array_equal(ia[1], ib[1]) and ia[2] == ib[2]


def test_triangulated_cube_agrees_with_the_reference_box_test(port):
    """The authored mesh test on the unit cube as 12 triangles against the reference's own boxIntersectionTest
    (src/intersections.h:48-90) under the same transform: same hit / miss decisions, distances and normals up to
    float rounding (different arithmetic, same geometry), same `outside` flag."""
    tri = scenes.parse_obj(scenes.obj_text(*scenes.box_quads()))
    rs = np.random.RandomState(4)
    for trs in (([0, 0, 0], [0, 0, 0], [1, 1, 1]), ([1.5, -2, 0.5], [20, 35, 10], [1.5, 3, 0.7])):
        T, inv, invT = port.build_transform(*(np.float32(x) for x in trs))
        g = np.zeros(2, orclib.GEOM_DT)
        for k, ty in enumerate((1, 2)):
            g[k]["type"], g[k]["transform"], g[k]["inverseTransform"], g[k]["invTranspose"] = ty, T, inv, invT
        n = 80000
        o = rs.uniform(-4, 4, (n, 3)).astype(np.float32) + np.float32(trs[0])
        o[::4] = np.float32(trs[0]) + rs.uniform(-0.2, 0.2, (len(o[::4]), 3)).astype(np.float32)   # origins inside the cube
        d = rs.normal(size=(n, 3)).astype(np.float32)
        box = port.closest_hit(g[:1], o, d)
        mesh = port.closest_hit(g[1:], o, d, (tri, np.int32([[0, 12]])))
        agree = (box[0] >= 0) == (mesh[0] >= 0)
        assert agree.mean() > 0.9995                                            # disagreements: rays grazing an edge
        both = (box[0] >= 0) & (mesh[0] >= 0)
        assert both.sum() > 10000
        assert np.abs(box[1][both] - mesh[1][both]).max() < 2e-4
        assert np.array_equal(box[3][both], mesh[3][both])                      # inside / outside
        close = np.abs(box[2][both] - mesh[2][both]).max(1) < 1e-4              # same face normal (corner hits may pick a neighbour)
        assert close.mean() > 0.999


def test_tie_rules(port):
    """lowest triangle index wins inside a mesh, lowest geom index wins across geoms (src/pathtrace.cu:192)"""
    quad = np.float32([[-1, -1, 0, 1, -1, 0, 1, 1, 0], [-1, -1, 0, 1, 1, 0, -1, 1, 0]])
    tri = np.concatenate([quad, quad])                                          # the same two triangles twice
    T, inv, invT = port.build_transform(np.float32([0, 0, 0]), np.float32([0, 0, 0]), np.float32([1, 1, 1]))
    g = np.zeros(2, orclib.GEOM_DT)
    for k in range(2):
        g[k]["type"], g[k]["transform"], g[k]["inverseTransform"], g[k]["invTranspose"] = 2, T, inv, invT
    o = np.float32([[0.3, -0.2, 2.0], [-0.3, 0.2, -2.0]])
    d = np.float32([[0, 0, -1], [0, 0, 1]])
    hit, t, n, outside = port.closest_hit(g, o, d, (tri, np.int32([[0, 2], [2, 2]])))
    assert list(hit) == [0, 0] and list(outside) == [1, 0]                      # first geom wins; front then back face
    assert np.array_equal(n, np.float32([[0, 0, 1], [0, 0, -1]]))               # the normal faces the incoming ray (-0 == 0)
    assert np.allclose(t, 2.0 - 1e-4, atol=1e-6)                                # getPointOnRay stops 1e-4 short (src/intersections.h:27-29)
