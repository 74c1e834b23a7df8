"""CPU tests of the host-side BVH builder (csrc/ptb_bvh.cpp) through the C ABI (ptb_bvh_build_host; no compute call).

The reference's closest-hit loop is brute force (src/pathtrace.cu:176-199); the BVH only prunes it, so what must hold is
(1) the flat tree is well formed -- every primitive in exactly one leaf, breadth-first links, child boxes enclosing what
hangs below them, depth within the traversal stack -- and (2) the padded boxes are CONSERVATIVE for the reference's own
tests: for every ray the oracle (tests/orclib.py, the CPU restatement of src/intersections.h) reports a hit for, every
box on the way from the root to the leaf of the hit primitive passes the slab test the traversal kernel applies
(ptb_kernels.cuh slabNearPair / bvhNodeStep), so the traversal cannot lose the hit.
"""
import numpy as np
import pytest

import orclib
import scenes
from common import pkg, pkg_tuning, to_oracle_scene

F = np.float32


@pytest.fixture(scope="module")
def port():
    return orclib.load("port")


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    m = pkg()
    d = tmp_path_factory.mktemp("bvh")
    path = str(d / "synthetic.txt")
    with open(path, "w") as f:
        f.write(scenes.synthetic_text(n_objects=3000, res=(400, 225), depth=4, iterations=1))
    return m.Scene.load(path)


def children(nodes, i):
    """(lo, hi, ref, count) of the two children of interior node i"""
    n = nodes[i]
    return ((n["llo"], n["lhi"], int(n["lref"]), int(n["lcount"])), (n["rlo"], n["rhi"], int(n["rref"]), int(n["rcount"])))


def walk(nodes):
    """parent link of every interior node and of every leaf slot: (parentNode, side)"""
    nodeParent = {0: None}
    slotParent = {}
    for i in range(len(nodes)):
        for side, (lo, hi, ref, count) in enumerate(children(nodes, i)):
            if ref < 0:
                continue
            if count > 0:
                for s in range(ref, ref + count):
                    assert s not in slotParent, "slot %d sits in two leaves" % s
                    slotParent[s] = (i, side)
            else:
                assert ref not in nodeParent, "node %d has two parents" % ref
                assert ref > i, "breadth-first order: child %d of node %d" % (ref, i)
                nodeParent[ref] = (i, side)
    return nodeParent, slotParent


@pytest.mark.parametrize("max_leaf", [1, 4])
def test_tree_is_well_formed(synth, max_leaf):
    m = pkg()
    nodes, order, depth = m.bvh_build_host(synth.geoms, max_leaf)
    n = len(synth.geoms)
    assert sorted(order.tolist()) == list(range(n)), "order is a permutation of the geoms"
    nodeParent, slotParent = walk(nodes)
    assert sorted(nodeParent) == list(range(len(nodes))), "every interior node hangs off exactly one parent"
    assert sorted(slotParent) == list(range(n)), "every primitive sits in exactly one leaf"
    assert 0 < depth < 60, "the traversal stack holds 64 entries"
    assert len(nodes) <= max(n, 1)
    for i in range(len(nodes)):
        for lo, hi, ref, count in children(nodes, i):
            if ref < 0:
                continue
            assert np.all(lo <= hi)
            assert count <= 8                                      # 27-bit ref | count << 27 stack entries, leaves of at most 8
            if max_leaf == 1:
                assert count <= 1
            if count == 0:
                # the box of an interior child is the union of the two boxes stored in its own node
                (alo, ahi, _, _), (blo, bhi, bref, _) = children(nodes, ref)
                ulo, uhi = (np.minimum(alo, blo), np.maximum(ahi, bhi)) if bref >= 0 else (alo, ahi)
                assert np.array_equal(lo, ulo) and np.array_equal(hi, uhi)
            else:
                # scene order inside a leaf (keeps the reference's first-index tie rule, src/pathtrace.cu:190)
                assert np.all(np.diff(order[ref:ref + count]) > 0)
    # deterministic: the same scene builds the same bytes
    nodes2, order2, _ = m.bvh_build_host(synth.geoms, max_leaf)
    assert nodes.tobytes() == nodes2.tobytes() and np.array_equal(order, order2)


def test_leaf_boxes_contain_their_primitive(synth):
    m = pkg()
    nodes, order, _ = m.bvh_build_host(synth.geoms, 1)
    _, slotParent = walk(nodes)
    g = synth.geoms
    corners = np.array([[x, y, z, 1.0] for x in (-.5, .5) for y in (-.5, .5) for z in (-.5, .5)])
    rs = np.random.RandomState(3)
    u = rs.normal(size=(64, 3))
    u = 0.5 * u / np.linalg.norm(u, axis=1, keepdims=True)
    surf = np.concatenate([u, np.ones((64, 1))], 1)
    for slot in range(0, len(g), 7):
        i, side = slotParent[slot]
        lo, hi, _, _ = children(nodes, i)[side]
        geom = g[order[slot]]
        T = geom["transform"].reshape(4, 4).T.astype(np.float64)        # column-major mat4
        pts = (corners if geom["type"] == 1 else surf) @ T.T
        assert np.all(pts[:, :3] >= lo) and np.all(pts[:, :3] <= hi)


def slab(lo, hi, o, inv):
    """slabNearPair (ptb_kernels.cuh) for one child box and n rays, in float32"""
    with np.errstate(invalid="ignore", over="ignore"):
        t0 = (lo[None, :] - o) * inv
        t1 = (hi[None, :] - o) * inv
    near = np.fmax(np.fmax(np.fmin(t0[:, 0], t1[:, 0]), np.fmin(t0[:, 1], t1[:, 1])), np.fmax(np.fmin(t0[:, 2], t1[:, 2]), F(0)))
    far = np.fmin(np.fmin(np.fmax(t0[:, 0], t1[:, 0]), np.fmax(t0[:, 1], t1[:, 1])), np.fmax(t0[:, 2], t1[:, 2]))
    return near, near <= far * F(1.00001) + F(1e-5)


def rays_of(port, scene, n_bounce, seed):
    """camera rays plus rays leaving their hit points in random directions (origins ON surfaces, like bounce rays)"""
    os_ = to_oracle_scene(scene)
    o, d = port.generate_rays(os_.camera)
    hit, t, nrm, _ = port.closest_hit(os_.geoms, o, d)
    rs = np.random.RandomState(seed)
    k = np.flatnonzero(hit >= 0)
    k = k[rs.permutation(len(k))[:n_bounce]]
    dn = d[k] / np.linalg.norm(d[k], axis=1, keepdims=True)
    o2 = (o[k] + dn * (t[k, None] - F(1e-4))).astype(F)
    d2 = rs.normal(size=(len(k), 3)).astype(F)
    d2[::40, 0] = 0.0                                              # axis-parallel directions: 1 / 0 in the slab test
    d2[::55, 1] = 0.0
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True).astype(F)
    return np.concatenate([o, o2]), np.concatenate([d, d2]), os_


def pruned_hits(nodes, order, o, d, hit, t):
    """Rays whose oracle hit the traversal would lose: some box between the root and the leaf of the hit primitive fails
    bvhNodeStep's visit rule with best.t = the hit's own distance (best.t is never smaller during a traversal).
    Returns (lost rays, number of (ray, box) pairs checked)."""
    nodeParent, slotParent = walk(nodes)
    slotOf = np.empty(len(order), np.int64)
    slotOf[order] = np.arange(len(order))
    with np.errstate(divide="ignore"):
        inv = (F(1) / d).astype(F)
    dl = (np.sqrt((d * d).sum(1, dtype=F)) * F(0.9999)).astype(F)
    todo = {}                                                      # (node, side) -> rays, over the whole root-to-leaf chain
    for r in np.flatnonzero(hit >= 0):
        link = slotParent[int(slotOf[hit[r]])]
        while link is not None:
            todo.setdefault(link, []).append(r)
            link = nodeParent[link[0]]
    lost, checked = set(), 0
    for (i, side), rr in todo.items():
        rr = np.asarray(rr)
        lo, hi, _, _ = children(nodes, i)[side]
        near, inside = slab(lo, hi, o[rr], inv[rr])
        visit = inside & (near * dl[rr] - F(1e-4) <= t[rr])
        lost.update(rr[~visit].tolist())
        checked += len(rr)
    return sorted(lost), checked


@pytest.mark.parametrize("max_leaf", [1, 4])
def test_boxes_are_conservative_for_the_reference_tests(port, synth, max_leaf):
    m = pkg()
    nodes, order, _ = m.bvh_build_host(synth.geoms, max_leaf)
    o, d, os_ = rays_of(port, synth, 40000, 11)
    hit, t, _, _ = port.closest_hit(os_.geoms, o, d)
    assert (hit >= 0).sum() > 40000
    lost, checked = pruned_hits(nodes, order, o, d, hit, t)
    assert not lost, "the boxes prune the oracle's hit of rays %s" % lost[:8]
    assert checked > 150000
    assert len(stage_b_rejects(m, os_.geoms, o, d, hit, t)) == 0
    assert len(sure_miss_rejects(os_.geoms, o, d, hit)) == 0


def test_far_sphere_hit_the_reference_reports_outside_the_sphere(port, tmp_path, monkeypatch):
    """CPU side of tests/test_gpu_scale.py's regression of the same name, on the full 100 k-object scene of BASELINE.json
    configs[4]: ~240 object units from the origin of a thin ellipsoid the reference's sphere test (src/intersections.h:102-144)
    cancels |ro|^2 - 0.25 against (ro . rd)^2 and reports hits for rays that pass just OUTSIDE sphere 31153.  The shipped leaf
    boxes keep every one of those hits; boxes hugging the ellipsoid (PTB_BVH_TIGHT_SPHERES=1, withdrawn) lose some --
    which also shows that this check can fail."""
    m = pkg()
    path = str(tmp_path / "synthetic.txt")
    with open(path, "w") as f:
        f.write(scenes.synthetic_text())
    s = m.Scene.load(path)
    assert len(s.geoms) == 100000
    o = np.array([[26.248161, 22.502022, -8.040187]], F)
    d = np.array([[0.16347983, 0.02360658, -0.9862642]], F)
    rs = np.random.RandomState(14)
    o = np.concatenate([o, o + rs.uniform(-2e-5, 2e-5, (4095, 3)).astype(F)])
    d = np.concatenate([d, d + rs.uniform(-2e-7, 2e-7, (4095, 3)).astype(F)])
    hit, t, _, _ = port.closest_hit(to_oracle_scene(s).geoms, o, d)
    assert hit[0] == 31153 and (hit == 31153).sum() > 100 and (hit != 31153).sum() > 100      # the grazing band
    monkeypatch.delenv("PTB_BVH_TIGHT_SPHERES", raising=False)
    nodes, order, depth = m.bvh_build_host(s.geoms, 1)
    assert depth < 60
    lost, checked = pruned_hits(nodes, order, o, d, hit, t)
    assert not lost and checked > 4096 * 10
    monkeypatch.setenv("PTB_BVH_TIGHT_SPHERES", "1")            # read by the TUNING build only
    assert m.bvh_build_host(s.geoms, 1)[0].tobytes() == nodes.tobytes()     # the release library ignores the environment
    tight, order2, _ = pkg_tuning().bvh_build_host(s.geoms, 1)
    lost, _ = pruned_hits(tight, order2, o, d, hit, t)
    assert lost and np.all(hit[lost] == 31153)


def random_spheres(orc, rs, n, scale_lo, scale_hi, kind=0):
    """SPHERE (kind 0) or CUBE (kind 1) geoms placed like the synthetic scene's (tests/scenes.py synthetic_text), per-axis scales"""
    g = np.zeros(n, orclib.GEOM_DT)
    g["type"] = kind
    g["materialid"] = 1
    t = np.stack([rs.uniform(-50, 50, n), rs.uniform(0, 100, n), rs.uniform(-100, 0, n)], 1).astype(F)
    r = rs.uniform(0, 360, (n, 3)).astype(F)
    s = rs.uniform(scale_lo, scale_hi, (n, 3)).astype(F)
    for i in range(n):
        g["translation"][i], g["rotation"][i], g["scale"][i] = t[i], r[i], s[i]
        g["transform"][i], g["inverseTransform"][i], g["invTranspose"][i] = orc.build_transform(t[i], r[i], s[i])
    return g


def grazing_band(orc, g, rs, origins, band, dist, scale_hi, in_scene=False):
    """Rays aimed at the silhouette of every sphere of g from `origins` far points each: the edge is found by bisection on the
    oracle's own answer, then `band` rays sweep from 2e-3 world units inside it to 2e-2 outside (log-spaced: the spurious hits of
    the reference's cancellation-prone sphere test live just outside).  Returns (o, d, owner sphere of every ray)."""
    K, Q = len(g), origins
    c = np.repeat(g["translation"].astype(np.float64), Q, 0)
    owner = np.repeat(np.arange(K), Q)
    u = rs.normal(size=(K * Q, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    org = c + u * rs.uniform(dist[0], dist[1], (K * Q, 1))
    if in_scene:                                                   # origins next to OTHER spheres, like bounce rays
        org = g["translation"].astype(np.float64)[rs.randint(0, K, K * Q)] + 2.0 * scale_hi * u
        u = org - c; u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = np.cross(u, rs.normal(size=(K * Q, 3))); w /= np.linalg.norm(w, axis=1, keepdims=True)

    def rays(c, org, w, sv):                                       # target = centre + w * s
        d = c + w * sv[:, None] - org
        return org.astype(F), (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(F)

    lo = np.zeros(K * Q); hi = np.full(K * Q, 2.0 * scale_hi)
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        inside = orc.closest_hit(g, *rays(c, org, w, mid))[0] == owner
        lo = np.where(inside, mid, lo); hi = np.where(inside, hi, mid)
    offs = np.concatenate([-np.geomspace(2e-3, 1e-7, band // 4), np.geomspace(1e-7, 2e-2, band - band // 4)])
    sv = (0.5 * (lo + hi)[:, None] + offs[None, :]).ravel()
    rep = lambda x: np.repeat(x, len(offs), 0)
    o, d = rays(rep(c), rep(org), rep(w), sv)
    return o, d, rep(owner)


def stage_b_rejects(m, g, o, d, hit, t):
    """Oracle hits that the per-primitive pruning of the closest-hit searches would throw away: boundEntryRaw (ptb_math.cuh) of the
    HIT primitive restated in float32 -- padded world box for cubes, padded bounding sphere with the errK slack for spheres --
    followed by the searches' acceptance rule `e >= 0 && e * 0.99998 - slack <= best.t` (ptb_kernels.cuh stage B, best.t >= t)."""
    lo, hi, c, r2, ek = m.geom_bounds_host(g.view(m.GEOM_DT))
    k = np.flatnonzero(hit >= 0)
    j = hit[k]
    ok_, dk, tk = o[k], d[k], t[k]
    rd = np.sqrt((dk * dk).sum(1, dtype=F)).astype(F)
    dn = (dk / rd[:, None]).astype(F)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = (F(1) / dk).astype(F)
        t0 = (lo[j] - ok_) * inv
        t1 = (hi[j] - ok_) * inv
    tn = np.fmax(np.fmax(np.fmin(t0[:, 0], t1[:, 0]), np.fmin(t0[:, 1], t1[:, 1])), np.fmax(np.fmin(t0[:, 2], t1[:, 2]), F(0)))
    tf = np.fmin(np.fmin(np.fmax(t0[:, 0], t1[:, 0]), np.fmax(t0[:, 1], t1[:, 1])), np.fmax(t0[:, 2], t1[:, 2]))
    e_cube = np.where(tn <= tf * F(1.00002) + F(1e-6), tn * rd, F(-1)).astype(F)
    oc = (c[j] - ok_).astype(F)
    oc2 = (oc * oc).sum(1, dtype=F)
    bq = (oc * dn).sum(1, dtype=F)
    det = (bq * bq + (r2[j] - oc2) + ek[j] * (oc2 + bq * bq)).astype(F)
    sq = np.sqrt(np.maximum(det, F(0))).astype(F)
    e_sph = np.where((det < 0) | (bq + sq * F(1.000001) < 0), F(-1), np.fmax(bq - sq * F(1.000001) - F(1e-6) * np.abs(bq), F(0))).astype(F)
    e = np.where(g["type"][j] == 1, e_cube, e_sph)
    slack = F(2e-5) * (np.abs(ok_).sum(1, dtype=F) + F(1))
    keep = (e >= 0) & (e * F(0.99998) - slack <= tk)
    return k[~keep]


def sure_miss_rejects(g, o, d, hit):
    """Oracle hits the object-space sure-miss filters (ptb_math.cuh boxSureMiss / sphereSureMiss, run on the hit primitive before the
    reference-exact test in every closest-hit search) would call a miss.  float32 restatement; the filters' margins (4e-6 / 2e-5
    relative) are far above the ulp-level differences between this and the device's approximate reciprocals."""
    k = np.flatnonzero(hit >= 0)
    j = hit[k]
    M = g["inverseTransform"].reshape(-1, 4, 4).transpose(0, 2, 1)[j].astype(F)         # row-major 4x4 of every hit primitive
    ok_, dk = o[k], d[k]
    row = lambda r, v, w: ((M[:, r, 0] * v[:, 0] + M[:, r, 1] * v[:, 1]) + (M[:, r, 2] * v[:, 2] + M[:, r, 3] * F(w))).astype(F)
    qo = np.stack([row(r, ok_, 1) for r in range(3)], 1)
    qu = np.stack([row(r, dk, 0) for r in range(3)], 1)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        # boxSureMiss
        usable = (np.abs(qu) > F(1e-30)).all(1)
        inv = (F(1) / qu).astype(F)
        a1 = (F(-0.5) - qo) * inv
        a2 = (F(0.5) - qo) * inv
        ta, tb = np.fmin(a1, a2), np.fmax(a1, a2)
        tmin = np.where(ta > 0, ta, F(-1e38)).max(1)
        tmax = np.fmin(tb.min(1), F(1e38))
        mag = np.abs(tmax) + np.abs(tmin)
        box_miss = usable & ((tmax - tmin < F(-4e-6) * mag) | (tmax < 0))
        # sphereSureMiss
        uu = (qu * qu).sum(1, dtype=F)
        ok2 = (uu > F(1e-30)) & (uu < F(1e30))
        b = ((qo * qu).sum(1, dtype=F) / np.sqrt(uu)).astype(F)
        oo = (qo * qo).sum(1, dtype=F)
        cc = oo - F(0.25)
        mg = b * b + (oo + F(0.25))
        sph_miss = ok2 & ((b * b - cc < F(-2e-5) * mg) | ((b > 0) & (b * b > F(1e-6) * mg) & (cc > F(2e-5) * mg)))
    return k[np.where(g["type"][j] == 1, box_miss, sph_miss)]


def leaf_boxes(nodes, order):
    """(lo, hi) of the leaf box of every geom, indexed by geom (one primitive per leaf)"""
    _, slotParent = walk(nodes)
    lo = np.zeros((len(order), 3), F); hi = np.zeros((len(order), 3), F)
    for slot, (i, side) in slotParent.items():
        lo[order[slot]], hi[order[slot]] = children(nodes, i)[side][:2]
    return lo, hi


def test_far_grazing_hits_need_the_rigorous_sphere_boxes(port, monkeypatch):
    """From 160-230 world units away -- distances the 100 k-object scene of BASELINE.json configs[4] has room for -- the reference's
    sphere test reports hits up to ~0.1 object units OUTSIDE a thin ellipsoid.  The boxes ptb_create builds (far pad over the
    reach of every sphere inside the box of everything a ray can start from) keep every one of them; the round-1 boxes
    (PTB_BVH_FAR_PAD=0) lose a few."""
    m = pkg()
    rs = np.random.RandomState(1)
    g = random_spheres(port, rs, 300, 0.3, 1.2)
    o, d, owner = grazing_band(port, g, rs, 8, 512, (160.0, 230.0), 1.2)
    # two anchor cubes put every ray origin inside the scene's box, as the origins of real paths are
    anchors = random_spheres(port, rs, 2, 1.0, 1.0, kind=1)
    for k, t in enumerate(([-290.0, -240.0, -340.0], [290.0, 340.0, 240.0])):
        anchors["translation"][k] = t
        anchors["transform"][k], anchors["inverseTransform"][k], anchors["invTranspose"][k] = \
            port.build_transform(np.float32(t), anchors["rotation"][k], anchors["scale"][k])
    g = np.concatenate([g, anchors])
    assert np.all(o >= anchors["translation"][0]) and np.all(o <= anchors["translation"][1])
    hit, t, _, _ = port.closest_hit(g, o, d)
    assert (hit == owner).sum() > 300000
    gm = g.view(m.GEOM_DT)
    monkeypatch.delenv("PTB_BVH_FAR_PAD", raising=False)
    nodes, order, _ = m.bvh_build_host(gm, 1)                       # the library's own choice: what ptb_create builds
    lost, checked = pruned_hits(nodes, order, o, d, hit, t)
    assert not lost and checked > 3000000
    assert len(stage_b_rejects(m, g, o, d, hit, t)) == 0            # the per-primitive bounds carry their own slack (errK)
    assert len(sure_miss_rejects(g, o, d, hit)) == 0                # and so do the object-space sure-miss filters
    nodes0, order0, _ = m.bvh_build_host(gm, 1, 0.0)                # the round-1 boxes: the failing control
    lost0, _ = pruned_hits(nodes0, order0, o, d, hit, t)
    assert lost0 and np.all(hit[lost0] == owner[lost0])
    # the rigorous boxes contain the round-1 boxes (so every hit those found is still found) and are no larger than the
    # uniform worst case over the diameter of the whole scene
    lo, hi = leaf_boxes(nodes, order)
    lo0, hi0 = leaf_boxes(nodes0, order0)
    assert np.all(lo <= lo0) and np.all(hi >= hi0) and np.any(lo < lo0)
    span = anchors["translation"][1].astype(np.float64) - anchors["translation"][0] + 4.0
    loU, hiU = leaf_boxes(*m.bvh_build_host(gm, 1, 1.05 * float(np.linalg.norm(span)))[:2])
    assert np.all(lo >= loU) and np.all(hi <= hiU) and np.any(lo > loU)
    monkeypatch.setenv("PTB_BVH_FAR_PAD", "0")                  # the knob exists in the TUNING build only
    assert pkg_tuning().bvh_build_host(gm, 1)[0].tobytes() == nodes0.tobytes()
    assert m.bvh_build_host(gm, 1)[0].tobytes() == nodes.tobytes()


def test_sphere_pruning_bound_carries_the_anisotropy_slack(port, monkeypatch):
    """The per-primitive bounding-sphere test (stage B of the BVH search, phase 1 of the brute-force one) must keep the hits the
    reference reports for rays passing outside a thin ellipsoid: the slack of its discriminant grows with (sigma_max / sigma_min)^2
    (ptb_bvh.cpp sphereBounds).  With the 2e-6 constant that only covers near-isotropic spheres the same rays are rejected."""
    m = pkg()
    rs = np.random.RandomState(5)
    g = random_spheres(port, rs, 150, 0.05, 2.0)
    o, d, owner = grazing_band(port, g, rs, 4, 256, (100.0, 250.0), 2.0)
    hit, t, _, _ = port.closest_hit(g, o, d)
    assert (hit == owner).sum() > 30000
    assert len(stage_b_rejects(m, g, o, d, hit, t)) == 0
    assert len(sure_miss_rejects(g, o, d, hit)) == 0
    shipped = m.geom_bounds_host

    def isotropic_only(geoms, reach=0.0):
        lo, hi, c, r2, ek = shipped(geoms, reach)
        assert ek.max() > 1e-4                                      # thin ellipsoids ask for far more than the constant
        return lo, hi, c, r2, np.full_like(ek, 2e-6)
    monkeypatch.setattr(m, "geom_bounds_host", isotropic_only)
    assert len(stage_b_rejects(m, g, o, d, hit, t)) > 100


def test_sure_miss_filters_prove_misses_and_keep_hits(port):
    """The restated object-space filters are not vacuous: they prove nearly every miss of random (ray, primitive) pairs and never
    call a primitive the oracle hits a miss."""
    rs = np.random.RandomState(5)
    n = 100000
    o = np.stack([rs.uniform(-50, 50, n), rs.uniform(0, 100, n), rs.uniform(-100, 0, n)], 1).astype(F)
    d = rs.normal(size=(n, 3)).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for kind in (0, 1):
        g = random_spheres(port, rs, 100, 0.3, 1.2, kind)
        hit = port.closest_hit(g, o, d)[0]
        other = rs.randint(0, 100, n).astype(np.int32)
        other[hit == other] = -1
        assert len(sure_miss_rejects(g, o, d, other)) > 0.99 * (other >= 0).sum()
        assert (hit >= 0).sum() > 100 and len(sure_miss_rejects(g, o, d, hit)) == 0


def test_degenerate_scenes():
    m = pkg()
    nodes, order, depth = m.bvh_build_host(np.zeros(0, m.GEOM_DT))
    assert len(nodes) == 1 and len(order) == 0 and nodes[0]["lref"] == -1 and nodes[0]["rref"] == -1
    g = np.zeros(1, m.GEOM_DT)
    g["type"] = 1
    g["transform"][0] = np.eye(4, dtype=F).ravel()
    g["inverseTransform"][0] = np.eye(4, dtype=F).ravel()
    nodes, order, depth = m.bvh_build_host(g)
    assert len(nodes) == 1 and order.tolist() == [0]
    assert nodes[0]["lref"] == 0 and nodes[0]["lcount"] == 1 and nodes[0]["rref"] == -1
    assert np.all(nodes[0]["llo"] < -0.5) and np.all(nodes[0]["lhi"] > 0.5) and np.all(nodes[0]["lhi"] < 0.51)
    # coincident primitives cannot be split by any plane: the builder still terminates with one primitive per leaf slot
    g = np.repeat(g, 40)
    nodes, order, depth = m.bvh_build_host(g)
    _, slotParent = walk(nodes)
    assert sorted(slotParent) == list(range(40)) and depth < 60
    # a singular transform (SCALE 0 on one axis: the inverse is not finite) still gets finite boxes
    port = orclib.load("port")
    g2 = random_spheres(port, np.random.RandomState(2), 20, 0.3, 1.2)
    g2["scale"][5] = [1.0, 0.0, 1.0]
    with np.errstate(all="ignore"):
        g2["transform"][5], g2["inverseTransform"][5], g2["invTranspose"][5] = port.build_transform(g2["translation"][5], g2["rotation"][5], g2["scale"][5])
    assert not np.isfinite(g2["inverseTransform"][5]).all()
    nodes, order, depth = m.bvh_build_host(g2.view(m.GEOM_DT))
    _, slotParent = walk(nodes)
    assert sorted(slotParent) == list(range(20))
    for name in ("llo", "lhi", "rlo", "rhi"):
        assert np.isfinite(nodes[name]).all()
    # a mesh geom (type 2) has no box of its own: triangles are boxed one by one inside ptb_create
    g["type"][3] = 2
    with pytest.raises(m.PtbError):
        m.bvh_build_host(g)
