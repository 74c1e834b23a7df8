"""The argument behind PtbOptions.leaveTest, checked on the CPU against the oracle (no GPU).

closestHitBrute<LEAVE> (csrc/ptb_kernels.cuh, leaveSkip) drops the cube a bounce ray starts on when ONE row of the object-space ray says

    s * qo_k > 0.5   and   s * qu_k > 0        (k, s = axis and side of the face the ray left)

with qo_k, qu_k computed in the order multiplyMV computes them (src/intersections.h:34-36, glm's mat4 * vec4).  The claim: the reference's
boxIntersectionTest (src/intersections.h:48-90) then reports a miss for that cube -- for every finite ray, whatever the other two axes
do, also for tilted, anisotropically scaled cubes.  Here: rays that start a hair off the faces of random such cubes (the way scatterRay
leaves them) and rays from anywhere, the row test restated in float32 numpy, the oracle's answer for the cube alone."""
import numpy as np
import pytest

import orclib
from test_bvh_host import random_spheres

F = np.float32


def row_test(g, o, d, axis, side):
    """float32 restatement of leaveSkip's row test for one cube: rays o, d (n x 3), face (axis, side)"""
    m = g["inverseTransform"].reshape(4, 4)                      # column-major glm mat4: m[j][i] = element (row i, column j)
    rx, ry, rz, rw = (F(m[0][axis]), F(m[1][axis]), F(m[2][axis]), F(m[3][axis]))
    ox, oy, oz = o[:, 0], o[:, 1], o[:, 2]
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    qo = (rx * ox + ry * oy) + (rz * oz + rw)                   # mvRowPoint: (m0 v0 + m1 v1) + (m2 v2 + m3)
    qu = (rx * dx + ry * dy) + (rz * dz + rw * F(0.0))         # mvRow with w = 0
    s = F(1.0) if side > 0 else F(-1.0)
    finite = (np.abs(o).sum(1) + 1.0) * 2e-5 < 20.0
    finite &= (d.astype(np.float64) ** 2).sum(1) < 1e12
    return finite & (s * qo > F(0.5)) & (s * qu > F(0.0))


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_row_test_implies_the_reference_misses_the_cube(kind):
    orc = orclib.load(kind)
    if orc is None:
        pytest.skip("oracle build %r not present" % kind)
    rs = np.random.RandomState(2024)
    cubes = random_spheres(orc, rs, 48, 0.05, 3.0, kind=1)       # tilted, per-axis scales 0.05 .. 3: strongly anisotropic
    n = 20000
    settled = 0
    for i in range(len(cubes)):
        g = cubes[i:i + 1]
        fwd = g["transform"][0].reshape(4, 4).T.astype(np.float64)     # row-major 4 x 4
        for axis in range(3):
            for side in (+1, -1):
                # points on the face, pushed off it along the WORLD normal the reference builds (transform * face normal, :86) by the
                # reference's offsets and by nothing at all; directions over the whole sphere (most leave, some re-enter)
                q = rs.uniform(-0.5, 0.5, (n, 3)); q[:, axis] = 0.5 * side
                p = (fwd[:3, :3] @ q.T).T + fwd[:3, 3]
                nrm = fwd[:3, axis] * side; nrm /= np.linalg.norm(nrm)
                off = rs.choice([0.0, 1e-6, 1e-4, 1e-3, 1e-2], (n, 1))
                o = (p + nrm * off).astype(F)
                d = rs.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
                d = d.astype(F)
                # and a share of rays from far away, unnormalised directions included
                far = rs.rand(n) < 0.2
                o[far] = (rs.normal(size=(int(far.sum()), 3)) * 200.0).astype(F)
                d[far] = (d[far] * rs.uniform(0.1, 10.0, (int(far.sum()), 1))).astype(F)
                # degenerate directions: components that are exactly +0 / -0 (the slab divisions by zero of :60-62)
                z = rs.rand(n, 3) < 0.05
                d[z] = np.where(rs.rand(int(z.sum())) < 0.5, F(0.0), F(-0.0))
                cond = row_test(g[0], o, d, axis, side)
                hit, t, _, _ = orc.closest_hit(g, o, d)
                bad = cond & (hit >= 0)
                assert not bad.any(), (i, axis, side, int(bad.sum()), o[bad][0].tolist(), d[bad][0].tolist(), float(t[bad][0]))
                settled += int(cond.sum())
    assert settled > 0.3 * len(cubes) * 6 * n                     # the test is not vacuous: the row test settles a large share of the rays
