"""The reference's own application over the library: oracle/_ref/ref_gl_shell is the reference's
UNMODIFIED src/main.cpp + preview.cpp + glslUtility.cpp + image.cpp + scene.cpp + utilities.cpp + stb.cpp, compiled where they
lie and linked against pathtraceInit / pathtraceFree / pathtrace of csrc/host/pathtrace.cpp over libptb200.so, with a stub for the
GL stack this image does not have (oracle/gl_stub.cpp: the mapped PBO is plain device memory).  It runs its GLFW main loop,
calls pathtrace(pbo, 0, iteration) every frame, saves <FILE>.<time>.<N>samp.png through its own stb writer after ITERATIONS
frames and exits (src/main.cpp:125-147)."""
import glob
import os
import subprocess

import numpy as np
import pytest

import scenes
from common import pkg, ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_gl_shell")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/ref_gl_shell is built where /root/reference is mounted (it travels with the snapshot)")
def test_unmodified_reference_app_renders_through_the_library(tmp_path):
    from PIL import Image
    m = pkg()
    n = 5
    scene = tmp_path / "cornell.txt"
    scene.write_text(scenes.cornell_text(res=(96, 64), iterations=n))
    run = subprocess.run([EXE, str(scene)], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert run.returncode == 0, run.stderr[-2000:]
    pngs = glob.glob(str(tmp_path / ("cornell.*.%dsamp.png" % n)))              # src/main.cpp:91-94 file name
    assert len(pngs) == 1, os.listdir(str(tmp_path))
    got = np.asarray(Image.open(pngs[0]).convert("RGB"))
    # the same frames through the library's own entry points: pathtrace() semantic, then saveImage's arithmetic on the device
    s = m.Scene.load(str(scene))
    with m.Renderer(s, device=0, batch_iters=1) as r:
        for it in range(1, n + 1):
            r.pathtrace(None, 0, it, s.image)
        want = r.resolve_rgb8(n)
    assert got.shape == want.shape == (64, 96, 3)
    assert np.array_equal(got, want)
    assert got.max() > 100                                                       # the light is in the picture
