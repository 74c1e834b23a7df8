"""CPU tests of the product's host side (no compute calls): the C-ABI library loads and exports every
symbol include/ptb200.h declares, the C++ scene loader / camera fix-up / PNG writer behave like the
reference's, and rendering fails loudly without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orclib
import scenes
from common import pkg, golden, has_gpu, ROOT


def test_library_exports_every_declared_symbol():
    m = pkg()
    L = m.lib()
    header = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(m.ABI_SYMBOLS), declared ^ set(m.ABI_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(L, name), "libptb200.so does not export %s" % name
    L.ptb_abi_version.restype = C.c_int
    assert L.ptb_abi_version() == 7      # PTB_ABI_VERSION: 2 added antialias / lens options, 3 triangle meshes in PtbSceneDesc, 4 deposit mode / stage timers / lazy image, 5 regroupMaterials, 6 leaveTest, 7 hitQueue


def test_struct_sizes_match_reference_layout():
    m = pkg()
    assert m.GEOM_DT.itemsize == 236 and m.MATERIAL_DT.itemsize == 44 and m.CAMERA_DT.itemsize == 84
    assert C.sizeof(m.PtbCamera) == 84
    o = m.PtbOptions()
    m.lib().ptb_default_options(C.byref(o))
    assert o.device == -1 and o.bsdfMode == m.BSDF_FULL and o.pipeline == m.PIPELINE_FUSED and o.useGraphs == 1


def test_options_mirror_matches_the_header():
    """The ctypes mirror of PtbOptions has the header's fields, in the header's order (ptb_default_options writes the whole C struct), and
    the defaults of the image-neutral options are the documented ones."""
    m = pkg()
    header = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    body = re.search(r"typedef struct PtbOptions\s*\{(.*?)\}\s*PtbOptions;", header, flags=re.S)
    if body is None:
        body = re.search(r"typedef struct\s*\{([^{}]*?)\}\s*PtbOptions;", header, flags=re.S)
    assert body, "PtbOptions not found in include/ptb200.h"
    declared = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*;", body.group(1))
    assert declared == [n for n, _ in m.PtbOptions._fields_], (declared, [n for n, _ in m.PtbOptions._fields_])
    o = m.PtbOptions()
    m.lib().ptb_default_options(C.byref(o))
    assert (o.leaveTest, o.hitQueue, o.lookahead, o.depositMode) == (1, -1, 1, m.DEPOSIT_ORDERED)


def test_scene_loader_matches_reference_pod(scene_dir):
    m = pkg()
    g = golden()
    for name in ("cornell", "sphere"):
        raw = m.Scene.load("%s/%s.txt" % (scene_dir, name), fixup=False)
        fixed = m.Scene.load("%s/%s.txt" % (scene_dir, name), fixup=True)
        assert raw.geoms.tobytes() == g["scene_%s_geoms" % name].tobytes()
        assert raw.materials.tobytes() == g["scene_%s_materials" % name].tobytes()
        assert raw.camera.tobytes() == g["scene_%s_camera_raw" % name].tobytes()
        assert fixed.camera.tobytes() == g["scene_%s_camera_fixed" % name].tobytes()
        assert [raw.iterations, raw.traceDepth] == list(g["scene_%s_meta" % name])
        assert raw.imageName == name


def test_scene_loader_matches_oracle_on_generated_scenes(tmp_path):
    m = pkg()
    port = orclib.load("port")
    p = tmp_path / "syn.txt"
    p.write_text(scenes.synthetic_text(n_objects=2000, seed=11))
    a = m.Scene.load(str(p))
    b = port.load_scene(str(p))
    assert len(a.geoms) == 2000
    assert a.geoms.tobytes() == b.geoms.tobytes() and a.materials.tobytes() == b.materials.tobytes()
    assert a.camera.tobytes() == b.camera.tobytes()


def test_loader_quirks(tmp_path):
    m = pkg()
    base = scenes.cornell_text()
    (tmp_path / "crlf.txt").write_text(base.replace("\n", "\r\n"))
    (tmp_path / "cr.txt").write_text(base.replace("\n", "\r"))
    (tmp_path / "notrail.txt").write_text(base.rstrip("\n"))
    ref = m.Scene.load(str(_write(tmp_path, "lf.txt", base)))
    for f in ("crlf.txt", "cr.txt", "notrail.txt"):
        s = m.Scene.load(str(tmp_path / f))
        assert s.geoms.tobytes() == ref.geoms.tobytes() and s.camera.tobytes() == ref.camera.tobytes()
    # a trailing blank after the type keyword defeats the whole-line compare (src/scene.cpp:48,51):
    # the object keeps the default type, as in the reference
    odd = base.replace("\nsphere\n", "\nsphere \n")
    s = m.Scene.load(str(_write(tmp_path, "odd.txt", odd)))
    assert len(s.geoms) == 7
    # out-of-sequence ids are skipped (src/scene.cpp:37-39)
    skipped = base.replace("OBJECT 6", "OBJECT 9")
    s = m.Scene.load(str(_write(tmp_path, "skip.txt", skipped)))
    assert len(s.geoms) == 6
    with pytest.raises(m.PtbError):
        m.Scene.load(str(tmp_path / "missing.txt"))


def _write(d, name, text):
    p = d / name
    p.write_text(text)
    return p


def test_png_writer_follows_image_cpp(tmp_path):
    from PIL import Image
    m = pkg()
    w, h = 13, 7
    rs = np.random.RandomState(3)
    summed = rs.uniform(-0.5, 3.0, (w * h, 3)).astype(np.float32) * 4
    m.save_png(str(tmp_path / "out"), w, h, summed, 4, flip_x=True)
    got = np.asarray(Image.open(str(tmp_path / "out.png")).convert("RGB"))
    assert got.shape == (h, w, 3)
    # src/main.cpp:83-89 (x flip, / samples) then src/image.cpp:27-30 (clamp, *255, truncate)
    pix = (summed / np.float32(4)).reshape(h, w, 3)[:, ::-1, :]
    want = (np.clip(pix, 0, 1) * np.float32(255)).astype(np.uint8)
    assert np.array_equal(got, want)
    m.save_png(str(tmp_path / "noflip"), w, h, summed, 4, flip_x=False)
    got2 = np.asarray(Image.open(str(tmp_path / "noflip.png")).convert("RGB"))
    assert np.array_equal(got2, want[:, ::-1, :])


def test_hdr_writer(tmp_path):
    m = pkg()
    img = np.linspace(0, 6, 5 * 4 * 3, dtype=np.float32).reshape(20, 3)
    rc = m.lib().ptb_save_hdr(str(tmp_path / "o").encode(), 5, 4, img.ctypes.data, C.c_float(1.0), 0)
    assert rc == 0
    data = open(str(tmp_path / "o.hdr"), "rb").read()
    assert data.startswith(b"#?RADIANCE") and data.endswith(bytes(0)) and len(data) > 80


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a machine without a GPU")
def test_render_fails_loudly_without_gpu(scene_dir):
    m = pkg()
    s = m.Scene.load(scene_dir + "/cornell.txt")
    with pytest.raises(m.PtbError) as e:
        m.Renderer(s)
    assert "CUDA" in str(e.value)
    # the reference-named entry points behave the same
    with pytest.raises(m.PtbError):
        m.pathtraceInit(s)
    m.pathtraceFree()          # free before/without init is a no-op (src/main.cpp:126)
    with pytest.raises(m.PtbError):
        m.pathtrace(None, 0, 1)


def test_product_does_not_touch_the_oracle():
    """No product source may reference oracle/ (parity would be void)."""
    pkgdir = os.path.join(ROOT, "project3-cuda-path-tracer_b200")
    for base, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "libptoracle" not in text and "libptref" not in text and "orc_" not in text, os.path.join(base, f)
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkgdir, "libptb200.so")], capture_output=True, text=True).stdout
    assert "ptoracle" not in out and "ptref" not in out


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="/root/reference not mounted")
def test_dropin_shim_compiles_and_links_against_the_reference_headers(tmp_path):
    """csrc/host/pathtrace.cpp touches only Scene's public members, so a maintainer can drop it in
    place of src/pathtrace.cu: compile it against the REFERENCE's scene.h / sceneStructs.h / glm and
    link it with the reference's scene.cpp + utilities.cpp and libptb200.so."""
    import shutil
    import subprocess
    ref = "/root/reference"
    pkgdir = os.path.join(ROOT, "project3-cuda-path-tracer_b200")
    shutil.copy(os.path.join(pkgdir, "csrc", "host", "pathtrace.cpp"), str(tmp_path / "pathtrace.cpp"))
    (tmp_path / "app.cpp").write_text('''
#include "scene.h"
#include "pathtrace.h"
Scene::~Scene() {}
int main(int argc, char **argv) {
    Scene *scene = new Scene(argv[1]);
    pathtraceFree();                 // before the first init, like src/main.cpp:126
    pathtraceInit(scene);            // exits with "CUDA error" when there is no device
    pathtrace(NULL, 0, 1);
    pathtraceFree();
    return 0;
}
''')
    inc = ["-I%s/src" % ref, "-I%s/external/include" % ref, "-I/usr/local/cuda/include", "-I%s/include" % ROOT]
    exe = str(tmp_path / "app")
    cmd = ["g++", "-std=c++14", "-O1", "-w", "-DPTB_USE_REAL_GLM"] + inc + [
        str(tmp_path / "pathtrace.cpp"), str(tmp_path / "app.cpp"), ref + "/src/scene.cpp", ref + "/src/utilities.cpp",
        "-L" + pkgdir, "-lptb200", "-Wl,-rpath," + pkgdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    scene = tmp_path / "cornell.txt"
    scene.write_text(scenes.cornell_text(res=(32, 32)))
    run = subprocess.run([exe, str(scene)], capture_output=True, text=True)
    if has_gpu():
        assert run.returncode == 0, run.stderr
    else:
        assert run.returncode == 1 and "CUDA error" in run.stderr      # checkCUDAError semantics: print + exit(EXIT_FAILURE)


def test_headless_driver_usage_and_errors(tmp_path):
    import subprocess
    exe = os.path.join(ROOT, "project3-cuda-path-tracer_b200", "ptb200_headless")
    assert os.path.exists(exe), "build() must produce the headless driver"
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stdout                    # src/main.cpp:36-39
    r = subprocess.run([exe, str(tmp_path / "nope.txt"), "--quiet"], capture_output=True, text=True)
    assert r.returncode == 1 and "Error reading from file" in r.stderr  # src/scene.cpp:12-15


def test_reference_gl_shell_links_over_the_library(tmp_path):
    """the reference's own display shell -- src/main.cpp, preview.cpp, glslUtility.cpp, image.cpp, scene.cpp,
    utilities.cpp, stb.cpp, UNMODIFIED and compiled where they lie (oracle/Makefile `glshell`) -- links against the product's
    pathtraceInit / pathtraceFree / pathtrace, with oracle/gl_stub.cpp standing in for the GL stack this image does not have.
    Here (no GPU): the binary exists, takes its three renderer symbols from the shim over libptb200.so, prints the app's usage
    line, and dies with checkCUDAError's message at pathtraceInit.  tests/test_gpu_dropin.py runs it for real."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gl_shell")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_gl_shell is built where /root/reference is mounted")
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for need in ("ptb_create", "ptb_destroy", "ptb_pathtrace"):
        assert need in syms, need
    assert "libptb200.so" in subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stdout + r.stderr              # src/main.cpp:36-39
    scene = tmp_path / "cornell.txt"
    scene.write_text(scenes.cornell_text(res=(32, 32), iterations=2))
    run = subprocess.run([exe, str(scene)], capture_output=True, text=True, cwd=str(tmp_path))
    if not has_gpu():
        assert run.returncode == 1 and "CUDA error" in run.stderr
