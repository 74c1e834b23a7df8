"""Golden fixtures for the camera / material extensions (stochastic antialiasing,
thin-lens depth of field, imperfect specular), generated from the REFERENCE-LINKED build
(oracle/_ref/libptref.so, where the formulas are written with the reference's own glm / thrust calls and
its calculateRandomDirectionInHemisphere frame).  The reference leaves these features as TODOs
(src/pathtrace.cu:118-120,134; src/interactions.h:44-79), so the semantics are the ones documented in
DESIGN.md section 8; what these vectors pin is the arithmetic.  Run in the build container only:

    python tests/golden/make_golden_ext.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orclib   # noqa: E402
import scenes   # noqa: E402

RES = (96, 64)
LENS_CASES = {"aa": dict(antialias=1), "dof": dict(lens_radius=0.35, focal_distance=9.5),
              "aadof": dict(antialias=1, lens_radius=0.2, focal_distance=11.0)}


def main():
    ref = orclib.load("reference")
    assert ref is not None and ref.kind == "reference", "build oracle/_ref first (make -C oracle ref)"
    import tempfile
    d = tempfile.mkdtemp()
    out = {}
    texts = {"cornell": scenes.cornell_text(res=RES), "glossy": scenes.glossy_text(res=RES)}
    sc = {}
    for name, text in texts.items():
        p = os.path.join(d, name + ".txt")
        open(p, "w").write(text)
        sc[name] = ref.load_scene(p)
    for tag, kw in LENS_CASES.items():
        o, dd = ref.generate_rays_lens(sc["cornell"].camera, 3, **kw)
        out["rays_%s_o" % tag], out["rays_%s_d" % tag] = o, dd
        img, alive, seg = ref.render(sc["cornell"], orclib.MODE_FULL, 1, 2, **kw)
        out["img_%s" % tag], out["alive_%s" % tag] = img, alive
    img, alive, seg = ref.render(sc["glossy"], orclib.MODE_FULL, 1, 2)
    out["img_glossy"], out["alive_glossy"] = img, alive
    # ---- triangle meshes: glm::intersectRayTriangle of the vendored glm, called by the
    # reference-linked build; the scene (OBJ files included) is parsed by the product's host loader
    from common import pkg, to_oracle_scene
    ms = to_oracle_scene(pkg().Scene.load(scenes.mesh_scene_files(os.path.join(d, "mesh"), res=RES)))
    o, dd = ref.generate_rays(ms.camera)
    hit, t, n, outside = ref.closest_hit(ms.geoms, o, dd, ms)
    out["mesh_triangles"], out["mesh_ranges"] = ms.triangles, ms.mesh_ranges
    out["mesh_hit"], out["mesh_t"], out["mesh_n"], out["mesh_outside"] = hit.astype(np.int8), t, n, outside.astype(np.int8)
    img, alive, seg = ref.render(ms, orclib.MODE_FULL, 1, 2)
    out["img_mesh"], out["alive_mesh"] = img, alive
    np.savez_compressed(os.path.join(HERE, "extension_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "extension_vectors.npz"), os.path.getsize(os.path.join(HERE, "extension_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
