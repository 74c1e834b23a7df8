"""Scene fixtures in the reference's scene-file format (INSTRUCTION.md:258-301, parser src/scene.cpp).

The reference ships two scenes, scenes/cornell.txt and scenes/sphere.txt.  /root/reference does not
exist on the GPU box and reference files are not copied into this repository, so the two scenes are
re-typed here as compact tables and written out in the reference's text format by `emit()`.
tests/test_oracle_vs_ref.py checks (in the build container, where /root/reference is mounted) that
the emitted files parse to byte-identical Geom/Material/Camera arrays as the reference's own files.

Also here: the synthetic many-object scene of BASELINE.json configs[4],
generated from a fixed seed so it is reproducible everywhere.
"""
import os
import numpy as np

# (comment, RGB, SPECEX, SPECRGB, REFL, REFR, REFRIOR, EMITTANCE)
_CORNELL_MATERIALS = [
    ("Emissive material (light)", "1 1 1", "0", "0 0 0", "0", "0", "0", "5"),
    ("Diffuse white", ".98 .98 .98", "0", "0 0 0", "0", "0", "0", "0"),
    ("Diffuse red", ".85 .35 .35", "0", "0 0 0", "0", "0", "0", "0"),
    ("Diffuse green", ".35 .85 .35", "0", "0 0 0", "0", "0", "0", "0"),
    ("Specular white", ".98 .98 .98", "0", ".98 .98 .98", "1", "0", "0", "0"),
]
# (comment, type, material, TRANS, ROTAT, SCALE)
_CORNELL_OBJECTS = [
    ("Ceiling light", "cube", 0, "0 10 0", "0 0 0", "3 .3 3"),
    ("Floor", "cube", 1, "0 0 0", "0 0 0", "10 .01 10"),
    ("Ceiling", "cube", 1, "0 10 0", "0 0 90", ".01 10 10"),
    ("Back wall", "cube", 1, "0 5 -5", "0 90 0", ".01 10 10"),
    ("Left wall", "cube", 2, "-5 5 0", "0 0 0", ".01 10 10"),
    ("Right wall", "cube", 3, "5 5 0", "0 0 0", ".01 10 10"),
    ("Sphere", "sphere", 4, "-1 4 -1", "0 0 0", "3 3 3"),
]
_SPHERE_MATERIALS = [_CORNELL_MATERIALS[0]]
_SPHERE_OBJECTS = [("Sphere", "sphere", 0, "0 0 0", "0 0 0", "3 3 3")]


def emit(materials, objects, res=(800, 800), fovy="45", iterations=5000, depth=8, name="cornell",
         eye="0.0 5 10.5", lookat="0 5 0", up="0 1 0"):
    """Return scene text.  Loader quirks honoured (SURVEY.md 7 hard part 7): sequential ids, exactly
    seven material lines, five fixed camera lines, blank line after every block, no trailing blanks."""
    out = []
    for i, (c, rgb, sx, srgb, refl, refr, ior, emit_) in enumerate(materials):
        out += ["// %s" % c, "MATERIAL %d" % i, "RGB         %s" % rgb, "SPECEX      %s" % sx,
                "SPECRGB     %s" % srgb, "REFL        %s" % refl, "REFR        %s" % refr,
                "REFRIOR     %s" % ior, "EMITTANCE   %s" % emit_, ""]
    out += ["// Camera", "CAMERA", "RES         %d %d" % res, "FOVY        %s" % fovy,
            "ITERATIONS  %d" % iterations, "DEPTH       %d" % depth, "FILE        %s" % name,
            "EYE         %s" % eye, "LOOKAT      %s" % lookat, "UP          %s" % up, "", ""]
    for i, (c, typ, mat, tr, ro, sc) in enumerate(objects):
        out += ["// %s" % c, "OBJECT %d" % i, typ, "material %d" % mat, "TRANS       %s" % tr,
                "ROTAT       %s" % ro, "SCALE       %s" % sc, ""]
    return "\n".join(out)


def cornell_text(**kw):
    return emit(_CORNELL_MATERIALS, _CORNELL_OBJECTS, name="cornell", **kw)


def closed_cornell_text(**kw):
    """cornell.txt with a front wall behind the camera: no path leaves the box, so every path runs until it meets the light or
    the depth limit (the "open vs closed scene" comparison of the stream-compaction analysis, INSTRUCTION.md:78-89, 339-355)."""
    objs = list(_CORNELL_OBJECTS) + [("Front wall", "cube", 1, "0 5 11", "0 90 0", ".01 10 10")]
    return emit(_CORNELL_MATERIALS, objs, name="cornell_closed", **kw)


def sphere_text(**kw):
    return emit(_SPHERE_MATERIALS, _SPHERE_OBJECTS, name="sphere", **kw)


def glass_text(**kw):
    """cornell with the sphere turned into glass (REFR 1, REFRIOR 1.5) and a second mirror sphere:
    exercises the refractive branch of the authored scatterRay body."""
    mats = list(_CORNELL_MATERIALS) + [("Glass", ".98 .98 .98", "0", ".98 .98 .98", "0", "1", "1.5", "0")]
    objs = list(_CORNELL_OBJECTS)
    objs[6] = ("Glass sphere", "sphere", 5, "-1 4 -1", "0 0 0", "3 3 3")
    objs.append(("Mirror sphere", "sphere", 4, "2.5 2 1.5", "0 30 0", "2 2.5 2"))
    objs.append(("Tilted box", "cube", 2, "-3 1.5 2", "20 35 10", "1.5 3 1.5"))
    return emit(mats, objs, name="glass", **kw)


def glossy_text(**kw):
    """cornell with imperfect-specular materials (REFL 1, SPECEX > 0; INSTRUCTION.md:202-216): a wide lobe
    (SPECEX 8) on the sphere, a tight one (SPECEX 200) on a tilted box and on the back wall."""
    mats = list(_CORNELL_MATERIALS) + [("Glossy wide", ".9 .9 .9", "8", ".9 .9 .9", "1", "0", "0", "0"),
                                       ("Glossy tight", ".9 .8 .6", "200", ".9 .8 .6", "1", "0", "0", "0")]
    objs = list(_CORNELL_OBJECTS)
    objs[3] = ("Back wall", "cube", 6, "0 5 -5", "0 90 0", ".01 10 10")
    objs[6] = ("Glossy sphere", "sphere", 5, "-1 4 -1", "0 0 0", "3 3 3")
    objs.append(("Tilted glossy box", "cube", 6, "2.5 1.5 1", "0 30 0", "2 3 2"))
    return emit(mats, objs, name="glossy", **kw)


# ---- triangle meshes (`mesh <file.obj>` objects; extension of the scene format, INSTRUCTION.md:291-293) ----
def icosphere(subdiv=1, radius=0.5):
    """Unit-diameter icosphere: 20 * 4^subdiv counter-clockwise (outward-facing) triangles."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        mid, nf = {}, []
        def m(a, b):
            k = (min(a, b), max(a, b))
            if k not in mid:
                p = v[a] + v[b]
                v.append(p / np.linalg.norm(p))
                mid[k] = len(v) - 1
            return mid[k]
        for a, b, c in f:
            ab, bc, ca = m(a, b), m(b, c), m(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v) * radius, f


def box_quads():
    """The unit cube of the `cube` primitive (+-0.5) as six outward-facing quads (the OBJ reader fans them)."""
    v = [(x, y, z) for x in (-.5, .5) for y in (-.5, .5) for z in (-.5, .5)]
    f = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    return np.array(v, np.float64), f


def obj_text(verts, faces, slashes=False, negative=False):
    """Wavefront OBJ text; `slashes` writes v/vt/vn style references, `negative` relative (negative) indices."""
    out = ["# generated by tests/scenes.py", "o mesh"]
    out += ["v %.7f %.7f %.7f" % tuple(p) for p in verts]
    if slashes:
        out += ["vt 0 0", "vn 0 0 1"]
    n = len(verts)
    for face in faces:
        ref = [(i - n) if negative else (i + 1) for i in face]
        out.append("f " + " ".join(("%d/1/1" % r) if slashes else str(r) for r in ref))
    return "\n".join(out) + "\n"


def parse_obj(text):
    """Independent (test-side) OBJ reader: -> (T, 9) float32 triangles, polygons fanned around their first vertex."""
    verts, tris = [], []
    for line in text.splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "v":
            verts.append([np.float32(float(x)) for x in tok[1:4]])
        elif tok[0] == "f":
            idx = [int(t.split("/")[0]) for t in tok[1:]]
            idx = [len(verts) + i if i < 0 else i - 1 for i in idx]
            for k in range(2, len(idx)):
                tris.append(verts[idx[0]] + verts[idx[k - 1]] + verts[idx[k]])
    return np.array(tris, np.float32).reshape(-1, 9)


def mesh_scene_files(dirname, res=(800, 800), subdiv=1, name="meshes"):
    """cornell with the sphere replaced by an icosphere MESH (mirror), a glass-like diffuse box mesh made of quads
    (rotated, non-uniformly scaled) and a small diffuse icosphere; writes the OBJ files next to the scene file and
    returns the scene path."""
    os.makedirs(dirname, exist_ok=True)
    v, f = icosphere(subdiv)
    open(os.path.join(dirname, "ico.obj"), "w").write(obj_text(v, f))
    bv, bf = box_quads()
    open(os.path.join(dirname, "box.obj"), "w").write(obj_text(bv, bf, slashes=True))
    v0, f0 = icosphere(0)
    open(os.path.join(dirname, "ico0.obj"), "w").write(obj_text(v0, f0, negative=True))
    mats = list(_CORNELL_MATERIALS) + [("Glass", ".98 .98 .98", "0", ".98 .98 .98", "0", "1", "1.5", "0")]
    objs = list(_CORNELL_OBJECTS)
    objs[6] = ("Mirror icosphere", "mesh ico.obj", 4, "-1 4 -1", "10 20 30", "3 3 3")
    objs.append(("Box mesh", "mesh box.obj", 2, "2.5 1.5 1", "20 35 10", "1.5 3 1.5"))
    objs.append(("Glass icosahedron", "mesh ico0.obj", 5, "-3 1.2 2.5", "0 15 0", "2.2 2.2 2.2"))
    objs.append(("Small sphere", "sphere", 3, "3 6 -2", "0 0 0", "1.5 1.5 1.5"))
    path = os.path.join(dirname, name + ".txt")
    open(path, "w").write(emit(mats, objs, res=res, name=name))
    return path


def synthetic_text(n_objects=100000, seed=565, res=(1920, 1080), depth=16, iterations=64):
    """BASELINE.json configs[4]: mixed SPHERE/CUBE scene, diffuse/REFL/REFR materials, one large
    ceiling light and a floor, horizontal -z view (the
    camera fix-up is only an identity for horizontal views along -z)."""
    rng = np.random.RandomState(seed)
    mats = [
        ("light", "1 1 1", "0", "0 0 0", "0", "0", "0", "5"),
        ("white", ".9 .9 .9", "0", "0 0 0", "0", "0", "0", "0"),
        ("red", ".85 .35 .35", "0", "0 0 0", "0", "0", "0", "0"),
        ("green", ".35 .85 .35", "0", "0 0 0", "0", "0", "0", "0"),
        ("blue", ".35 .35 .85", "0", "0 0 0", "0", "0", "0", "0"),
        ("mirror", ".98 .98 .98", "0", ".98 .98 .98", "1", "0", "0", "0"),
        ("glass", ".98 .98 .98", "0", ".98 .98 .98", "0", "1", "1.5", "0"),
    ]
    objs = [("ceiling light", "cube", 0, "0 110 -50", "0 0 0", "140 1 140"),
            ("floor", "cube", 1, "0 -1 -50", "0 0 0", "140 1 140")]
    n = n_objects - len(objs)
    centres = np.stack([rng.uniform(-50, 50, n), rng.uniform(0, 100, n), rng.uniform(-100, 0, n)], 1)
    scales = rng.uniform(0.3, 1.2, (n, 3))
    rots = rng.uniform(0, 360, (n, 3))
    kinds = rng.randint(0, 2, n)
    matsel = rng.randint(1, 7, n)
    emitters = rng.uniform(0, 1, n) < 0.01
    for i in range(n):
        m = 0 if emitters[i] else int(matsel[i])
        objs.append(("o%d" % i, "sphere" if kinds[i] == 0 else "cube", m,
                     "%.4f %.4f %.4f" % tuple(centres[i]), "%.3f %.3f %.3f" % tuple(rots[i]),
                     "%.4f %.4f %.4f" % tuple(scales[i])))
    return emit(mats, objs, res=res, depth=depth, iterations=iterations, name="synthetic",
                eye="0 50 60", lookat="0 50 0")


def write_all(dirname, synthetic_objects=0):
    os.makedirs(dirname, exist_ok=True)
    files = {"cornell.txt": cornell_text(), "sphere.txt": sphere_text(), "glass.txt": glass_text(), "glossy.txt": glossy_text()}
    if synthetic_objects:
        files["synthetic.txt"] = synthetic_text(synthetic_objects)
    for name, text in files.items():
        with open(os.path.join(dirname, name), "w") as f:
            f.write(text)
    return {k: os.path.join(dirname, k) for k in files}
