"""Generates the committed golden fixtures under tests/golden/ from the REFERENCE-LINKED build
(oracle/_ref/libptref.so: the reference's own intersections.h / interactions.h / scene.cpp compiled
where they lie under /root/reference, see oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own except one image,
img/REFERENCE_cornell.5000samp.png, which is stored here re-encoded as an npz array.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orclib   # noqa: E402
import scenes   # noqa: E402

REF_ROOT = "/root/reference"


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def sparse(img):
    idx = np.flatnonzero(img.any(1)).astype(np.int32)
    return idx, img[idx]


def main():
    ref = orclib.load("reference")
    assert ref is not None and ref.kind == "reference", "build oracle/_ref first (make -C oracle ref)"
    import tempfile
    d = tempfile.mkdtemp()
    files = scenes.write_all(d)
    out = {}

    # ---- scenes: parsed POD of the reference's own scene files ----
    for name in ("cornell", "sphere"):
        s = ref.load_scene(os.path.join(REF_ROOT, "scenes", name + ".txt"), fixup=False)
        f = ref.load_scene(os.path.join(REF_ROOT, "scenes", name + ".txt"), fixup=True)
        out["scene_%s_geoms" % name] = s.geoms.view(np.uint8)
        out["scene_%s_materials" % name] = s.materials.view(np.uint8)
        out["scene_%s_camera_raw" % name] = s.camera.view(np.uint8)
        out["scene_%s_camera_fixed" % name] = f.camera.view(np.uint8)
        out["scene_%s_meta" % name] = np.array([s.iterations, s.depth], np.int32)
    g = ref.load_scene(files["glass.txt"], fixup=True)
    out["scene_glass_geoms"] = g.geoms.view(np.uint8)
    out["scene_glass_camera_fixed"] = g.camera.view(np.uint8)

    # ---- RNG known answers (makeSeededRandomEngine + thrust u01) ----
    rs = np.random.RandomState(565)
    trip = np.stack([rs.randint(1, 1 << 22, 4096), rs.randint(0, 1 << 23, 4096), rs.randint(0, 512, 4096)], 1).astype(np.int32)
    trip[0] = (1, 0, 0)
    out["rng_triples"] = trip
    out["rng_u01"] = np.stack([ref.rng_u01(int(a), int(b), int(c), 4) for a, b, c in trip])

    # ---- transforms (buildTransformationMatrix + inverse + inverseTranspose) ----
    T = rs.uniform(-50, 50, (512, 3)).astype(np.float32)
    R = rs.uniform(-360, 360, (512, 3)).astype(np.float32)
    R[::4] = np.round(R[::4] / 90) * 90
    S = rs.uniform(0.01, 10, (512, 3)).astype(np.float32)
    mats = np.stack([np.stack(ref.build_transform(T[i], R[i], S[i])) for i in range(512)])
    out["xf_trs"] = np.stack([T, R, S], 1)
    out["xf_matrices"] = mats

    # ---- first bounce, cornell + sphere + glass (bit-exact contract) ----
    for name in ("cornell", "sphere", "glass"):
        s = ref.load_scene(files[name + ".txt"])
        o, dd = ref.generate_rays(s.camera)
        hit, t, n, outside = ref.closest_hit(s.geoms, o, dd)
        out["fb_%s_hit" % name] = hit.astype(np.int8)
        out["fb_%s_crc" % name] = np.array([crc(o), crc(dd), crc(t), crc(n), crc(outside)], np.uint32)
        out["fb_%s_t_stride7" % name] = t[::7]
        out["fb_%s_n_stride7" % name] = n[::7]
        out["fb_%s_dir_stride7" % name] = dd[::7]
        out["fb_%s_sum_t" % name] = np.array([t[hit >= 0].astype(np.float64).sum()])

    # ---- hemisphere sampler on the cornell first-bounce normals ----
    s = ref.load_scene(files["cornell.txt"])
    o, dd = ref.generate_rays(s.camera)
    hit, t, n, outside = ref.closest_hit(s.geoms, o, dd)
    sel = np.flatnonzero(hit > 0)[::97][:4096]
    its = rs.randint(1, 5000, len(sel)).astype(np.int32)
    dps = rs.randint(0, 8, len(sel)).astype(np.int32)
    out["hemi_normals"] = n[sel]
    out["hemi_seed"] = np.stack([its, sel.astype(np.int32), dps], 1)
    out["hemi_dirs"] = ref.hemisphere(n[sel], its, sel.astype(np.int32), dps)

    # ---- the normative loop: per-depth counts, radiance sums, sparse iter-1 images ----
    for name, mode, tag in (("cornell", 0, "cornell_diffuse"), ("cornell", 1, "cornell_full"), ("glass", 1, "glass_full")):
        s = ref.load_scene(files[name + ".txt"])
        img, alive, seg = ref.render(s, mode, 1, 1)
        idx, val = sparse(img)
        out["loop_%s_alive" % tag] = alive
        out["loop_%s_idx" % tag] = idx
        out["loop_%s_rgb" % tag] = val
        out["loop_%s_sum" % tag] = img.astype(np.float64).sum(0)
        img4, alive4, seg4 = ref.render(s, mode, 1, 4)
        out["loop_%s_alive4" % tag] = alive4
        out["loop_%s_crc4" % tag] = np.array([crc(img4)], np.uint32)
        out["loop_%s_sum4" % tag] = img4.astype(np.float64).sum(0)
    # small full images for cheap exact CPU tests
    for name, mode, tag in (("cornell", 0, "cornell_diffuse"), ("cornell", 1, "cornell_full")):
        p = os.path.join(d, "small.txt")
        with open(p, "w") as f:
            f.write(scenes.cornell_text(res=(96, 64)))
        s = ref.load_scene(p)
        img, alive, seg = ref.render(s, mode, 1, 8)
        out["small_%s_img8" % tag] = img
        out["small_%s_alive8" % tag] = alive
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)

    # ---- the one golden artefact the reference ships ----
    from PIL import Image
    png = np.asarray(Image.open(os.path.join(REF_ROOT, "img", "REFERENCE_cornell.5000samp.png")).convert("RGB"))
    assert png.shape == (800, 800, 3)
    np.savez_compressed(os.path.join(HERE, "reference_cornell_5000samp.npz"), rgb8=png)
    print("wrote", os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), os.path.getsize(os.path.join(HERE, "reference_cornell_5000samp.npz")))


if __name__ == "__main__":
    main()
