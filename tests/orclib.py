"""ctypes access to the CPU checkers under oracle/ (orc_api.h).  TEST INFRASTRUCTURE.

`load("port")`      -> oracle/libptoracle.so   (plain-C restatement; always available after build())
`load("reference")` -> oracle/_ref/libptref.so (the reference's own headers compiled; None if absent)
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GEOM_DT = np.dtype([("type", "<i4"), ("materialid", "<i4"), ("translation", "<f4", 3),
                    ("rotation", "<f4", 3), ("scale", "<f4", 3), ("transform", "<f4", 16),
                    ("inverseTransform", "<f4", 16), ("invTranspose", "<f4", 16)])
MATERIAL_DT = np.dtype([("color", "<f4", 3), ("specExponent", "<f4"), ("specColor", "<f4", 3),
                        ("hasReflective", "<f4"), ("hasRefractive", "<f4"),
                        ("indexOfRefraction", "<f4"), ("emittance", "<f4")])
CAMERA_DT = np.dtype([("resolution", "<i4", 2), ("position", "<f4", 3), ("lookAt", "<f4", 3),
                      ("view", "<f4", 3), ("up", "<f4", 3), ("right", "<f4", 3), ("fov", "<f4", 2),
                      ("pixelLength", "<f4", 2)])
assert GEOM_DT.itemsize == 236 and MATERIAL_DT.itemsize == 44 and CAMERA_DT.itemsize == 84

MODE_GOLDEN_DIFFUSE, MODE_FULL = 0, 1
RAYGEN_DEPTH_KEY = 511


MESH = 2            # orc_api.h ORC_MESH
TRIANGLE_DT = np.dtype([("v0", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3)])


class OrcMeshes(C.Structure):
    """orc_api.h OrcMeshes: triangles + (first, count) per geom"""
    _fields_ = [("triangles", C.c_void_p), ("ntriangles", C.c_int), ("ranges", C.c_void_p)]


def _meshes(scene_or_pair):
    """-> (OrcMeshes or None, keepalive): accepts a Scene with .triangles / .mesh_ranges or a (triangles, ranges) pair"""
    if scene_or_pair is None:
        return None, None
    tri, rng = scene_or_pair if isinstance(scene_or_pair, tuple) else (getattr(scene_or_pair, "triangles", None), getattr(scene_or_pair, "mesh_ranges", None))
    if tri is None or len(tri) == 0:
        return None, None
    tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
    rng = np.ascontiguousarray(rng, np.int32).reshape(-1, 2)
    return OrcMeshes(tri.ctypes.data, len(tri), rng.ctypes.data), (tri, rng)


class OrcLens(C.Structure):
    """orc_api.h OrcLens: stochastic antialiasing + thin lens"""
    _fields_ = [("antialias", C.c_int), ("lensRadius", C.c_float), ("focalDistance", C.c_float)]

_PATHS = {
    "port": os.path.join(ROOT, "oracle", "libptoracle.so"),
    "reference": os.path.join(ROOT, "oracle", "_ref", "libptref.so"),
    "reference_fast": os.path.join(ROOT, "oracle", "_ref", "libptref_fast.so"),
}


class Scene:
    def __init__(self, geoms, materials, camera, iterations, depth, name, triangles=None, mesh_ranges=None):
        self.geoms, self.materials, self.camera = geoms, materials, camera
        self.iterations, self.depth, self.name = iterations, depth, name
        # triangle meshes (geoms of type MESH): (T, 9) object-space vertices and (G, 2) first / count per geom
        self.triangles = None if triangles is None else np.ascontiguousarray(triangles, np.float32).reshape(-1, 9)
        self.mesh_ranges = None if mesh_ranges is None else np.ascontiguousarray(mesh_ranges, np.int32).reshape(-1, 2)

    @property
    def width(self):
        return int(self.camera["resolution"][0][0])

    @property
    def height(self):
        return int(self.camera["resolution"][0][1])

    @property
    def pixels(self):
        return self.width * self.height


class Oracle:
    def __init__(self, path):
        self.path = path
        self.lib = L = C.CDLL(path)
        L.orc_kind.restype = C.c_char_p
        L.orc_scene_load.restype = C.c_void_p
        L.orc_scene_load.argtypes = [C.c_char_p]
        L.orc_scene_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_scene_get.argtypes = [C.c_void_p] * 6 + [C.c_char_p, C.c_int]
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_camera_fixup.argtypes = [C.c_void_p]
        L.orc_build_transform.argtypes = [C.c_void_p] * 6
        L.orc_rng_u01.argtypes = [C.c_int] * 4 + [C.c_void_p]
        L.orc_generate_rays.argtypes = [C.c_void_p] * 3
        L.orc_closest_hit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.orc_hemisphere.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.orc_shade.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 12
        L.orc_render.restype = C.c_longlong
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_generate_rays_lens.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_render_ex.restype = C.c_longlong
        L.orc_render_ex.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_closest_hit_mesh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        self.kind = L.orc_kind().decode()
        if hasattr(L, "orc_sincos"):           # plain-C restatement only
            L.orc_sincos.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
            L.orc_sincos_exhaustive.restype = C.c_longlong
            L.orc_sincos_exhaustive.argtypes = [C.c_uint, C.c_uint, C.c_int]

    def sincos(self, angles, restated):
        """(sin, cos) per angle: libm's sinf / cosf (restated=False, what the oracle loops call) or the
        double-precision restatement the CUDA path mirrors (restated=True)."""
        a = np.ascontiguousarray(angles, np.float32)
        sn, cs = np.zeros_like(a), np.zeros_like(a)
        self.lib.orc_sincos(a.ctypes.data, len(a), int(bool(restated)), sn.ctypes.data, cs.ctypes.data)
        return sn, cs

    def sincos_mismatches(self, lo_bits, hi_bits, nthreads=0):
        return int(self.lib.orc_sincos_exhaustive(int(lo_bits), int(hi_bits), nthreads))

    # ---- scene ----
    def load_scene(self, path, fixup=True):
        h = self.lib.orc_scene_load(path.encode())
        if not h:
            raise IOError("cannot open scene file %s" % path)
        ng, nm = C.c_int(), C.c_int()
        self.lib.orc_scene_counts(h, C.addressof(ng), C.addressof(nm))
        geoms = np.zeros(ng.value, GEOM_DT)
        mats = np.zeros(nm.value, MATERIAL_DT)
        cam = np.zeros(1, CAMERA_DT)
        it, dp = C.c_int(), C.c_int()
        name = C.create_string_buffer(256)
        self.lib.orc_scene_get(h, geoms.ctypes.data, mats.ctypes.data, cam.ctypes.data,
                               C.addressof(it), C.addressof(dp), name, 256)
        self.lib.orc_scene_free(h)
        if fixup:
            self.lib.orc_camera_fixup(cam.ctypes.data)
        return Scene(geoms, mats, cam, it.value, dp.value, name.value.decode())

    def camera_fixup(self, cam):
        cam = cam.copy()
        self.lib.orc_camera_fixup(cam.ctypes.data)
        return cam

    def build_transform(self, t, r, s):
        t, r, s = (np.ascontiguousarray(a, np.float32) for a in (t, r, s))
        out = [np.zeros(16, np.float32) for _ in range(3)]
        self.lib.orc_build_transform(t.ctypes.data, r.ctypes.data, s.ctypes.data, *(o.ctypes.data for o in out))
        return out

    # ---- hot path pieces ----
    def rng_u01(self, it, index, depth, n):
        out = np.zeros(n, np.float32)
        self.lib.orc_rng_u01(it, index, depth, n, out.ctypes.data)
        return out

    def generate_rays(self, cam):
        P = int(cam["resolution"][0][0]) * int(cam["resolution"][0][1])
        o = np.zeros((P, 3), np.float32)
        d = np.zeros((P, 3), np.float32)
        self.lib.orc_generate_rays(cam.ctypes.data, o.ctypes.data, d.ctypes.data)
        return o, d

    def generate_rays_lens(self, cam, iteration, antialias=0, lens_radius=0.0, focal_distance=0.0):
        P = int(cam["resolution"][0][0]) * int(cam["resolution"][0][1])
        o = np.zeros((P, 3), np.float32)
        d = np.zeros((P, 3), np.float32)
        lens = OrcLens(int(antialias), float(lens_radius), float(focal_distance))
        self.lib.orc_generate_rays_lens(cam.ctypes.data, C.addressof(lens), int(iteration), o.ctypes.data, d.ctypes.data)
        return o, d

    def closest_hit(self, geoms, o, d, meshes=None):
        """meshes: None, a Scene carrying .triangles / .mesh_ranges, or a (triangles, ranges) pair"""
        o = np.ascontiguousarray(o, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        n = o.shape[0]
        hit = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32)
        outside = np.zeros(n, np.int32)
        ms, keep = _meshes(meshes)
        self.lib.orc_closest_hit_mesh(geoms.ctypes.data, len(geoms), C.addressof(ms) if ms else None, o.ctypes.data, d.ctypes.data, n,
                                      hit.ctypes.data, t.ctypes.data, nrm.ctypes.data, outside.ctypes.data)
        return hit, t, nrm, outside

    def hemisphere(self, normals, it, index, depth):
        normals = np.ascontiguousarray(normals, np.float32)
        n = normals.shape[0]
        it, index, depth = (np.ascontiguousarray(a, np.int32) for a in (it, index, depth))
        out = np.zeros((n, 3), np.float32)
        self.lib.orc_hemisphere(normals.ctypes.data, n, it.ctypes.data, index.ctypes.data,
                                depth.ctypes.data, out.ctypes.data)
        return out

    def shade(self, materials, mode, hit, material_ids, t, normal, outside, it, index, depth,
              origins, dirs, colors, remaining):
        """Returns new (origins, dirs, colors, remaining); inputs are not modified."""
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        f32 = lambda a: np.array(a, np.float32, copy=True, order="C")
        hit, material_ids, outside, it, index, depth = map(i32, (hit, material_ids, outside, it, index, depth))
        t, normal = np.ascontiguousarray(t, np.float32), np.ascontiguousarray(normal, np.float32)
        o, d, c = f32(origins), f32(dirs), f32(colors)
        rem = np.array(remaining, np.int32, copy=True)
        self.lib.orc_shade(materials.ctypes.data, mode, len(hit), hit.ctypes.data, material_ids.ctypes.data,
                           t.ctypes.data, normal.ctypes.data, outside.ctypes.data, it.ctypes.data,
                           index.ctypes.data, depth.ctypes.data, o.ctypes.data, d.ctypes.data,
                           c.ctypes.data, rem.ctypes.data)
        return o, d, c, rem

    def render(self, scene, mode, iter_begin, n_iters, depth=None, image=None, nthreads=0,
               antialias=0, lens_radius=0.0, focal_distance=0.0):
        """Returns (image_sum[P,3] float32, alive_per_depth int64, segments)."""
        depth = scene.depth if depth is None else depth
        if image is None:
            image = np.zeros((scene.pixels, 3), np.float32)
        alive = np.zeros(max(depth, 1), np.int64)
        ms, keep = _meshes(scene)
        if antialias or lens_radius > 0.0 or ms is not None:
            lens = OrcLens(int(antialias), float(lens_radius), float(focal_distance))
            seg = self.lib.orc_render_ex(scene.geoms.ctypes.data, len(scene.geoms), C.addressof(ms) if ms else None,
                                         scene.materials.ctypes.data, len(scene.materials), scene.camera.ctypes.data,
                                         C.addressof(lens), depth, mode,
                                         iter_begin, n_iters, image.ctypes.data, alive.ctypes.data, nthreads)
            return image, alive, int(seg)
        seg = self.lib.orc_render(scene.geoms.ctypes.data, len(scene.geoms), scene.materials.ctypes.data,
                                  len(scene.materials), scene.camera.ctypes.data, depth, mode,
                                  iter_begin, n_iters, image.ctypes.data, alive.ctypes.data, nthreads)
        return image, alive, int(seg)


_cache = {}


def load(kind="port"):
    """Return the Oracle for `kind`, or None when that library has not been built."""
    if kind not in _cache:
        p = _PATHS[kind]
        _cache[kind] = Oracle(p) if os.path.exists(p) else None
    return _cache[kind]
