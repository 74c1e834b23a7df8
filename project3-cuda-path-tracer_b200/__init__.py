"""Host-side Python mirror of the reference's renderer interface over the C ABI (include/ptb200.h).

The reference drives its renderer through three free functions (src/pathtrace.h:6-8):
    pathtraceInit(Scene*)  /  pathtrace(pbo, frame, iteration)  /  pathtraceFree()
This module exposes the same three names with the same argument meaning for tests and benchmarks,
plus a `Renderer` object for the batched / multi-GPU entry points.  All compute happens in
libptb200.so (hand-written sm_100a kernels); there is no Python or CPU fallback -- importing works
without a GPU (so the symbol table can be checked), every render call raises without one.

The package directory name contains '-' (it mirrors the reference's repository name), so import it
with   importlib.import_module("project3-cuda-path-tracer_b200")   or via `ptb200_loader.py`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PTB200_LIB: load another build of the same library (tuning variants of tools/build_variant.sh under the GPU test suite)
LIB_PATH = os.environ.get("PTB200_LIB") or os.path.join(_HERE, "libptb200.so")

GEOM_DT = np.dtype([("type", "<i4"), ("materialid", "<i4"), ("translation", "<f4", 3),
                    ("rotation", "<f4", 3), ("scale", "<f4", 3), ("transform", "<f4", 16),
                    ("inverseTransform", "<f4", 16), ("invTranspose", "<f4", 16)])
MATERIAL_DT = np.dtype([("color", "<f4", 3), ("specularExponent", "<f4"), ("specularColor", "<f4", 3),
                        ("hasReflective", "<f4"), ("hasRefractive", "<f4"),
                        ("indexOfRefraction", "<f4"), ("emittance", "<f4")])
CAMERA_DT = np.dtype([("resolution", "<i4", 2), ("position", "<f4", 3), ("lookAt", "<f4", 3),
                      ("view", "<f4", 3), ("up", "<f4", 3), ("right", "<f4", 3), ("fov", "<f4", 2),
                      ("pixelLength", "<f4", 2)])
TRIANGLE_DT = np.dtype([("v0", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3)])     # PtbTriangle: object-space vertices
assert GEOM_DT.itemsize == 236 and MATERIAL_DT.itemsize == 44 and CAMERA_DT.itemsize == 84 and TRIANGLE_DT.itemsize == 36
GEOM_SPHERE, GEOM_CUBE, GEOM_MESH = 0, 1, 2

BSDF_GOLDEN_DIFFUSE, BSDF_FULL = 0, 1
PIPELINE_FUSED, PIPELINE_STAGED = 0, 1
ACCEL_AUTO, ACCEL_BRUTE, ACCEL_BVH = 0, 1, 2
DEPOSIT_ORDERED, DEPOSIT_ATOMIC = 0, 1
NUM_STAGES = 12
STAGE_NAMES = ["raygen", "intersect", "sort", "shade", "compact", "gather", "fused_bounce0", "fused_bounces", "control"]


class PtbCamera(C.Structure):
    _fields_ = [("resolution", C.c_int32 * 2), ("position", C.c_float * 3), ("lookAt", C.c_float * 3),
                ("view", C.c_float * 3), ("up", C.c_float * 3), ("right", C.c_float * 3),
                ("fov", C.c_float * 2), ("pixelLength", C.c_float * 2)]


class PtbSceneDesc(C.Structure):
    _fields_ = [("geoms", C.c_void_p), ("numGeoms", C.c_int32), ("materials", C.c_void_p),
                ("numMaterials", C.c_int32), ("camera", PtbCamera), ("traceDepth", C.c_int32),
                ("numTriangles", C.c_int32), ("triangles", C.c_void_p), ("meshRanges", C.c_void_p)]


class PtbOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("bsdfMode", C.c_int32), ("pipeline", C.c_int32), ("accel", C.c_int32),
                ("batchIters", C.c_int32), ("compaction", C.c_int32), ("sortMaterials", C.c_int32),
                ("firstBounceCache", C.c_int32), ("useGraphs", C.c_int32), ("antialias", C.c_int32),
                ("stream", C.c_void_p), ("externalAccum", C.c_void_p),
                ("lensRadius", C.c_float), ("focalDistance", C.c_float),
                ("launchPerBounce", C.c_int32), ("depositMode", C.c_int32), ("timeStages", C.c_int32), ("lazyImage", C.c_int32), ("regroupMaterials", C.c_int32), ("lookahead", C.c_int32), ("leaveTest", C.c_int32), ("hitQueue", C.c_int32)]


class PtbStats(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("segments", C.c_int64), ("kernelLaunches", C.c_int64),
                ("alivePerDepth", C.c_int64 * 64), ("traceDepth", C.c_int32), ("numGeoms", C.c_int32),
                ("bvhNodes", C.c_int32), ("batchIters", C.c_int32), ("pathStateBytes", C.c_int64),
                ("sceneUploadBytes", C.c_int64), ("stageUs", C.c_double * NUM_STAGES), ("stageLaunches", C.c_int64 * NUM_STAGES)]


class PtbError(RuntimeError):
    pass


# every symbol include/ptb200.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = [
    "ptb_default_options", "ptb_last_error", "ptb_abi_version", "ptb_create", "ptb_update_scene", "ptb_destroy",
    "ptb_pathtrace", "ptb_render", "ptb_download", "ptb_write_pbo", "ptb_reset", "ptb_sync",
    "ptb_stats", "ptb_accum_device_ptr", "ptb_resolve_rgb8", "ptb_stage_raygen", "ptb_stage_raygen_iter", "ptb_stage_intersect",
    "ptb_stage_shade", "ptb_stage_compact", "ptb_stage_sort_keys", "ptb_stage_rng", "ptb_stage_arith", "ptb_stage_sincos",
    "ptb_scene_load", "ptb_scene_counts", "ptb_scene_get", "ptb_scene_mesh_counts", "ptb_scene_get_meshes", "ptb_scene_free", "ptb_camera_fixup",
    "ptb_save_png", "ptb_save_hdr", "ptb_bvh_build_host", "ptb_geom_bounds_host",
    "ptb_comm_unique_id", "ptb_comm_init_rank", "ptb_comm_reduce", "ptb_frame_device_ptr", "ptb_frame_download", "ptb_frame_resolve_rgb8",
    "ptb_multi_create", "ptb_multi_destroy", "ptb_multi_reset", "ptb_multi_render", "ptb_multi_frame", "ptb_multi_stats", "ptb_multi_context",
]
COMM_ID_BYTES = 128

_lib = None


def lib():
    """Load libptb200.so; fail loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with __graft_entry__.build() (nvcc, sm_100a); "
                              "there is no fallback implementation" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.ptb_last_error.restype = C.c_char_p
        L.ptb_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(PtbSceneDesc), C.POINTER(PtbOptions)]
        L.ptb_destroy.argtypes = [C.c_void_p]
        L.ptb_update_scene.argtypes = [C.c_void_p, C.POINTER(PtbSceneDesc)]
        L.ptb_pathtrace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ptb_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ptb_download.argtypes = [C.c_void_p, C.c_void_p]
        L.ptb_write_pbo.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ptb_reset.argtypes = [C.c_void_p]
        L.ptb_sync.argtypes = [C.c_void_p]
        L.ptb_stats.argtypes = [C.c_void_p, C.POINTER(PtbStats)]
        L.ptb_accum_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.ptb_resolve_rgb8.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ptb_stage_raygen.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ptb_stage_raygen_iter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ptb_stage_intersect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.ptb_stage_shade.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11
        L.ptb_stage_compact.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        L.ptb_stage_sort_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ptb_stage_rng.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ptb_stage_arith.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.ptb_stage_sincos.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
        L.ptb_comm_unique_id.argtypes = [C.c_void_p]
        L.ptb_comm_init_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ptb_comm_reduce.argtypes = [C.c_void_p, C.c_int]
        L.ptb_frame_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.ptb_frame_download.argtypes = [C.c_void_p, C.c_void_p]
        L.ptb_frame_resolve_rgb8.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ptb_multi_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(PtbSceneDesc), C.POINTER(PtbOptions), C.c_void_p, C.c_int32]
        L.ptb_multi_destroy.argtypes = [C.c_void_p]
        L.ptb_multi_reset.argtypes = [C.c_void_p]
        L.ptb_multi_render.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ptb_multi_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.ptb_multi_stats.argtypes = [C.c_void_p, C.POINTER(PtbStats)]
        L.ptb_multi_context.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        if hasattr(L, "ptb_scene_load"):
            L.ptb_scene_load.restype = C.c_int
            L.ptb_scene_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
            L.ptb_scene_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
            L.ptb_scene_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.c_char_p, C.c_int]
            L.ptb_scene_free.argtypes = [C.c_void_p]
            L.ptb_scene_mesh_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
            L.ptb_scene_get_meshes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.ptb_camera_fixup.argtypes = [C.c_void_p]
            L.ptb_save_png.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int]
            L.ptb_save_hdr.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int]
            L.ptb_bvh_build_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
            L.ptb_geom_bounds_host.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 5
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise PtbError("ptb200 error %d: %s" % (rc, lib().ptb_last_error().decode()))


class Scene:
    """Mirror of the reference's `Scene` public members (src/scene.h:20-25): geoms, materials, state."""

    def __init__(self, geoms, materials, camera, iterations=5000, trace_depth=8, image_name="image",
                 triangles=None, mesh_ranges=None):
        self.geoms = np.ascontiguousarray(geoms, GEOM_DT)
        # `mesh` objects (extension): (T, 9) object-space vertices and (G, 2) first / count per geom
        self.triangles = np.ascontiguousarray(np.zeros((0, 9)) if triangles is None else triangles, np.float32).reshape(-1, 9)
        self.mesh_ranges = np.ascontiguousarray(np.zeros((len(self.geoms), 2)) if mesh_ranges is None else mesh_ranges, np.int32).reshape(-1, 2)
        self.materials = np.ascontiguousarray(materials, MATERIAL_DT)
        self.camera = np.ascontiguousarray(camera, CAMERA_DT).reshape(1)
        self.iterations = int(iterations)
        self.traceDepth = int(trace_depth)
        self.imageName = image_name
        # state.image: the host accumulation buffer pathtrace() fills (src/sceneStructs.h:58)
        self.image = np.zeros((self.pixels, 3), np.float32)

    @property
    def width(self):
        return int(self.camera["resolution"][0][0])

    @property
    def height(self):
        return int(self.camera["resolution"][0][1])

    @property
    def pixels(self):
        return self.width * self.height

    @classmethod
    def load(cls, filename, fixup=True):
        """`new Scene(filename)` (src/scene.cpp:7-33) followed, when `fixup`, by the camera fix-up the
        app performs before the first frame (src/main.cpp:53-67, 102-120).  Host C++ in libptb200.so."""
        L = lib()
        h = C.c_void_p()
        _check(L.ptb_scene_load(os.fsencode(filename), C.byref(h)))
        ng, nm = C.c_int32(), C.c_int32()
        L.ptb_scene_counts(h, C.byref(ng), C.byref(nm))
        geoms = np.zeros(ng.value, GEOM_DT)
        mats = np.zeros(nm.value, MATERIAL_DT)
        cam = np.zeros(1, CAMERA_DT)
        it, dp = C.c_int32(), C.c_int32()
        name = C.create_string_buffer(256)
        L.ptb_scene_get(h, geoms.ctypes.data, mats.ctypes.data, cam.ctypes.data, C.byref(it), C.byref(dp), name, 256)
        nt = C.c_int32()
        L.ptb_scene_mesh_counts(h, C.byref(nt))
        tris = np.zeros((nt.value, 9), np.float32)
        ranges = np.zeros((ng.value, 2), np.int32)
        L.ptb_scene_get_meshes(h, tris.ctypes.data, ranges.ctypes.data)
        L.ptb_scene_free(h)
        if fixup:
            L.ptb_camera_fixup(cam.ctypes.data)
        return cls(geoms, mats, cam, it.value, dp.value, name.value.decode(), tris, ranges)


class Renderer:
    """One device context (PtbContext).  Keyword options map 1:1 onto PtbOptions."""

    def __init__(self, scene, device=-1, bsdf=BSDF_FULL, pipeline=PIPELINE_FUSED, accel=ACCEL_AUTO,
                 batch_iters=0, compaction=1, sort_materials=0, first_bounce_cache=0, use_graphs=1,
                 stream=None, external_accum=None, trace_depth=None, antialias=0, lens_radius=0.0, focal_distance=0.0,
                 launch_per_bounce=0, deposit_mode=DEPOSIT_ORDERED, time_stages=0, lazy_image=0, regroup_materials=None, lookahead=None, leave_test=None, hit_queue=None):
        L = lib()
        self.scene = scene
        self.trace_depth = scene.traceDepth if trace_depth is None else int(trace_depth)
        desc = self._desc(scene)
        opt = PtbOptions()
        L.ptb_default_options(C.byref(opt))
        opt.device, opt.bsdfMode, opt.pipeline, opt.accel = device, bsdf, pipeline, accel
        opt.batchIters, opt.compaction, opt.sortMaterials = batch_iters, compaction, sort_materials
        opt.firstBounceCache, opt.useGraphs = first_bounce_cache, use_graphs
        opt.stream = stream
        opt.externalAccum = external_accum
        opt.antialias, opt.lensRadius, opt.focalDistance = int(antialias), float(lens_radius), float(focal_distance)
        opt.launchPerBounce, opt.depositMode, opt.timeStages, opt.lazyImage = int(launch_per_bounce), int(deposit_mode), int(time_stages), int(lazy_image)
        if regroup_materials is not None:                                  # None: the library's default
            opt.regroupMaterials = int(regroup_materials)
        if lookahead is not None:
            opt.lookahead = int(lookahead)
        if leave_test is not None:
            opt.leaveTest = int(leave_test)
        if hit_queue is not None:                                          # None: PTB_HIT_QUEUE_AUTO
            opt.hitQueue = int(hit_queue)
        self._h = C.c_void_p()
        _check(L.ptb_create(C.byref(self._h), C.byref(desc), C.byref(opt)))

    def _desc(self, scene):
        self._keep = (scene.geoms, scene.materials, scene.triangles, scene.mesh_ranges)
        desc = PtbSceneDesc()
        desc.numTriangles = len(scene.triangles)
        desc.triangles = scene.triangles.ctypes.data if len(scene.triangles) else None
        desc.meshRanges = scene.mesh_ranges.ctypes.data if len(scene.mesh_ranges) else None
        desc.geoms = scene.geoms.ctypes.data
        desc.numGeoms = len(scene.geoms)
        desc.materials = scene.materials.ctypes.data
        desc.numMaterials = len(scene.materials)
        C.memmove(C.addressof(desc.camera), scene.camera.ctypes.data, 84)
        desc.traceDepth = self.trace_depth
        return desc

    def update_scene(self, scene):
        """Re-upload geoms/materials/camera (same counts and resolution) and reset the accumulation:
        the camera-move path of the app (src/main.cpp:125-128) without reallocating."""
        desc = self._desc(scene)
        _check(lib().ptb_update_scene(self._h, C.byref(desc)))
        self.scene = scene

    def close(self):
        if self._h:
            lib().ptb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- the reference's loop ----
    def pathtrace(self, pbo, frame, iteration, host_image=None):
        ptr = host_image.ctypes.data if host_image is not None else None
        _check(lib().ptb_pathtrace(self._h, pbo, frame, iteration, ptr))

    def render(self, iter_begin, count, iter_stride=1):
        _check(lib().ptb_render(self._h, iter_begin, count, iter_stride))

    def download(self, out=None):
        if out is None:
            out = np.zeros((self.scene.pixels, 3), np.float32)
        _check(lib().ptb_download(self._h, out.ctypes.data))
        return out

    def download_to(self, host_ptr):
        """ptb_download into caller-owned host memory given as a raw address (e.g. a pinned buffer)."""
        _check(lib().ptb_download(self._h, host_ptr))

    def write_pbo(self, device_ptr, samples):
        _check(lib().ptb_write_pbo(self._h, device_ptr, samples))

    def reset(self):
        _check(lib().ptb_reset(self._h))

    def sync(self):
        _check(lib().ptb_sync(self._h))

    def stats(self):
        s = PtbStats()
        _check(lib().ptb_stats(self._h, C.byref(s)))
        return _stats_dict(s)

    # ---- several GPUs, one process per GPU: NCCL inside the library (include/ptb200.h, ptb_comm_*) ----
    def comm_init_rank(self, unique_id, rank, nranks):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        _check(lib().ptb_comm_init_rank(self._h, buf, int(rank), int(nranks)))

    def comm_reduce(self, root=0):
        """ncclReduce(sum) of this rank's accumulation buffer into the root's frame buffer (asynchronous, collective)"""
        _check(lib().ptb_comm_reduce(self._h, int(root)))

    def frame_device_ptr(self):
        p = C.c_void_p()
        _check(lib().ptb_frame_device_ptr(self._h, C.byref(p)))
        return p.value

    def frame_download(self, out=None):
        if out is None:
            out = np.zeros((self.scene.pixels, 3), np.float32)
        _check(lib().ptb_frame_download(self._h, out.ctypes.data))
        return out

    def frame_download_to(self, host_ptr):
        _check(lib().ptb_frame_download(self._h, host_ptr))

    def frame_resolve_rgb8(self, samples):
        out = np.zeros((self.scene.height, self.scene.width, 3), np.uint8)
        _check(lib().ptb_frame_resolve_rgb8(self._h, samples, out.ctypes.data))
        return out

    def accum_device_ptr(self):
        p = C.c_void_p()
        _check(lib().ptb_accum_device_ptr(self._h, C.byref(p)))
        return p.value

    def resolve_rgb8(self, samples):
        out = np.zeros((self.scene.height, self.scene.width, 3), np.uint8)
        _check(lib().ptb_resolve_rgb8(self._h, samples, out.ctypes.data))
        return out

    # ---- stage-level entry points ----
    def stage_raygen(self, iteration=1):
        o = np.zeros((self.scene.pixels, 3), np.float32)
        d = np.zeros((self.scene.pixels, 3), np.float32)
        _check(lib().ptb_stage_raygen_iter(self._h, int(iteration), o.ctypes.data, d.ctypes.data))
        return o, d

    def stage_intersect(self, origins, dirs, accel=ACCEL_AUTO):
        o = np.ascontiguousarray(origins, np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        n = o.shape[0]
        hit = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32)
        outside = np.zeros(n, np.int32)
        _check(lib().ptb_stage_intersect(self._h, accel, o.ctypes.data, d.ctypes.data, n, hit.ctypes.data,
                                         t.ctypes.data, nrm.ctypes.data, outside.ctypes.data))
        return hit, t, nrm, outside

    def stage_shade(self, hit, t, normal, outside, it, pixel_index, depth, origins, dirs, colors, remaining):
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        hit, outside, it, pixel_index, depth = map(i32, (hit, outside, it, pixel_index, depth))
        t = np.ascontiguousarray(t, np.float32)
        normal = np.ascontiguousarray(normal, np.float32)
        o = np.array(origins, np.float32, copy=True, order="C")
        d = np.array(dirs, np.float32, copy=True, order="C")
        c = np.array(colors, np.float32, copy=True, order="C")
        rem = np.array(remaining, np.int32, copy=True)
        _check(lib().ptb_stage_shade(self._h, len(hit), hit.ctypes.data, t.ctypes.data, normal.ctypes.data,
                                     outside.ctypes.data, it.ctypes.data, pixel_index.ctypes.data,
                                     depth.ctypes.data, o.ctypes.data, d.ctypes.data, c.ctypes.data, rem.ctypes.data))
        return o, d, c, rem

    def stage_compact(self, ids, flags):
        ids = np.ascontiguousarray(ids, np.int32)
        flags = np.ascontiguousarray(flags, np.int32)
        out = np.zeros(len(ids), np.int32)
        cnt = C.c_int32()
        _check(lib().ptb_stage_compact(self._h, len(ids), ids.ctypes.data, flags.ctypes.data, out.ctypes.data, C.byref(cnt)))
        return out[:cnt.value], cnt.value

    def stage_sort_keys(self, keys, num_keys):
        keys = np.ascontiguousarray(keys, np.int32)
        order = np.zeros(len(keys), np.int32)
        _check(lib().ptb_stage_sort_keys(self._h, len(keys), keys.ctypes.data, num_keys, order.ctypes.data))
        return order

    def stage_arith(self, a, b):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        q, r, ny = np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)
        _check(lib().ptb_stage_arith(self._h, len(a), a.ctypes.data, b.ctypes.data, q.ctypes.data, r.ctypes.data, ny.ctypes.data))
        return q, r, ny

    def stage_sincos(self, angles):
        a = np.ascontiguousarray(angles, np.float32)
        sn, cs = np.zeros_like(a), np.zeros_like(a)
        _check(lib().ptb_stage_sincos(self._h, len(a), a.ctypes.data, sn.ctypes.data, cs.ctypes.data))
        return sn, cs

    def stage_rng(self, it, index, depth, n):
        out = np.zeros(n, np.float32)
        _check(lib().ptb_stage_rng(self._h, it, index, depth, n, out.ctypes.data))
        return out


def comm_unique_id():
    """ncclGetUniqueId through the library: rank 0 calls it, the launcher's plumbing (torch.distributed, MPI, a file) broadcasts it."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    _check(lib().ptb_comm_unique_id(buf))
    return bytes(buf)


def _stats_dict(s):
    return {"iterations": s.iterations, "segments": s.segments, "kernel_launches": s.kernelLaunches,
            "alive_per_depth": [int(s.alivePerDepth[d]) for d in range(s.traceDepth)],
            "trace_depth": s.traceDepth, "num_geoms": s.numGeoms, "bvh_nodes": s.bvhNodes,
            "batch_iters": s.batchIters, "path_state_bytes": s.pathStateBytes,
            "scene_upload_bytes": s.sceneUploadBytes,
            "stage_us": {n: float(s.stageUs[i]) for i, n in enumerate(STAGE_NAMES)},
            "stage_launches": {n: int(s.stageLaunches[i]) for i, n in enumerate(STAGE_NAMES)}}


class MultiRenderer:
    """One process, several devices (PtbMulti): n contexts + ncclCommInitAll inside the library; device g renders iterations
    g+1, g+1+G, ... and `frame` is ONE grouped ncclReduce of the accumulation buffers to device 0."""

    def __init__(self, scene, devices, **kw):
        L = lib()
        self.scene = scene
        tmp = Renderer.__new__(Renderer)
        tmp.trace_depth = scene.traceDepth if kw.get("trace_depth") is None else int(kw["trace_depth"])
        desc = Renderer._desc(tmp, scene)
        self._keep = tmp._keep
        opt = PtbOptions()
        L.ptb_default_options(C.byref(opt))
        for k_py, k_c in (("bsdf", "bsdfMode"), ("pipeline", "pipeline"), ("accel", "accel"), ("batch_iters", "batchIters"),
                          ("first_bounce_cache", "firstBounceCache"), ("deposit_mode", "depositMode"), ("antialias", "antialias")):
            if k_py in kw:
                setattr(opt, k_c, int(kw[k_py]))
        devs = np.ascontiguousarray(devices, np.int32)
        self.devices = [int(d) for d in devs]
        self._h = C.c_void_p()
        _check(L.ptb_multi_create(C.byref(self._h), C.byref(desc), C.byref(opt), devs.ctypes.data, len(devs)))

    def close(self):
        if self._h:
            lib().ptb_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _check(lib().ptb_multi_reset(self._h))

    def render(self, iter_begin, count):
        _check(lib().ptb_multi_render(self._h, iter_begin, count))

    def frame(self, samples, want_sum=True, want_rgb8=False):
        """-> (sum image or None, rgb8 or None, milliseconds of the NCCL reduce on the root's stream)"""
        img = np.zeros((self.scene.pixels, 3), np.float32) if want_sum else None
        rgb = np.zeros((self.scene.height, self.scene.width, 3), np.uint8) if want_rgb8 else None
        ms = C.c_float(0.0)
        _check(lib().ptb_multi_frame(self._h, int(samples), img.ctypes.data if want_sum else None,
                                     rgb.ctypes.data if want_rgb8 else None, C.byref(ms)))
        return img, rgb, float(ms.value)

    def stats(self):
        s = PtbStats()
        _check(lib().ptb_multi_stats(self._h, C.byref(s)))
        return _stats_dict(s)


# ---- the reference's three entry points (src/pathtrace.h:6-8), global-state form ----
_current = None
_current_opts = {}


def set_options(**kw):
    """Options applied by the next pathtraceInit (the reference has compile-time toggles instead)."""
    _current_opts.clear()
    _current_opts.update(kw)


def pathtraceInit(scene):
    global _current
    if _current is not None:
        _current.close()
    _current = Renderer(scene, **_current_opts)


def pathtraceFree():
    """No-op when nothing is initialised: the app frees before the first init (src/main.cpp:126)."""
    global _current
    if _current is not None:
        _current.close()
        _current = None


def pathtrace(pbo, frame, iteration):
    """Traces iteration `iteration` and leaves the accumulated SUM in scene.image, exactly the
    post-condition of the reference (src/pathtrace.cu:276, 389-390)."""
    if _current is None:
        raise PtbError("pathtrace() before pathtraceInit()")
    _current.pathtrace(pbo, frame, iteration, _current.scene.image)


def save_png(filename_base, width, height, image_sum, samples, flip_x=True):
    """saveImage + image::savePNG (src/main.cpp:78-99, src/image.cpp:22-39): writes <base>.png"""
    img = np.ascontiguousarray(image_sum, np.float32)
    _check(lib().ptb_save_png(os.fsencode(filename_base), width, height, img.ctypes.data, float(samples), int(flip_x)))


BVH_NODE_DT = np.dtype([("llo", "<f4", 3), ("lref", "<i4"), ("lhi", "<f4", 3), ("lcount", "<i4"),
                        ("rlo", "<f4", 3), ("rref", "<i4"), ("rhi", "<f4", 3), ("rcount", "<i4")])   # csrc/ptb_bvh.h HostBvhNode


def bvh_build_host(geoms, max_leaf=1, far_pad_diameter=-1.0):
    """The flat BVH ptb_create builds over SPHERE / CUBE geoms (csrc/ptb_bvh.cpp), on the host: (nodes, order, max_depth).
    far_pad_diameter > 0: the rigorous sphere boxes for ray origins up to that far away (csrc/ptb_bvh.cpp geomBounds); 0: none;
    < 0: what ptb_create uses (the diameter of the scene, or none under PTB_BVH_FAR_PAD=0)."""
    geoms = np.ascontiguousarray(geoms, GEOM_DT)
    n, depth = C.c_int32(0), C.c_int32(0)
    _check(lib().ptb_bvh_build_host(geoms.ctypes.data, len(geoms), int(max_leaf), float(far_pad_diameter), None, 0, None, C.byref(n), None))
    nodes = np.zeros(n.value, BVH_NODE_DT)
    order = np.zeros(len(geoms), np.int32)
    _check(lib().ptb_bvh_build_host(geoms.ctypes.data, len(geoms), int(max_leaf), float(far_pad_diameter), nodes.ctypes.data, len(nodes), order.ctypes.data,
                                    C.byref(n), C.byref(depth)))
    return nodes, order, depth.value


def geom_bounds_host(geoms, reach=0.0):
    """Pruning bounds of SPHERE / CUBE geoms as the kernels receive them (csrc/ptb_bvh.cpp): (lo, hi, centre, radius2, errK);
    the last three are zero for cubes."""
    geoms = np.ascontiguousarray(geoms, GEOM_DT)
    n = len(geoms)
    lo, hi, c = (np.zeros((n, 3), np.float32) for _ in range(3))
    r2, ek = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for i in range(n):
        _check(lib().ptb_geom_bounds_host(geoms[i:i + 1].ctypes.data, float(reach), lo[i:].ctypes.data, hi[i:].ctypes.data,
                                          c[i:].ctypes.data, r2[i:].ctypes.data, ek[i:].ctypes.data))
    return lo, hi, c, r2, ek
