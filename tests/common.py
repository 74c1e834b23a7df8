"""Shared helpers for the tests."""
import importlib
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "project3-cuda-path-tracer_b200"
REF_ROOT = "/root/reference"


def pkg():
    return importlib.import_module(PKG)


_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))
    return _golden


def golden_image():
    return np.load(os.path.join(HERE, "golden", "reference_cornell_5000samp.npz"))["rgb8"]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_bits(a, b):
    return np.array_equal(bits(a), bits(b))


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def densify(idx, rgb, pixels):
    img = np.zeros((pixels, 3), np.float32)
    img[idx] = rgb
    return img


def rel_err_lit(img, ref):
    """per-pixel relative error over pixels that carry light in either image"""
    lit = (ref.max(1) > 0) | (img.max(1) > 0)
    denom = np.maximum(np.maximum(ref.max(1), img.max(1)), 1e-12)
    return np.abs(img - ref).max(1)[lit] / denom[lit], lit


def to_oracle_scene(ms):
    """A product-side Scene (pkg().Scene, loaded by the host C++ loader) as an orclib.Scene for the CPU checkers:
    the POD arrays are byte-identical (src/sceneStructs.h), only the numpy field names differ."""
    import orclib
    return orclib.Scene(ms.geoms.view(orclib.GEOM_DT), ms.materials.view(orclib.MATERIAL_DT), ms.camera.view(orclib.CAMERA_DT),
                        ms.iterations, ms.traceDepth, ms.imageName,
                        ms.triangles if len(ms.triangles) else None, ms.mesh_ranges if len(ms.triangles) else None)


_tuning = None


def pkg_tuning():
    """The package bound to libptb200_tuning.so (csrc/Makefile `tuning`, -DPTB_TUNING): the same library with the environment
    knobs of A/B and validation runs compiled in.  Tests use it ONLY for failing controls (boxes that are known to lose
    reference hits); the release library never reads the environment."""
    global _tuning
    if _tuning is None:
        import importlib.util
        base = pkg()
        spec = importlib.util.spec_from_file_location("ptb200_tuning", os.path.join(os.path.dirname(base.__file__), "__init__.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.LIB_PATH = os.path.join(os.path.dirname(base.__file__), "libptb200_tuning.so")
        _tuning = mod
    return _tuning
