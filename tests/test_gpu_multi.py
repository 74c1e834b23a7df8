"""Several GPUs through the C++ library's own NCCL call sites (include/ptb200.h: ptb_multi_*, ptb_comm_*): SURVEY.md 8(e).
Skipped on boxes with fewer than two GPUs (the driver's multi-GPU tiers and `gpurun --gpus N` run them)."""
import numpy as np
import pytest

from common import pkg, same_bits

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.fixture(scope="module")
def m():
    return pkg()


def test_multi_with_one_device_is_the_single_gpu_render(m, scene_dir):
    """G = 1: no communicator, the frame is the context's own accumulation; same bits as a plain Renderer."""
    s = m.Scene.load(scene_dir + "/cornell.txt")
    with m.Renderer(s, device=0, bsdf=1, batch_iters=4) as r:
        r.render(1, 8)
        want = r.download()
        want8 = r.resolve_rgb8(8)
    with m.MultiRenderer(s, [0], bsdf=1, batch_iters=4) as mr:
        mr.render(1, 8)
        img, rgb, ms = mr.frame(8, want_sum=True, want_rgb8=True)
        st = mr.stats()
    assert same_bits(img, want) and np.array_equal(rgb, want8) and st["iterations"] == 8 and ms == 0.0


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_spp_split_with_nccl_reduce_inside_the_library(m, scene_dir):
    """G devices under one process: interleaved iterations, one grouped ncclReduce of the 3*W*H float accumulation buffers to
    device 0, RGB8 resolve fused after it.  The reduced frame equals the 1-GPU frame up to float summation order
    (n * 2^-24 relative); the per-depth path counts are exact; a second frame keeps accumulating."""
    G = min(_ngpus(), 8)
    s = m.Scene.load(scene_dir + "/cornell.txt")
    n = 8 * G
    with m.Renderer(s, device=0, bsdf=1, batch_iters=4) as r:
        r.render(1, n)
        want = r.download()
        wst = r.stats()
        r.render(n + 1, n)
        want2 = r.download()
    with m.MultiRenderer(s, list(range(G)), bsdf=1, batch_iters=4) as mr:
        mr.render(1, n)
        img, rgb, ms = mr.frame(n, want_sum=True, want_rgb8=True)
        st = mr.stats()
        mr.render(n + 1, n)
        img2, _, _ = mr.frame(2 * n)
    assert st["iterations"] == n and st["alive_per_depth"] == wst["alive_per_depth"] and st["segments"] == wst["segments"]
    assert ms > 0.0
    bound = 2.0 * n * 2.0 ** -24
    lit = want.max(1) > 0
    assert np.all(np.abs(img - want)[lit] <= bound * want[lit].max(1, keepdims=True))
    assert np.array_equal(img == 0, want == 0)
    assert np.all(np.abs(img2 - want2)[lit] <= 2 * bound * want2[lit].max(1, keepdims=True))
    flip = np.clip(want.reshape(s.height, s.width, 3)[:, ::-1] / n, 0, 1)
    assert np.abs(rgb.astype(np.int32) - (flip * 255).astype(np.int32)).max() <= 1
