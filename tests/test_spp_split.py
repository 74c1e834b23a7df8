"""CPU tests of the N>1 path: the iteration split and the single reduce per frame, world_size 2 over
gloo.  The renderer on each rank here is the CPU oracle (no GPU in this container); the host-side
logic under test (spp_split.py) is the same code bench.py runs over NCCL."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from common import PKG, ROOT, HERE


def split():
    return importlib.import_module(PKG + ".spp_split")


def test_rank_iterations_partition():
    sp = split()
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 5, 8, 17, 5000):
            seen = []
            for rank in range(world):
                b, c, s = sp.rank_iterations(1, total, rank, world)
                seen += [b + k * s for k in range(c)]
            assert sorted(seen) == list(range(1, total + 1)), (world, total)
            counts = [sp.rank_iterations(1, total, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sp.rank_iterations(1, 10, 2, 2)


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    import orclib
    import scenes
    sp = importlib.import_module(PKG + ".spp_split")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    path = os.path.join(tmpdir, "tiny_%d.txt" % rank)
    with open(path, "w") as f:
        f.write(scenes.cornell_text(res=(40, 30)))
    orc = orclib.load("port")
    scene = orc.load_scene(path)
    total = 7
    accum = np.zeros((scene.pixels, 3), np.float32)
    frames = []
    for frame_idx, (first, n) in enumerate(((1, 4), (5, 3))):        # two frames, progressive
        b, c, s = sp.rank_iterations(first, n, rank, world)
        for k in range(c):                                           # stride > 1: one iteration at a time
            orc.render(scene, orclib.MODE_FULL, b + k * s, 1, image=accum, nthreads=1)
        frames.append(sp.reduce_frame(torch.from_numpy(accum)).numpy().copy())
    ok = True
    if rank == 0:
        want4, _, _ = orc.render(scene, orclib.MODE_FULL, 1, 4, nthreads=1)
        want7, _, _ = orc.render(scene, orclib.MODE_FULL, 1, total, nthreads=1)
        ok = np.allclose(frames[0], want4, rtol=2e-6, atol=0) and np.allclose(frames[1], want7, rtol=2e-6, atol=0)
        ok = ok and want7.any()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if not flag.item():
        raise SystemExit(3)


def test_two_rank_spp_split_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
