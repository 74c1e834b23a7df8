"""BASELINE.json configs[3] (cornell 3840x2160, depth 8, 5000 spp split across the GPUs with one NCCL reduce of the
accumulation buffer per frame) and configs[4] (synthetic 100k objects, 1920x1080, depth 16) on N GPUs of one box.
One process per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
       tools/gpu_multi_configs.py [--check] [cornell4k] [synthetic]
Rank r renders iterations r+1, r+1+N, ... (spp_split.rank_iterations) into its private buffer; the frame is the
NCCL sum-reduce to rank 0.  Time: CUDA events on the launching stream around render + reduce, max over ranks.
--check: rank 0 afterwards renders ALL the iterations alone and compares the reduced frame with it (the float
summation order differs between the two, SURVEY.md 8(e) caveat; the bar is the recursive-summation bound, iterations * 2^-24
relative per pixel channel of the mean -- pixels whose samples are all equal reach 5e-5 at 5000 spp).
One JSON line per config on rank 0."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as entry
import scenes

m = entry.load_package()
m.lib()
split = __import__("importlib").import_module(entry.PKG + ".spp_split")
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
argv = sys.argv[1:]
check = "--check" in argv
which = [a for a in argv if not a.startswith("--")] or ["cornell4k", "synthetic"]
CONFIGS = {  # name -> (scene text, total iterations of the frame, batch_iters)
    "cornell4k": (lambda: scenes.cornell_text(res=(3840, 2160)), 5000, 4),
    "synthetic": (lambda: scenes.synthetic_text(100000), 512, 16),
}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def allred(x, op):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=op)
    return float(t.item())


tmp = tempfile.mkdtemp()
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
for name in which:
    text, total, batch = CONFIGS[name]
    path = os.path.join(tmp, "%s_%d.txt" % (name, rank))
    open(path, "w").write(text())
    scene = m.Scene.load(path)
    P = scene.pixels
    accum = torch.zeros(P * 3, dtype=torch.float32, device=dev)
    frame = torch.zeros_like(accum)
    r = m.Renderer(scene, device=local, bsdf=m.BSDF_FULL, pipeline=m.PIPELINE_FUSED, batch_iters=batch,
                   stream=stream.cuda_stream, external_accum=accum.data_ptr())
    b, c, s = split.rank_iterations(1, total, rank, world)
    # warm-up: a short frame (graph capture, NCCL channel setup), then reset
    r.render(b, min(c, 4 * batch), s); split.reduce_frame(accum, frame); r.reset(); accum.zero_()
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(stream)
    r.render(b, c, s)
    e1.record(stream)
    split.reduce_frame(accum, frame)
    e2.record(stream)
    barrier()
    ms_render, ms_total = allred(e0.elapsed_time(e1), dist.ReduceOp.MAX), allred(e0.elapsed_time(e2), dist.ReduceOp.MAX)
    ms_reduce = allred(e1.elapsed_time(e2), dist.ReduceOp.MIN)       # the rank that arrives last waits for nobody
    st = r.stats()
    seg = allred(st["segments"], dist.ReduceOp.SUM)
    iters = allred(st["iterations"], dist.ReduceOp.SUM)
    out = {"config": name, "n_gpus": world, "pixels": P, "geoms": len(scene.geoms), "trace_depth": scene.traceDepth,
           "iterations_total": int(iters), "batch_iters": st["batch_iters"], "bvh_nodes": st["bvh_nodes"],
           "ms_frame": ms_total, "ms_render_max_rank": ms_render, "ms_reduce_min_rank": ms_reduce, "reduce_bytes": P * 12,
           "iters_per_s": iters / (ms_total * 1e-3), "segments_per_s": seg / (ms_total * 1e-3), "segments_per_iter": seg / max(1.0, iters)}
    r.close()
    if check and rank == 0:
        ref = torch.zeros(P * 3, dtype=torch.float32, device=dev)
        r1 = m.Renderer(scene, device=local, bsdf=m.BSDF_FULL, pipeline=m.PIPELINE_FUSED, batch_iters=batch,
                        stream=stream.cuda_stream, external_accum=ref.data_ptr())
        torch.cuda.synchronize(dev); t0 = time.perf_counter()
        r1.render(1, total, 1); r1.sync()
        out["one_gpu_same_frame_ms"] = (time.perf_counter() - t0) * 1e3
        r1.close()
        a, bb = frame.double() / total, ref.double() / total
        out["max_abs_diff_of_mean"] = float((a - bb).abs().max().item())
        out["max_rel_diff_of_mean"] = float(((a - bb).abs() / bb.abs().clamp_min(1e-3)).max().item())
        out["frame_mean"] = float(a.mean().item())
        out["check_bar"] = total * 2.0 ** -24                  # recursive float32 summation of `total` samples
        out["check_ok"] = bool(out["max_rel_diff_of_mean"] <= out["check_bar"])
    if rank == 0:
        print(json.dumps(out), flush=True)
    barrier()
if world > 1:
    dist.destroy_process_group()
