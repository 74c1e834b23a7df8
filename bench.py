#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: path segments/s (and iterations/s) on
cornell 800x800, traceDepth 8, through the B200-native wavefront path tracer.

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, libptb200.so)
    python bench.py --impl reference [...]                        the reference's CPU path on host cores

A "step" is one output frame: SPP_PER_STEP iterations of the per-bounce loop over all 640 000 pixels
(plus, for N > 1, the single NCCL sum-reduce of the accumulation buffer to rank 0).
Rank r of N renders iterations r+1, r+1+N, ... so per-GPU work is fixed as N grows ("weak").

`value`  : whole-job path segments/s with everything resident in HBM, timed with CUDA events on the
           launching stream, max over ranks.
`e2e`    : the same metric through the host-buffer API a user calls -- per step the scene is re-uploaded
           from pinned host memory (ptb_update_scene), the frame is traced (ptb_render) and the
           accumulated image is copied to a pinned host buffer (ptb_download); copies are inside the
           timed region.
Only this file's cpu_baseline leg and --impl reference execute anything under oracle/ (as the thing
being timed on the CPU); the product path never does.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "scenes/cornell.txt 800x800 traceDepth 8 (BASELINE.json configs[2] scene/size; configs[0] on the CPU arm)"
METRIC = "path_segments_per_s"
UNIT = "segments/s"
SPP_PER_STEP = 256            # iterations per step (frame)
BATCH_ITERS = 128             # iterations per wavefront: 81.92 M paths, 2 x 3.28 GB of path state (>> 126 MB L2); 32: 109.3, 64: 107.2, 128: 106.1, 256: 105.5 us/iteration
CPU_ITERS = 16                # BASELINE.json configs[0]: 16 iterations through the reference's host functions
FLOP_PER_SEGMENT = 700        # ~7 ray-primitive tests + scatter, FMA counted as 2
BYTES_PER_SEGMENT = 88        # 44 B read + 44 B write of path state
BYTES_PER_PIXEL_ITER = 24     # accumulation read + write


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def write_scene(tmpdir):
    import scenes
    path = os.path.join(tmpdir, "cornell.txt")
    with open(path, "w") as f:
        f.write(scenes.cornell_text())
    return path


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(scene_path, iters_per_sample, repeats, nthreads=0):
    """Times the reference's CPU implementation of the path: the reference's own __host__ __device__
    intersect / scatter functions compiled for the host (oracle/_ref, kind "reference"), else the
    plain-C port.  Returns (segments/s, kind, cores, sample description, seconds)."""
    import orclib
    orc = orclib.load("reference") or orclib.load("port")
    if orc is None:
        raise RuntimeError("neither oracle/_ref/libptref.so nor oracle/libptoracle.so is built")
    scene = orc.load_scene(scene_path)
    # the thread count is passed explicitly (omp_set_num_threads in the oracle): torchrun exports OMP_NUM_THREADS=1 to
    # its workers, which would silently turn the all-cores baseline of an N > 1 launch into a one-core run
    if not nthreads:
        try:
            nthreads = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            nthreads = os.cpu_count() or 1
    cores = nthreads
    best = None
    total_seg, total_t = 0, 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, _, seg = orc.render(scene, orclib.MODE_FULL, 1, iters_per_sample, nthreads=nthreads)
        dt = time.perf_counter() - t0
        total_seg += seg
        total_t += dt
        best = seg / dt if best is None else max(best, seg / dt)
    sample = "%d x %d iterations of cornell 800x800 depth 8 (full BSDF), OpenMP over pixel rows, parity build -O2 -ffp-contract=off" % (repeats, iters_per_sample)
    return total_seg / total_t, orc.kind, cores, sample, total_t, seg / iters_per_sample


def run_reference_arm(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    tmp = tempfile.mkdtemp()
    scene_path = write_scene(tmp)
    iters = CPU_ITERS                           # one step = BASELINE.json configs[0]: 16 iterations (~39.6 M segments, ~1.2 s on 16 cores)
    cpu_reference_run(scene_path, 2, max(1, args.warmup))      # warm-up steps
    t0 = time.perf_counter()
    val, kind, cores, sample, secs, seg_per_iter = cpu_reference_run(scene_path, iters, args.steps)
    ms_per_step = secs / args.steps * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "reference cornell scene regenerated from tests/scenes.py; no external data",
        "config": {"workload": WORKLOAD, "bsdf": "full", "iterations_per_step": iters, "host": "CPU, all cores"},
        "iters_per_s": val / seg_per_iter,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()
    return 0


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries loaded later write there too (NCCL prints its
    version banner to fd 1 under torchrun), so fd 1 is pointed at stderr for the whole run and the JSON line
    goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def skeleton_gpu_baseline(scene_path, iters=1000, warm=50):
    """The reference skeleton's own CUDA loop (unmodified src/pathtrace.cu rebuilt for sm_100 behind oracle/skeleton_main.cpp,
    built by oracle/Makefile where /root/reference exists; the binary travels in oracle/_ref/): the denominator of
    BASELINE.json's ">= 10x" target.  Runs on this box's GPU OUTSIDE every timed region.  None when the binary is absent."""
    exe = os.path.join(ROOT, "oracle", "_ref", "skeleton_bench")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, scene_path, str(iters), str(warm)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=300)
        for line in r.stdout.splitlines():
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        return None
    return None


def roofline_of(seg, iters, pixels, ms, hbm_peak, sm_max):
    """algorithmic bytes = 88 B per path segment + 24 B per pixel per iteration; flops = 700 per segment."""
    alg = BYTES_PER_SEGMENT * seg + BYTES_PER_PIXEL_ITER * pixels * iters
    gbs = alg / (ms * 1e-3) / 1e9
    tf = FLOP_PER_SEGMENT * seg / (ms * 1e-3) / 1e12
    return {"algorithmic_gbs": gbs, "hbm_frac": gbs / hbm_peak, "fp32_tflops": tf, "fp32_frac": tf / (148 * 128 * 2 * sm_max * 1e6 / 1e12)}


def timed_render(torch, stream, r, first, count, reps=1):
    """count iterations starting at `first`, timed with CUDA events on the launching stream; best of reps -> (ms, stats delta)"""
    best = None
    for _ in range(reps):
        r.sync()
        s0 = r.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r.render(first, count)
        e1.record(stream)
        r.sync()
        ms = e0.elapsed_time(e1)
        s1 = r.stats()
        if best is None or ms < best[0]:
            best = (ms, s1["segments"] - s0["segments"], s1["iterations"] - s0["iterations"])
        first += count
    return best


def extra_configs(m, torch, stream, local, tmp, hbm_peak, sm_max):
    """BASELINE.json configs[1], [3] (one GPU, short) and [4] beside the headline line: us/iteration, segments/s, roofline fractions.
    Single-GPU only (world == 1); everything is timed with CUDA events after a warm-up run of the same length."""
    import scenes
    out = {}
    specs = [
        ("configs[1] sphere.txt 800x800 depth 8", scenes.sphere_text(), dict(batch_iters=128), 1024),
        ("configs[3] cornell.txt 3840x2160 depth 8 (one GPU, 64 of the 5000 spp)", scenes.cornell_text(res=(3840, 2160)), dict(), 64),
        ("configs[4] synthetic 100k SPHERE/CUBE 1920x1080 depth 16 (flat BVH)", scenes.synthetic_text(100000), dict(), 64),
    ]
    for name, text, kw, n in specs:
        path = os.path.join(tmp, "extra.txt")
        with open(path, "w") as f:
            f.write(text)
        sc = m.Scene.load(path)
        with m.Renderer(sc, device=local, bsdf=m.BSDF_FULL, stream=stream.cuda_stream, **kw) as r:
            r.render(1, n)                                              # warm-up (graph capture, first-touch)
            ms, seg, its = timed_render(torch, stream, r, 1 + n, n, reps=2)
            st = r.stats()
        d = {"us_per_iteration": ms * 1e3 / its, "segments_per_s": seg / (ms * 1e-3), "iters_per_s": its / (ms * 1e-3),
             "segments_per_iter": seg / its, "iterations_timed": its, "batch_iters": st["batch_iters"], "bvh_nodes": st["bvh_nodes"]}
        d.update(roofline_of(seg, its, sc.pixels, ms, hbm_peak, sm_max))
        out[name] = d
    return out


def golden_image_error(m, torch, stream, local, scene):
    """north_star: the converged cornell image within relative RMSE 1e-2 of img/REFERENCE_cornell.5000samp.png (committed as
    tests/golden/reference_cornell_5000samp.npz).  Same metric as tests/test_gpu_scale.py: 16x16 block means, golden-diffuse mode."""
    import numpy as np
    gp = os.path.join(ROOT, "tests", "golden", "reference_cornell_5000samp.npz")
    if not os.path.exists(gp):
        return None
    gold = np.load(gp)["rgb8"].astype(np.float64) / 255.0
    with m.Renderer(scene, device=local, bsdf=m.BSDF_GOLDEN_DIFFUSE, batch_iters=64, stream=stream.cuda_stream) as r:
        r.render(1, 5000)
        ours = r.resolve_rgb8(5000).astype(np.float64) / 255.0
    rms = np.sqrt((gold ** 2).mean())
    blk = lambda a: a.reshape(50, 16, 50, 16, 3).mean((1, 3))
    return {"rel_rmse_16x16_blocks": float(np.sqrt(((blk(ours) - blk(gold)) ** 2).mean()) / rms),
            "rel_rmse_per_pixel": float(np.sqrt(((ours - gold) ** 2).mean()) / rms),
            "mean_ratio_rgb": [float(x) for x in ours.mean((0, 1)) / gold.mean((0, 1))], "spp": 5000, "bound": 1e-2,
            "note": "golden-diffuse mode vs img/REFERENCE_cornell.5000samp.png; per-pixel floor between two 5000-spp renders is ~0.086"}


def main():
    out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=SPP_PER_STEP)
    ap.add_argument("--batch-iters", type=int, default=BATCH_ITERS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline line only: no extra configs, skeleton, golden image, strong-scaling leg")
    ap.add_argument("--strong-spp", type=int, default=5000, help="samples per pixel of the configs[3] strong-scaling frame")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args, out)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    m = entry.load_package()
    m.lib()                                           # fail loudly if the CUDA extension is missing
    split = __import__("importlib").import_module(entry.PKG + ".spp_split")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    tmp = tempfile.mkdtemp()
    scene_path = write_scene(tmp)
    scene = m.Scene.load(scene_path)
    P = scene.pixels
    spp = args.spp_per_step
    # a dedicated (non-default) stream: the library launches on it (kernels AND its own NCCL reduce), torch's events are recorded
    # on it, so CUDA-event timing sees exactly the work of the step
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    host_frame = torch.empty(P * 3, dtype=torch.float32, pin_memory=True)

    def make_renderer(sc, **kw):
        r = m.Renderer(sc, device=local, bsdf=m.BSDF_FULL, pipeline=m.PIPELINE_FUSED, stream=stream.cuda_stream, **kw)
        if world > 1:
            # the frame collective is the LIBRARY's: ncclCommInitRank + ncclReduce inside libptb200.so (include/ptb200.h
            # ptb_comm_*); torch.distributed only carries the 128-byte ncclUniqueId, the barriers and the max over ranks
            split.init_library_comm(r, rank, world, m)
        return r

    r = make_renderer(scene, batch_iters=args.batch_iters)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    step_counter = [0]

    def device_step():
        first = 1 + step_counter[0] * spp * world
        b, c, s = split.rank_iterations(first, spp * world, rank, world)
        r.render(b, c, s)
        if world > 1:
            split.reduce_frame_library(r, 0)                           # ONE ncclReduce(sum) of 3*W*H floats per frame
        step_counter[0] += 1

    def e2e_step():
        r.update_scene(scene)                                          # H2D from pinned staging, resets accumulation
        b, c, s = split.rank_iterations(1, spp * world, rank, world)
        r.render(b, c, s)
        if world > 1:
            split.reduce_frame_library(r, 0)
        if rank == 0:
            r.frame_download_to(host_frame.data_ptr())                 # D2H of the (reduced) frame into pinned host memory + sync
        else:
            r.sync()

    # ------------------------------------------------------------------ device-resident timing (`value`)
    for _ in range(args.warmup):
        device_step()
    barrier()
    st0 = r.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        device_step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    st1 = r.stats()
    seg = st1["segments"] - st0["segments"]
    iters = st1["iterations"] - st0["iterations"]
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms = allmax(ms)
    seg_all, iters_all, launches_all = allsum(seg), allsum(iters), allsum(launches)

    # ------------------------------------------------------------------ end-to-end timing (`e2e`)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    seg_e2e = 0
    for _ in range(e2e_steps):
        e2e_step()
        seg_e2e += r.stats()["segments"]                                # stats reset with every scene upload
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    seg_e2e_all = allsum(seg_e2e)
    h2d = r.stats()["scene_upload_bytes"]
    d2h = P * 3 * 4 if rank == 0 else 0

    # ------------------------------------------------------------------ the reference's own call: pathtrace(pbo, frame, iter), N = 1
    # strict = the host image is valid on return every iteration (src/pathtrace.cu:389-390); lazy = PtbOptions.lazyImage: the copy
    # of iteration i overlaps iteration i + 1 on a copy stream, valid after ptb_sync.  One iteration per wavefront in both.
    dropin = None
    if world == 1:
        dropin = {}
        hp = host_frame.data_ptr()
        # *_lookahead = PtbOptions.lookahead: the call for iteration i traces the wavefront i .. i + 63 at the batched rate and every call adds
        # only its own iteration to the image (bit-identical: tests/test_gpu_parity.py::test_pathtrace_lookahead_is_bit_identical)
        for mode, lazy, la_batch in (("strict", 0, 0), ("lazy_image", 1, 0), ("strict_lookahead", 0, 64), ("lazy_image_lookahead", 1, 64)):
            with m.Renderer(scene, device=local, bsdf=m.BSDF_FULL, batch_iters=la_batch or 1, lazy_image=lazy, lookahead=1 if la_batch else 0) as rd:
                for it in range(1, 9):
                    m._check(m.lib().ptb_pathtrace(rd._h, None, 0, it, hp))
                rd.sync()
                n_it = 256
                t0 = time.perf_counter()
                for it in range(9, 9 + n_it):
                    m._check(m.lib().ptb_pathtrace(rd._h, None, 0, it, hp))
                rd.sync()
                dt = time.perf_counter() - t0
                sdi = rd.stats()
            dropin[mode] = {"iters_per_s": n_it / dt, "us_per_call": dt / n_it * 1e6,
                            "segments_per_s": sdi["segments"] / sdi["iterations"] * n_it / dt, "d2h_bytes_per_iter": P * 12}
            if la_batch:
                dropin[mode]["iterations_per_window"] = la_batch
        # the PCIe floor of the strict semantic: 7.68 MB device->host per call at the measured pinned D2H rate
        src = torch.empty(P * 3, dtype=torch.float32, device=dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host_frame.copy_(src, non_blocking=True); stream.synchronize()
        c0.record(stream)
        for _ in range(20):
            host_frame.copy_(src, non_blocking=True)
        c1.record(stream); stream.synchronize()
        d2h_ms = c0.elapsed_time(c1) / 20
        dropin["pcie_floor"] = {"d2h_gbs_pinned": P * 12 / (d2h_ms * 1e-3) / 1e9, "us_per_7.68MB_copy": d2h_ms * 1e3,
                                "note": "ptb_pathtrace with a host image every iteration cannot run faster than this copy (src/pathtrace.cu:389-390 semantic)"}

    # the same device-resident loop with the first-bounce cache on (INSTRUCTION.md:87-89; result-neutral,
    # tests/test_gpu_parity.py::test_first_bounce_cache_is_result_neutral) -- reported beside `value`, not as `value`
    r.close()
    r = make_renderer(scene, batch_iters=args.batch_iters, first_bounce_cache=1)
    for _ in range(args.warmup):
        device_step()
    barrier()
    sc0 = r.stats()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for _ in range(args.steps):
        device_step()
    c1.record(stream)
    barrier()
    ms_fbc = allmax(c0.elapsed_time(c1))
    sc1 = r.stats()
    seg_fbc = allsum(sc1["segments"] - sc0["segments"])
    iters_fbc = allsum(sc1["iterations"] - sc0["iterations"])
    r.close()

    # ------------------------------------------------------------------ strong scaling: BASELINE.json configs[3]
    # cornell.txt at 3840 x 2160, depth 8, a FIXED number of samples per pixel split over the N GPUs, one NCCL reduce of the
    # 99.5 MB accumulation buffer per frame (inside the library), the frame copied to the host on rank 0.
    strong = None
    if not args.no_extras:
        import scenes
        p4 = os.path.join(tmp, "cornell4k.txt")
        with open(p4, "w") as f:
            f.write(scenes.cornell_text(res=(3840, 2160)))
        s4 = m.Scene.load(p4)
        r4 = make_renderer(s4)
        host4 = torch.empty(s4.pixels * 3, dtype=torch.float32, pin_memory=True) if rank == 0 else None
        b, c, s = split.rank_iterations(1, args.strong_spp, rank, world)
        r4.render(b, min(c, 32), s)                                    # warm-up: graph capture, NCCL connection set-up
        if world > 1:
            split.reduce_frame_library(r4, 0)
        r4.reset()
        barrier()
        q0, q1, q2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        q0.record(stream)
        r4.render(b, c, s)
        q1.record(stream)
        if world > 1:
            split.reduce_frame_library(r4, 0)
        q2.record(stream)
        if rank == 0:
            r4.frame_download_to(host4.data_ptr())
        barrier()
        ms_frame = allmax(q0.elapsed_time(q2))
        ms_reduce = q1.elapsed_time(q2)
        st4 = r4.stats()
        seg4 = allsum(st4["segments"])
        if rank == 0:
            strong = {"workload": "scenes/cornell.txt 3840x2160 depth 8, %d spp per frame split over %d GPU(s) (BASELINE.json configs[3])" % (args.strong_spp, world),
                      "scaling": "strong", "ms_per_frame": ms_frame, "segments_per_s": seg4 / (ms_frame * 1e-3),
                      "reduce_ms_rank0": ms_reduce if world > 1 else 0.0, "reduce_bytes": s4.pixels * 12 if world > 1 else 0,
                      "iterations_per_gpu": c, "batch_iters": st4["batch_iters"],
                      "frame_mean": float(host4.mean().item()) / args.strong_spp}
        r4.close()

    if rank == 0:
        hbm_peak, peak_src, sm_max = load_peaks()
        alg_bytes = BYTES_PER_SEGMENT * seg_all + BYTES_PER_PIXEL_ITER * P * iters_all
        achieved = alg_bytes / (ms * 1e-3) / 1e9 / world             # per GPU
        # launches of the timed region per GPU, from the library's own counter: per wavefront bounce 0 + the bounce loop + the ordered-deposit
        # reduce + advance, per ptb_render call one set_control
        batches = iters_all / world / st1["batch_iters"]
        bounce_launches = max(1.0, 2.0 * batches)
        traffic = None
        import glob
        profs = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_bounce_traffic.json")))    # newest ncu --set full capture
        traffic_src = None
        if profs:
            try:
                traffic = json.load(open(profs[-1])).get("dram_bytes_per_launch")
                traffic_src = os.path.relpath(profs[-1], ROOT)
            except Exception:
                traffic = None
        # issue-slot view of the same kernels, from the newest committed ncu --set full summary (recorded, not measured live: ncu replays kernels)
        issue = None
        try:
            import csv
            csvs = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_bounce_ncu.csv")))
            rows = list(csv.reader(open(csvs[-1])))
            col = lambda key: [i for i, h in enumerate(rows[0]) if h.startswith(key)][0]
            iB, iL, iT = col("smsp__issue_active.avg.pct"), col("smsp__thread_inst_executed_per_inst_executed"), col("gpu__time_duration")
            issue = {"source": os.path.relpath(csvs[-1], ROOT),
                     "kernels": [{"kernel": r[0].split("(")[0].replace("void ", ""), "issue_slots_busy_pct": float(r[iB]), "lanes_per_instruction": float(r[iL]),
                                  "ms_under_ncu": float(r[iT])} for r in rows[1:]]}
        except Exception:
            issue = None
        line = {
            "metric": METRIC, "value": seg_all / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "reference cornell scene regenerated from tests/scenes.py; no external data",
            "config": {"workload": WORKLOAD, "bsdf": "full", "iterations_per_step_per_gpu": spp, "batch_iters": st1["batch_iters"],
                       "deposit": "ordered (bit-identical to one iteration per wavefront; tests/test_gpu_parity.py::test_bench_configuration_against_the_oracle)",
                       "pipeline": "fused bounce kernels + CUDA graph", "options": "library defaults (leaveTest 1, hitQueue AUTO: on at this wavefront size)",
                       "parallelism": "spp split x%d, 1 ncclReduce/frame inside libptb200.so" % world,
                       "l2": "inputs larger than L2: %.0f MB of path state per wavefront vs 126 MB L2" % (st1["path_state_bytes"] / 2 / 1e6)},
            "iters_per_s": iters_all / (ms * 1e-3),
            "segments_per_iter": seg_all / iters_all,
            "alive_per_depth_per_iter": [a / st1["iterations"] for a in st1["alive_per_depth"]],
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "e2e": {"value": seg_e2e_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps},
            "e2e_per_iteration_dropin": dropin,
            "first_bounce_cache": {"value": seg_fbc / (ms_fbc * 1e-3), "unit": UNIT, "iters_per_s": iters_fbc / (ms_fbc * 1e-3),
                                   "note": "PtbOptions.firstBounceCache = 1: bounce-0 hit records reused across iterations (bit-identical images)"},
            "strong_scaling": strong,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "ptb::bounce0_shared_kernel (bounce 0) + ptb::bounce_loop_kernel (bounces 1..)",
                         "algorithmic_bytes": "88 B/segment + 24 B/pixel/iteration (SURVEY.md 8(d))",
                         "bounce_launches_in_timed_region_per_gpu": bounce_launches,
                         "avg_launch_us": ms * 1e3 / max(1.0, bounce_launches),
                         "issue": issue,
                         "fp32": {"achieved_tflops": FLOP_PER_SEGMENT * seg_all / (ms * 1e-3) / 1e12 / world,
                                  "peak_tflops": 148 * 128 * 2 * sm_max * 1e6 / 1e12, "flop_per_segment": FLOP_PER_SEGMENT,
                                  "note": "the bounce kernel is instruction-issue bound, not HBM bound: bit-exact parity forbids FFMA in the exact tests (DESIGN.md 4.4)"}},
        }
        if world == 1 and not args.no_extras:
            # beside the headline: the other BASELINE.json configs on one GPU, the reference skeleton's own CUDA loop on this
            # GPU (the ">= 10x" denominator) and the converged-image error -- all outside the timed regions above
            try:
                line["configs"] = extra_configs(m, torch, stream, local, tmp, hbm_peak, sm_max)
            except Exception as ex:
                line["configs"] = {"error": str(ex)}
            try:
                line["golden_image"] = golden_image_error(m, torch, stream, local, scene)
            except Exception as ex:
                line["golden_image"] = {"error": str(ex)}
            sk = skeleton_gpu_baseline(scene_path)
            if sk:
                ours_it, ours_seg = line["iters_per_s"], line["value"]
                line["skeleton_gpu_baseline"] = {
                    "impl": "unmodified reference src/pathtrace.cu rebuilt for sm_100 (oracle/_ref/skeleton_bench), one bounce per iteration, D2H of the image every iteration",
                    "iters_per_s": sk["iters_per_s"], "segments_per_s": sk["segments_per_s"],
                    "ours_over_skeleton": {
                        "batched_api": {"segments_per_s": ours_seg / sk["segments_per_s"], "iters_per_s": ours_it / sk["iters_per_s"]},
                        "pathtrace_call_strict": {"segments_per_s": dropin["strict"]["segments_per_s"] / sk["segments_per_s"],
                                                  "iters_per_s": dropin["strict"]["iters_per_s"] / sk["iters_per_s"]},
                        "pathtrace_call_lazy_image": {"segments_per_s": dropin["lazy_image"]["segments_per_s"] / sk["segments_per_s"],
                                                      "iters_per_s": dropin["lazy_image"]["iters_per_s"] / sk["iters_per_s"]},
                        "pathtrace_call_strict_lookahead": {"segments_per_s": dropin["strict_lookahead"]["segments_per_s"] / sk["segments_per_s"],
                                                            "iters_per_s": dropin["strict_lookahead"]["iters_per_s"] / sk["iters_per_s"]},
                        "pathtrace_call_lazy_image_lookahead": {"segments_per_s": dropin["lazy_image_lookahead"]["segments_per_s"] / sk["segments_per_s"],
                                                                "iters_per_s": dropin["lazy_image_lookahead"]["iters_per_s"] / sk["iters_per_s"]}},
                    "target": "BASELINE.json north_star: >= 10x the skeleton's own CUDA loop on cornell 800x800 depth 8"}
        if not args.no_cpu_baseline and world == 1:
            try:
                val, kind, cores, sample, secs, _ = cpu_reference_run(scene_path, CPU_ITERS, 8)     # ~10 s of CPU work
                line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": secs}
            except Exception as ex:                                     # the checker is optional for the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
