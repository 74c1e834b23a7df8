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
    # a dedicated (non-default) stream: the library launches on it, torch's events are recorded on it and
    # NCCL orders against it, so CUDA-event timing sees exactly the work of the step
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    accum = torch.zeros(P * 3, dtype=torch.float32, device=dev)        # private partial sums of this rank
    frame = torch.zeros_like(accum)                                     # reduced frame (meaningful on rank 0)
    host_frame = torch.empty(P * 3, dtype=torch.float32, pin_memory=True)
    r = m.Renderer(scene, device=local, bsdf=m.BSDF_FULL, pipeline=m.PIPELINE_FUSED, batch_iters=args.batch_iters,
                   stream=stream.cuda_stream, external_accum=accum.data_ptr())

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
            split.reduce_frame(accum, frame)
        step_counter[0] += 1

    def e2e_step():
        r.update_scene(scene)                                          # H2D from pinned staging, resets accumulation
        b, c, s = split.rank_iterations(1, spp * world, rank, world)
        r.render(b, c, s)
        if world > 1:
            split.reduce_frame(accum, frame)
            if rank == 0:
                host_frame.copy_(frame, non_blocking=True)
                stream.synchronize()
        else:
            r.download_to(host_frame.data_ptr())                       # D2H into pinned host memory + sync

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
    se0 = r.stats()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    e2e_steps = max(3, args.steps // 2)
    seg_e2e = 0
    for _ in range(e2e_steps):
        e2e_step()
        seg_e2e += r.stats()["segments"]                                # stats reset with every scene upload
    g1.record(stream)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    seg_e2e_all = allsum(seg_e2e)
    h2d = r.stats()["scene_upload_bytes"]
    d2h = P * 3 * 4 if rank == 0 else 0

    # per-iteration drop-in semantic (pathtrace(pbo, frame, iter) + D2H every call), N = 1 only, reported as an extra
    dropin = None
    if world == 1:
        r.reset()
        img = np.empty((P, 3), np.float32)
        hp = host_frame.data_ptr()
        for it in range(1, 4):
            m._check(m.lib().ptb_pathtrace(r._h, None, 0, it, hp))
        t0 = time.perf_counter()
        n_it = 32
        for it in range(4, 4 + n_it):
            m._check(m.lib().ptb_pathtrace(r._h, None, 0, it, hp))
        dt = time.perf_counter() - t0
        sdi = r.stats()
        dropin = {"iters_per_s": n_it / dt, "segments_per_s": sdi["segments"] / sdi["iterations"] * n_it / dt,
                  "d2h_bytes_per_iter": P * 12, "note": "ptb_pathtrace with a host image every iteration (src/pathtrace.cu:389-390 semantic)"}

    # the same device-resident loop with the first-bounce cache on (INSTRUCTION.md:87-89; result-neutral,
    # tests/test_gpu_parity.py::test_first_bounce_cache_is_result_neutral) -- reported beside `value`, not as `value`
    r.close()
    r = m.Renderer(scene, device=local, bsdf=m.BSDF_FULL, pipeline=m.PIPELINE_FUSED, batch_iters=args.batch_iters,
                   stream=stream.cuda_stream, external_accum=accum.data_ptr(), first_bounce_cache=1)
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

    if rank == 0:
        hbm_peak, peak_src, sm_max = load_peaks()
        alg_bytes = BYTES_PER_SEGMENT * seg_all + BYTES_PER_PIXEL_ITER * P * iters_all
        achieved = alg_bytes / (ms * 1e-3) / 1e9 / world             # per GPU
        # per GPU: every batch is `bounce launches` + one advance_kernel, every ptb_render call one set_control_kernel
        batches = iters_all / world / st1["batch_iters"]
        bounce_launches = max(1.0, launches_all / world - batches - args.steps)
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
        line = {
            "metric": METRIC, "value": seg_all / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "reference cornell scene regenerated from tests/scenes.py; no external data",
            "config": {"workload": WORKLOAD, "bsdf": "full", "iterations_per_step_per_gpu": spp, "batch_iters": st1["batch_iters"],
                       "pipeline": "fused bounce kernels + CUDA graph", "parallelism": "spp split x%d, 1 NCCL reduce/frame" % world,
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
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "ptb::bounce_kernel (bounce 0) + ptb::bounce_loop_kernel (bounces 1..)",
                         "algorithmic_bytes": "88 B/segment + 24 B/pixel/iteration (SURVEY.md 8(d))",
                         "bounce_launches_in_timed_region_per_gpu": bounce_launches,
                         "avg_launch_us": ms * 1e3 / max(1.0, bounce_launches),
                         "fp32": {"achieved_tflops": FLOP_PER_SEGMENT * seg_all / (ms * 1e-3) / 1e12 / world,
                                  "peak_tflops": 148 * 128 * 2 * sm_max * 1e6 / 1e12, "flop_per_segment": FLOP_PER_SEGMENT,
                                  "note": "the bounce kernel is instruction-issue bound, not HBM bound: bit-exact parity forbids FFMA in the exact tests (DESIGN.md 4.4)"}},
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                val, kind, cores, sample, secs, _ = cpu_reference_run(scene_path, CPU_ITERS, 8)     # ~10 s of CPU work
                line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": secs}
            except Exception as ex:                                     # the checker is optional for the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
        out.write(json.dumps(line) + "\n")
        out.flush()
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
