"""The toggle table the assignment's analysis asks for (INSTRUCTION.md:78-89, 339-355): stream compaction on / off, material sort
on / off, first-bounce cache at depths 2 / 4 / 8 / 16, open vs closed box, one launch per bounce vs the bounce loop, ordered vs
atomic deposit -- microseconds per iteration by CUDA events, plus the per-stage device time of the staged pipeline
(PtbOptions.timeStages) and the paths alive per depth.   python tools/gpu_toggles.py > profiles/<tag>_toggles.txt"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import scenes
m = g.load_package()
d = tempfile.mkdtemp()


def scene_of(text, name):
    p = os.path.join(d, name + ".txt")
    open(p, "w").write(text)
    return m.Scene.load(p)


def time_us(scene, n, **kw):
    """us per iteration (CUDA events on the library's own stream handle), segments per iteration, alive per depth"""
    stream = torch.cuda.Stream()
    with m.Renderer(scene, device=0, bsdf=kw.pop("bsdf", 1), stream=stream.cuda_stream, **kw) as r:
        r.render(1, n); r.sync()
        best = 1e30
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); r.render(1 + (rep + 1) * n, n); e1.record(stream); r.sync()
            best = min(best, e0.elapsed_time(e1) * 1e3 / n)
        st = r.stats()
    return best, st["segments"] / st["iterations"], [a / st["iterations"] for a in st["alive_per_depth"]]


def stages(scene, n, **kw):
    with m.Renderer(scene, device=0, bsdf=1, time_stages=1, **kw) as r:
        r.render(1, n); r.reset(); r.render(1, n)
        st = r.stats()
    return {k: (v / n, st["stage_launches"][k] // n) for k, v in st["stage_us"].items() if st["stage_launches"][k]}


cornell = scene_of(scenes.cornell_text(), "cornell")
closed = scene_of(scenes.closed_cornell_text(), "closed")
print("# cornell 800x800 depth 8, full BSDF, one B200; us per iteration, CUDA events, best of 3")
print("## fused pipeline (the fast path)")
for tag, sc, kw, n in (("batch 128, bounce loop, ordered deposit (bench configuration)", cornell, dict(batch_iters=128), 256),
                       ("batch 128, bounce loop, atomic deposit", cornell, dict(batch_iters=128, deposit_mode=1), 256),
                       ("batch 128, one launch per bounce", cornell, dict(batch_iters=128, launch_per_bounce=1), 256),
                       ("batch 128, first-bounce cache", cornell, dict(batch_iters=128, first_bounce_cache=1), 256),
                       ("batch 1 (one iteration per wavefront)", cornell, dict(batch_iters=1), 64),
                       ("batch 8", cornell, dict(batch_iters=8), 64),
                       ("closed box, batch 128", closed, dict(batch_iters=128), 256),
                       ("closed box, batch 128, first-bounce cache", closed, dict(batch_iters=128, first_bounce_cache=1), 256)):
    us, seg, alive = time_us(sc, n, **kw)
    print("%-70s %9.1f us/iter  %.3e seg/s  segments/iter %.0f" % (tag, us, seg / us * 1e6, seg))
print("alive per depth, open box  :", " ".join("%.0f" % a for a in time_us(cornell, 16, batch_iters=16)[2]))
print("alive per depth, closed box:", " ".join("%.0f" % a for a in time_us(closed, 16, batch_iters=16)[2]))
print("## first-bounce cache against trace depth (fused, batch 64): us/iter without -> with (bit-identical images)")
for depth in (2, 4, 8, 16):
    sc = scene_of(scenes.cornell_text(depth=depth), "cd%d" % depth)
    a = time_us(sc, 128, batch_iters=64)[0]
    b = time_us(sc, 128, batch_iters=64, first_bounce_cache=1)[0]
    print("depth %2d: %8.1f -> %8.1f us/iter (%.1f %% saved)" % (depth, a, b, 100 * (a - b) / a))
print("## staged pipeline (the reference's stage structure), batch 8")
for tag, sc, kw in (("compaction on", cornell, dict(pipeline=1, compaction=1, batch_iters=8)),
                    ("compaction off (skeleton layout: every slot at every bounce)", cornell, dict(pipeline=1, compaction=0, batch_iters=8)),
                    ("compaction on + material sort", cornell, dict(pipeline=1, compaction=1, sort_materials=1, batch_iters=8)),
                    ("closed box, compaction on", closed, dict(pipeline=1, compaction=1, batch_iters=8)),
                    ("closed box, compaction off", closed, dict(pipeline=1, compaction=0, batch_iters=8))):
    us, seg, _ = time_us(sc, 32, **kw)
    print("%-70s %9.1f us/iter  %.3e seg/s" % (tag, us, seg / us * 1e6))
    st = stages(sc, 16, **kw)
    print("    per-stage us/iter (launches/iter):", "  ".join("%s %.1f (%d)" % (k, v[0], v[1]) for k, v in st.items()))
print("## fused pipeline per-stage device time (launches go out one by one under the timers), batch 32")
st = stages(cornell, 64, batch_iters=32)
print("    ", "  ".join("%s %.1f us/iter" % (k, v[0]) for k, v in st.items()))
if os.environ.get("PTB_TOGGLES_SYNTH", "1") != "0":
    syn = scene_of(scenes.synthetic_text(100000), "synthetic")
    print("# synthetic 100k objects 1920x1080 depth 16 (flat BVH), batch 16")
    us, seg, alive = time_us(syn, 32, batch_iters=16)
    print("%-70s %9.1f us/iter  %.3e seg/s" % ("fused: trace_kernel + bounce_kernel per depth", us, seg / us * 1e6))
    us2, _, _ = time_us(syn, 32, batch_iters=16, first_bounce_cache=1)
    print("%-70s %9.1f us/iter" % ("fused + first-bounce cache", us2))
    st = stages(syn, 16, batch_iters=16)
    print("     per-stage us/iter:", "  ".join("%s %.1f" % (k, v[0]) for k, v in st.items()))
    print("alive per depth:", " ".join("%.0f" % a for a in alive))
    us3, seg3, _ = time_us(syn, 8, pipeline=1, compaction=1, batch_iters=4)
    us4, _, _ = time_us(syn, 8, pipeline=1, compaction=1, sort_materials=1, batch_iters=4)
    print("staged, compaction on: %.1f us/iter; + material sort (7 materials): %.1f us/iter" % (us3, us4))
    st = stages(syn, 8, pipeline=1, compaction=1, sort_materials=1, batch_iters=4)
    print("     per-stage us/iter:", "  ".join("%s %.1f" % (k, v[0]) for k, v in st.items()))
