"""Do the jobs of several contexts overlap on one device?  The 256-archive batch as 1, 2 and 4 jobs launched back to back.
python tools/overlap_probe.py [archives] [steps]   (NAFGPU_NO_PRIORITIES=1: all streams at one priority)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nafcodec_b200 as N
n_arch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
arcs = bench.make_workload(8, 5_000_000, 19, 0)
batch = [N.parse_archive(arcs[i % len(arcs)]) for i in range(n_arch)]
for parts in (1, 2, 4):
    ctxs = [N.Context(0) for _ in range(parts)]
    per = n_arch // parts
    for k, c in enumerate(ctxs):
        c.prepare(batch[k * per:(k + 1) * per]); c.sync()
    for _ in range(3):
        for c in ctxs: c.run()
        for c in ctxs: c.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        for c in ctxs: c.run()
        for c in ctxs: c.sync()
    dt = (time.perf_counter() - t0) / steps
    print(f"{parts} job(s) of {per}: {dt * 1e3:.3f} ms per step, {n_arch * 5e6 / dt / 1e9:.1f} GB/s ASCII (wall clock, kernels enqueued per step)")
    for c in ctxs: c.close()
