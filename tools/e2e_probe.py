"""Where the end-to-end (host in -> host out) time of the bench workload goes: raw PCIe rates, lane sweep, per-call split."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import nafcodec_b200 as N
from nafcodec_b200 import _ffi
import bench

lib = _ffi.default_library()
uniq = bench.make_workload(8, 5_000_000, 19, 0)
import ctypes as C
pinned = []
for a in uniq:
    p = lib.dll.nafgpu_host_alloc(len(a)); C.memmove(p, a, len(a)); pinned.append((p, len(a)))
archives = []
for i in range(64):
    p, n = pinned[i % 8]
    arc = _ffi.Archive(); assert lib.dll.nafgpu_parse_archive(p, n, C.byref(arc)) == 0
    archives.append(arc)
want = _ffi.WANT_ALL
ascii_bytes = sum(int(a.header.number_of_sequences) * 0 + int(a.sections[4].original_size) if hasattr(a, "sections") else 0 for a in archives) or 64 * 5_000_000

# raw PCIe
dev = torch.device("cuda:0")
hp = torch.empty(320 << 20, dtype=torch.uint8).pin_memory()
hq = torch.empty(80 << 20, dtype=torch.uint8).pin_memory()
dp = torch.empty(320 << 20, dtype=torch.uint8, device=dev)
dq = torch.empty(80 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n
t = timed(lambda: hp.copy_(dp, non_blocking=True)); print(f"D2H 320 MiB alone: {t*1e3:.2f} ms, {hp.numel()/t/1e9:.1f} GB/s")
t = timed(lambda: dq.copy_(hq, non_blocking=True)); print(f"H2D 80 MiB alone: {t*1e3:.2f} ms, {hq.numel()/t/1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): hp.copy_(dp, non_blocking=True)
    with torch.cuda.stream(s2): dq.copy_(hq, non_blocking=True)
t = timed(both); print(f"both concurrently: {t*1e3:.2f} ms")

# per-call split on one context, 16 archives
ctx = N.Context(0, lib)
sub = archives[:16]
for _ in range(3):
    ctx.prepare(sub, want); ctx.run(); ctx.fetch()
tp = tr = tf = 0.0
for _ in range(10):
    t0 = time.perf_counter(); ctx.prepare(sub, want); ctx.sync(); t1 = time.perf_counter()
    ctx.run(); ctx.sync(); t2 = time.perf_counter(); ctx.fetch(); t3 = time.perf_counter()
    tp += t1 - t0; tr += t2 - t1; tf += t3 - t2
print(f"16 archives, one context: prepare(+H2D) {tp*100:.2f} ms, run {tr*100:.2f} ms, fetch(D2H) {tf*100:.2f} ms")

for cnt in (16, 64):
    sub = archives[:cnt]
    arr = (_ffi.Archive * cnt)(*sub); res = (_ffi.Result * cnt)()
    tp = tr = tf = 0.0
    for it in range(13):
        t0 = time.perf_counter(); assert lib.dll.nafgpu_job_prepare(ctx._ctx, arr, cnt, want) == 0; lib.dll.nafgpu_job_sync(ctx._ctx); t1 = time.perf_counter()
        assert lib.dll.nafgpu_job_run(ctx._ctx) == 0; lib.dll.nafgpu_job_sync(ctx._ctx); t2 = time.perf_counter()
        assert lib.dll.nafgpu_job_fetch(ctx._ctx, res, cnt) == 0; t3 = time.perf_counter()
        if it >= 3: tp += t1 - t0; tr += t2 - t1; tf += t3 - t2
    st = ctx.stats()
    print(f"{cnt} archives raw C calls: prepare+H2D {tp*100:.2f} ms, run {tr*100:.2f} ms, fetch {tf*100:.2f} ms ({st.d2h_bytes/(tf/10)/1e9:.1f} GB/s D2H), h2d {st.h2d_bytes} d2h {st.d2h_bytes}")
from concurrent.futures import ThreadPoolExecutor
def stream_mode(lanes, steps=10):
    """every lane runs its share of every step without a per-step barrier"""
    ctxs = [N.Context(0, lib) for _ in range(lanes)]
    bounds = [(64 * k) // lanes for k in range(lanes + 1)]
    def work(k, nsteps):
        lo, hi = bounds[k], bounds[k + 1]
        cnt = hi - lo
        arr = (_ffi.Archive * cnt)(*archives[lo:hi]); res = (_ffi.Result * cnt)()
        for _ in range(nsteps):
            rc = lib.dll.nafgpu_decode_batch(ctxs[k]._ctx, arr, cnt, want, res); assert rc == 0
    with ThreadPoolExecutor(lanes) as pool:
        list(pool.map(lambda k: work(k, 3), range(lanes)))
        t0 = time.perf_counter()
        list(pool.map(lambda k: work(k, steps), range(lanes)))
        dt = (time.perf_counter() - t0) / steps
    print(f"stream mode lanes {lanes}: {dt*1e3:.2f} ms/step, {64*5e6/dt/1e9:.1f} GB/s")
    for c in ctxs: c.close()
import threading
def interference():
    """thread A: D2H of a finished 16-archive job in a loop; thread B: prepare-only / run-only / nothing"""
    ca, cb = N.Context(0, lib), N.Context(0, lib)
    cnt = 16
    arr = (_ffi.Archive * cnt)(*archives[:cnt]); resa = (_ffi.Result * cnt)(); resb = (_ffi.Result * cnt)()
    for c in (ca, cb):
        assert lib.dll.nafgpu_job_prepare(c._ctx, arr, cnt, want) == 0
        assert lib.dll.nafgpu_job_run(c._ctx) == 0
        lib.dll.nafgpu_job_sync(c._ctx)
    for mode in ("idle", "run", "prepare", "prepare+run", "fetch"):
        stop = [False]
        def b():
            while not stop[0]:
                if mode in ("prepare", "prepare+run"): assert lib.dll.nafgpu_job_prepare(cb._ctx, arr, cnt, want) == 0
                if mode in ("run", "prepare+run"): assert lib.dll.nafgpu_job_run(cb._ctx) == 0
                if mode == "fetch": assert lib.dll.nafgpu_job_fetch(cb._ctx, resb, cnt) == 0
                lib.dll.nafgpu_job_sync(cb._ctx)
                if mode == "idle": time.sleep(0.001)
        th = threading.Thread(target=b); th.start()
        time.sleep(0.05)
        t0 = time.perf_counter()
        for _ in range(40): assert lib.dll.nafgpu_job_fetch(ca._ctx, resa, cnt) == 0
        dt = (time.perf_counter() - t0) / 40
        stop[0] = True; th.join()
        if mode in ("prepare", "prepare+run"):
            assert lib.dll.nafgpu_job_run(cb._ctx) == 0; lib.dll.nafgpu_job_sync(cb._ctx)
        print(f"D2H of 80 MB while the other context does {mode}: {dt*1e3:.2f} ms ({80.0/dt/1e3:.1f} GB/s)")
interference()
def reverse_interference():
    """thread A: prepare / run of a 16-archive job, timed; thread B: D2H loop on another context"""
    ca, cb = N.Context(0, lib), N.Context(0, lib)
    cnt = 16
    arr = (_ffi.Archive * cnt)(*archives[:cnt]); resb = (_ffi.Result * cnt)()
    for c in (ca, cb):
        assert lib.dll.nafgpu_job_prepare(c._ctx, arr, cnt, want) == 0
        assert lib.dll.nafgpu_job_run(c._ctx) == 0
        lib.dll.nafgpu_job_sync(c._ctx)
    for other in ("idle", "fetch"):
        stop = [False]
        def b():
            while not stop[0]:
                if other == "fetch": assert lib.dll.nafgpu_job_fetch(cb._ctx, resb, cnt) == 0
                else: time.sleep(0.001)
        th = threading.Thread(target=b); th.start()
        time.sleep(0.05)
        tp = tr = 0.0
        for _ in range(40):
            t0 = time.perf_counter(); assert lib.dll.nafgpu_job_prepare(ca._ctx, arr, cnt, want) == 0; lib.dll.nafgpu_job_sync(ca._ctx); t1 = time.perf_counter()
            assert lib.dll.nafgpu_job_run(ca._ctx) == 0; lib.dll.nafgpu_job_sync(ca._ctx); t2 = time.perf_counter()
            tp += t1 - t0; tr += t2 - t1
        stop[0] = True; th.join()
        print(f"while the other context does {other}: prepare+H2D {tp/40*1e3:.2f} ms, run {tr/40*1e3:.2f} ms")
reverse_interference()
for lanes in (1, 2, 3, 4):
    stream_mode(lanes)
for lanes in (4,):
    pipe = N.Pipeline(0, lanes, lib)
    def consume(i, r): pass
    for _ in range(3): pipe.decode(archives, want, consume)
    t0 = time.perf_counter()
    for _ in range(10): pipe.decode(archives, want, consume)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"lanes {lanes}: {dt*1e3:.2f} ms/step, {64*5e6/dt/1e9:.1f} GB/s")
    pipe.close()
