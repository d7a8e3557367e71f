"""Phase times of nafgpu_job_prepare on the cfg4 shape (NAFGPU_DEBUG_PREP=1 prints them).  python tools/prep_probe.py [reads]
NAFGPU_EMUL=1: on the CPU emulator build (host phases only); NAFGPU_WALK_SPLIT=0: long sections walked in one piece."""
import os, sys, time
os.environ["NAFGPU_DEBUG_PREP"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as K
import nafcodec_b200 as N
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
data = K.cfg4_fastq(reads)
lib = None
if os.environ.get("NAFGPU_EMUL"):          # host-side phases only: the CPU emulator build of the same sources (tests/emul)
    import _harness
    lib = _harness.emul_library()
ctx = N.Context(0, lib)
arc = N.parse_archive(data, lib)
for i in range(4):
    t = time.perf_counter(); ctx.prepare([arc]); t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
    print(f"prepare {1e3 * (t1 - t):.2f} ms, sync {1e3 * (t2 - t1):.2f} ms", flush=True)
print("cores", os.cpu_count())
