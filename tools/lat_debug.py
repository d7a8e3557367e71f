"""Single-archive latency probe: device time (CUDA graph, L2 flushed) and serial stage times for the cfg1 fixture and one cfg2
archive.  NAFGPU_DEBUG_HUF=1 adds the cycle accounting of k_decode_sequences / k_lz_small; NAFGPU_LZ_SMALL=N moves the
one-CTA match stage's limit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nafcodec_b200 as N
fx = open(os.path.join(ROOT, "tests", "golden", "NZ_AAEN01000029.naf"), "rb").read()
ctx = N.Context(0)
for name, data in (("cfg1", fx), ("cfg2", bench.make_workload(1, 5_000_000, 19, 0)[0])):
    ctx.prepare([N.parse_archive(data)]); ctx.sync()
    ctx.time_runs(5, True)
    ms = ctx.time_runs(20, True) / 20
    ctx.run(); ctx.fetch_raw()
    st = ctx.stats()
    print(f"{name}: device {ms * 1e3:.1f} us, {st.n_sequences} matches, {st.lz_rounds} rounds, {st.kernel_launches} launches", file=sys.stderr)
    print({k: round(v, 4) for k, v in ctx.profile_stages() if v > 0.002}, file=sys.stderr)
