"""LZ stage probe: pending matches per dependency round + stage times for a cfg3-shaped archive.  python tools/lz_probe.py [Mbp]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as K
import nafcodec_b200 as N
fastq = len(sys.argv) > 1 and sys.argv[1] == "fastq"          # python tools/lz_probe.py fastq [reads] [quality 0/1]
mbp = float(sys.argv[2 if fastq else 1]) if len(sys.argv) > (2 if fastq else 1) else 20
p = os.path.join(ROOT, "bench_cache", "cfg3_n250000000_s3_l19.naf")
if fastq:
    data = K.cfg4_fastq(int(mbp))
else:
    data = open(p, "rb").read() if mbp >= 250 and os.path.exists(p) else K.cfg3_chromosome(int(mbp * 1e6), workers=0)
ctx = N.Context(0)
arc = N.parse_archive(data)
ctx.prepare([arc]); ctx.sync()
ctx.time_runs(2, True)
ms = ctx.time_runs(5, True) / 5
ctx.run(); ctx.fetch_raw()
st = ctx.stats()
print(f"{mbp} Mbp: device {ms:.3f} ms; blocks {st.n_blocks} sequences {st.n_sequences}; rounds {st.lz_rounds} handover {st.lz_handover} unresolved {st.lz_unresolved}")
print("pending after round k:", [int(x) for x in st.lz_pending if x])
print("stages:", {k: round(v, 3) for k, v in ctx.profile_stages() if v > 0.003})
