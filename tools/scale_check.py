"""Parity + timing of the BASELINE.json config shapes at sizes the oracle finishes in seconds to a minute.
   python tools/scale_check.py [cfg3_mbp] [cfg4_reads]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as K
import _oracle as O
import nafcodec_b200 as N
from _harness import assert_same_as_oracle

cfg3_mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 50
cfg4_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
ctx = N.Context(0)


def run(name, data, **fields):
    t0 = time.perf_counter()
    d = O.decode(data, **fields)
    t_cpu = time.perf_counter() - t0
    arc = N.parse_archive(data)
    want = N.decoder._want_bits(fields.get("id", True), fields.get("comment", True), fields.get("sequence", True),
                                fields.get("quality", True), fields.get("mask", True))
    t0 = time.perf_counter()
    ctx.prepare([arc], want)
    ctx.sync()
    t_prep = time.perf_counter() - t0
    st = ctx.stats()
    ctx.time_runs(2, True)
    ms = ctx.time_runs(5, True) / 5
    ctx.run()
    res = ctx.fetch()[0]
    assert_same_as_oracle(res, d, name)
    handover = ctx.stats().lz_handover
    stages = ", ".join(f"{k}={v:.3f}" for k, v in ctx.profile_stages() if v > 0.02)
    out_bytes = st.ascii_bytes + st.quality_bytes + st.id_bytes + st.comment_bytes
    print(f"{name}: PARITY OK | {len(data) / 1e6:.1f} MB archive, {st.n_blocks} zstd blocks, {st.n_sequences} sequences | device {ms:.3f} ms "
          f"({out_bytes / ms / 1e6:.1f} GB/s out, {st.algorithmic_bytes / ms / 1e6 / 6557.8 * 100:.2f}% of HBM peak) | host prepare {t_prep * 1e3:.1f} ms | "
          f"cpu oracle {t_cpu * 1e3:.0f} ms ({out_bytes / t_cpu / 1e9:.2f} GB/s) | lz rounds {ctx.stats().lz_rounds}, handover at round {handover} ({ctx.stats().lz_unresolved} bytes to level 2) | stages[ms]: {stages}", flush=True)


t0 = time.time()
n3 = int(cfg3_mbp * 1e6)
cfg3 = K.genome(3, n3, level=int(os.environ.get("CFG3_LEVEL", "19")), gaps=20, gap_len=max(n3 // 5000, 100), telomere=10_000, mask=True, records=1,
                mean_u=300.0, mean_m=300.0)
print(f"generated cfg3-shape {cfg3_mbp} Mbp in {time.time() - t0:.1f}s", flush=True)
run(f"cfg3 ({cfg3_mbp} Mbp, dense mask, N gaps)", cfg3)
t0 = time.time()
cfg4 = K.fastq_reads(4, cfg4_reads, level=0, with_mask=True)
print(f"generated cfg4-shape {cfg4_reads} reads in {time.time() - t0:.1f}s", flush=True)
run(f"cfg4 ({cfg4_reads} x 150 bp FASTQ, all fields)", cfg4)
run(f"cfg4 ({cfg4_reads} x 150 bp FASTQ, quality=False)", cfg4, quality=False)
run("cfg1 (NZ_AAEN01000029.naf)", open(os.path.join(ROOT, "tests", "golden", "NZ_AAEN01000029.naf"), "rb").read())
