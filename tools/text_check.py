"""FASTA text of the bench workload (64 x 5 Mbp archives) formatted on the device: parity with the CPU formatter and
the device time of the two text kernels."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nafcodec_b200 as N
from nafcodec_b200 import _ffi
import _oracle as O
import _cases as K
import bench

lib = _ffi.default_library()
ctx = N.Context(0, lib)
uniq = bench.make_workload(8, 5_000_000, 19, 0)
for name, datas, fmt in (("cfg2 x64 FASTA", [uniq[i % 8] for i in range(64)], "fasta"),
                         ("cfg4 100k reads FASTQ", [K.fastq_reads(11, 100_000)], "fastq")):
    arcs = [N.parse_archive(d, lib) for d in datas]
    want = _ffi.WANT_ALL
    code = {"fasta": _ffi.TEXT_FASTA, "fastq": _ffi.TEXT_FASTQ}[fmt]
    out = ctx.format(arcs, want, code)
    for k in range(min(len(datas), 8)):
        assert out[k] == O.format_text(datas[k], fmt), (name, k)
    ms = []
    for _ in range(5):
        t0 = time.perf_counter(); ctx.format(arcs, want, code); wall = time.perf_counter() - t0
        ms.append(ctx.stats().text_kernel_ms)
    st = ctx.stats()
    m = sorted(ms)[len(ms) // 2]
    # algorithmic bytes: the text written once + ASCII / quality / ids / comments read once
    alg = st.text_bytes + st.ascii_bytes + st.quality_bytes + st.id_bytes + st.comment_bytes
    print(f"{name}: PARITY OK | text {st.text_bytes / 1e6:.1f} MB | text kernels {m:.3f} ms ({st.text_bytes / m / 1e6:.0f} GB/s text out, "
          f"{alg / m / 1e6 / 6557.8 * 100:.1f}% of HBM peak) | whole call incl. decode and D2H of the text {wall * 1e3:.1f} ms")
