"""Concurrent contexts whose jobs all take the finisher path (cooperative kernels from several streams at once): parity + no hang."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nafcodec_b200 as N
from nafcodec_b200 import _ffi
import _cases as K
import _oracle as O
from _harness import assert_same_as_oracle

lib = _ffi.default_library()
datas = [K.fastq_reads(20 + i, 20000) for i in range(4)] + [K.genome(5, 2_000_000, level=19)]
want = [O.decode(d) for d in datas]
arcs = [N.parse_archive(d, lib) for d in datas]
batch = [arcs[i % len(arcs)] for i in range(24)]
pipe = N.Pipeline(0, 6, lib)
t0 = time.time()
for it in range(20):
    res = pipe.decode(batch)
    if it % 5 == 0:
        for i, r in enumerate(res):
            assert_same_as_oracle(r, want[i % len(arcs)], f"iter {it} archive {i}")
print(f"20 iterations x 24 archives on 6 lanes ok in {time.time() - t0:.1f} s; handover flags", [int(s.lz_handover) for s in pipe.stats()])
pipe.close()
