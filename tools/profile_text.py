"""FASTA text of a few cfg2 archives, formatted on the device twice: the command ncu wraps for k_text_write."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nafcodec_b200 as N
from nafcodec_b200 import _ffi
import bench

n_arch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lib = _ffi.default_library()
ctx = N.Context(0, lib)
uniq = bench.make_workload(8, 5_000_000, 19, 0)
arcs = [N.parse_archive(uniq[i % 8], lib) for i in range(n_arch)]
for _ in range(2):
    out = ctx.format(arcs, _ffi.WANT_ALL, _ffi.TEXT_FASTA)
print("formatted", len(out), "archives,", sum(len(o) for o in out), "bytes")
