"""Diagnostic: per-phase cycle counts of k_huf_decode (NAFGPU_DEBUG_HUF=1) on a batch of cfg2 archives."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["NAFGPU_DEBUG_HUF"] = "1"
import nafcodec_b200 as N
import bench
arcs = bench.make_workload(2, 5_000_000, 19, 0)
res = N.decode_batch(arcs * 8)
print("decoded", len(res))
