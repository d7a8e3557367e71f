"""Decodes a small batch of cfg2 archives a few times (the command ncu wraps; see profiles/README.md)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nafcodec_b200 as N
import bench

n_arch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
arcs = bench.make_workload(min(n_arch, 8), 5_000_000, 19, 0)
batch = [arcs[i % len(arcs)] for i in range(n_arch)]
if len(sys.argv) > 3 and sys.argv[3] == "fixture":      # cfg1 first (latency of a single small archive), then the cfg2 batch
    fx = open(os.path.join(ROOT, "tests", "golden", "NZ_AAEN01000029.naf"), "rb").read()
    for _ in range(reps):
        N.decode_batch([fx])
for _ in range(reps):
    res = N.decode_batch(batch)
print("decoded", len(res), "archives,", sum(r.total_residues for r in res), "residues")
