"""Cycle accounting of the single-archive path (NAFGPU_DEBUG_HUF=1 python tools/lat_debug.py): cfg1 fixture, then cfg2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nafcodec_b200 as N
fx = open(os.path.join(ROOT, "tests", "golden", "NZ_AAEN01000029.naf"), "rb").read()
ctx = N.Context(0)
for name, data in (("cfg1", fx), ("cfg2", bench.make_workload(1, 5_000_000, 19, 0)[0])):
    print(name, file=sys.stderr)
    for i in range(2):
        ctx.decode([N.parse_archive(data)])
    print({k: round(v, 4) for k, v in ctx.profile_stages() if v > 0.002}, file=sys.stderr)
