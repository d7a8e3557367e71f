import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench, nafcodec_b200 as N
arcs = bench.make_workload(1, 5_000_000, 19, 0)
ctx = N.Context(0)
for i in range(2):
    ctx.decode([N.parse_archive(arcs[0])])
