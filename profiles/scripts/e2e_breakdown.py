"""host-side time breakdown of epg.simulate on the bench workload (EPGX_TIMING=1) + the host expansion bandwidth"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["EPGX_TIMING"] = "1"
import bench
from epgpy_b200 import epg
T1, T2, B1 = bench.grid_axes(bench.GRID)
seq = bench.fisp_sequence(epg, T1, T2, B1, 1000)
for i in range(4):
    t0 = time.perf_counter()
    out = epg.simulate(seq)
    print("call %d: %.1f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
    del out
