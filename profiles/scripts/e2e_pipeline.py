"""anatomy of the end-to-end pipeline of epg.simulate on the bench workload: Plan.run_to_host_real (kernel chunks -> D2H of
rows of reals -> host widening) with different chunk counts / host threads / staging depths, with the widening switched
off (kernel + PCIe only) and with the kernel alone.  python profiles/scripts/e2e_pipeline.py > gpurun_out/e2e_pipeline.txt"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from epgpy_b200 import epg, engine, lowering, functions

T1, T2, B1 = bench.grid_axes(bench.GRID)
seq = bench.fisp_sequence(epg, T1, T2, B1, 1000)
t0 = time.perf_counter(); low = lowering.lower(seq); t1 = time.perf_counter()
plan = engine.Plan(low); t2 = time.perf_counter()
print("cpu_count %d; lower %.1f ms, plan %.1f ms" % (os.cpu_count(), 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
out = functions._host_buffer((low.nadc, low.natoms, 1), torch.complex128)
L = engine.lib()
real_expand = L.epgx_expand_real

def run(tag, reps=3, **kw):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter()
        plan.run_to_host_real(0, out, 0, low.natoms, **kw)
        ts.append(1e3 * (time.perf_counter() - t))
    print("%-44s %s" % (tag, " ".join("%.1f" % x for x in ts)), flush=True)

run("32 chunks, every chunk as rows of reals", nchunk=32, complex_every=0)
for ce in (2, 3, 4, 5, 6, 8):
    run("32 chunks, every %d-th chunk complex" % ce, nchunk=32, complex_every=ce)
run("64 chunks, every 4th chunk complex", nchunk=64, complex_every=4)
run("32 chunks, every 4th complex, 12 threads", nchunk=32, complex_every=4, nthreads=12)
# D2H alone: 8 GB of rows of reals, device -> pinned host, in 32 pieces
per = low.natoms // 32
d = torch.empty((low.nadc, per), dtype=torch.float64, device="cuda:0")
h = [torch.empty((low.nadc, per), dtype=torch.float64, pin_memory=True) for _ in range(4)]
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(32):
        h[i % 4].copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    print("D2H alone, 32 x %d MB: %.1f ms" % (d.numel() * 8 >> 20, 1e3 * (time.perf_counter() - t)), flush=True)
# widening alone (16 threads), then with the D2H stream running beside it
import threading
def widen(nt=16):
    t = time.perf_counter()
    for i in range(32):
        real_expand(engine.DTYPES["f64"], h[i % 4].data_ptr(), per, out.data_ptr() + i * per * 16, low.natoms, low.nadc, per, nt)
    return 1e3 * (time.perf_counter() - t)
print("widening alone, 16 threads: %.1f ms" % widen(), flush=True)
print("widening alone, 32 threads: %.1f ms" % widen(32), flush=True)
def copies():
    for i in range(32):
        h[2 + i % 2].copy_(d, non_blocking=True)
    torch.cuda.synchronize()
th = threading.Thread(target=copies); th.start()
print("widening beside the D2H stream, 16 threads: %.1f ms" % widen(), flush=True)
th.join()
