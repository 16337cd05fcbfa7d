"""profiling driver: BASELINE configs[3] (RF-spoiled GRE, 3-d gradients, diffusion, 500 TRs) on the complex register kernel"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench_configs as bc
from epgpy_b200 import engine, epg, lowering
seq, opts, jac = bc.cfg_gre_diffusion(epg)
low = lowering.lower(seq, options=opts)
plan = engine.Plan(low)
print(plan.config())
for _ in range(3):
    plan.run(0)
    torch.cuda.synchronize()
print("ok")
