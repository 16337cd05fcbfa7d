import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench_configs as bc
from epgpy_b200 import engine, epg, lowering
seq, opts, jac = bc.cfg_fisp(epg, (30,30,30), 1000, jac=True)
low = lowering.lower(seq, probe=[None, epg.Jacobian(jac)], options=opts)
plan = engine.Plan(low)
print(plan.config())
for _ in range(2):
    s,j = plan.run(0)
    torch.cuda.synchronize()
print('ok')
