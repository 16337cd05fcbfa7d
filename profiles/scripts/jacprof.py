import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench_configs as bc
from epgpy_b200 import engine, epg, lowering
# usage: jacprof.py [kernel variant (4 = realjac, 5 = setjac)] [dtype]
kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dtype = sys.argv[2] if len(sys.argv) > 2 else "f64"
seq, opts, jac = bc.cfg_fisp(epg, (30,30,30), 1000, jac=True)
low = lowering.lower(seq, probe=[None, epg.Jacobian(jac)], options=opts, dtype=dtype)
plan = engine.Plan(low)
if kernel:
    plan.set_variant(kernel=kernel)
print(plan.config())
for _ in range(2):
    s,j = plan.run(0)
    torch.cuda.synchronize()
print('ok')
