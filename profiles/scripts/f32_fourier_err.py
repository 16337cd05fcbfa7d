import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import cases
from util import product_namespace
epg=product_namespace()
for name in sorted(cases.FOURIER_CASES):
    ref=np.load(f'tests/golden/{name}.npz')
    case=cases.FOURIER_CASES[name](epg)
    vals=epg.simulate(case["seq"], asarray=False, dtype="float32", **case["options"])
    print(name, max(np.abs(np.asarray(v)-ref[f"probe{i}"]).max()/max(np.abs(ref[f"probe{i}"]).max(),1e-30) for i,v in enumerate(vals)))
