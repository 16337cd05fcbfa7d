"""shared helpers of the parity suite"""

import numpy as np

import cases

RTOL64 = 1e-10  # BASELINE.json north_star: FP64 within 1e-10 relative of the reference
RTOL32 = 1e-5   # BASELINE.json north_star: FP32 mode within 1e-5


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / scale)


def product_namespace():
    import epgpy_b200

    return cases.namespace(epgpy_b200)


def run_case(simulate, epg, case, **extra):
    """run a tests/cases.py case through `simulate` (product or tape interpreter) -> (signal, jac|None)"""
    opts = dict(case.get("options") or {})
    opts.update(extra)
    if case.get("density") is not None:
        opts["init"] = epg.StateMatrix(density=case["density"])
    if case.get("init") is not None:
        opts["init"] = np.array(case["init"])
    if case.get("jac"):
        sig, jac = simulate(case["seq"], probe=[None, epg.Jacobian(case["jac"])], **opts)
        return np.asarray(sig), np.asarray(jac)
    return np.asarray(simulate(case["seq"], **opts)), None
