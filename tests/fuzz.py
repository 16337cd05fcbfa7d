"""Random sequence generator for the fuzz tests (test infrastructure)."""

import numpy as np


def random_case(epg, seed, real_only=False, with_jac=False):
    """a random but valid sequence of hot-path operators on a small random grid"""
    rng = np.random.RandomState(seed)
    n1, n2, n3 = rng.randint(1, 5), rng.randint(1, 4), rng.randint(1, 4)
    T1 = rng.uniform(200, 2000, n1)
    T2 = rng.uniform(20, 200, (1, n2))
    B1 = rng.uniform(0.7, 1.3, (1, 1, n3))
    g = rng.uniform(-0.05, 0.05, (1, n2))
    nops = rng.randint(8, 60)
    seq = []
    if rng.rand() < 0.3:
        seq.append(epg.PD(rng.uniform(0.5, 2.0, n1)))
    o1 = (lambda **kw: kw) if with_jac else (lambda **kw: {})
    nvar = 0
    last_shift = None
    for i in range(nops):
        r = rng.rand()
        prev_shift, last_shift = last_shift, None
        if r < 0.3:
            alpha = rng.uniform(5, 175) * (B1 if rng.rand() < 0.6 else 1.0)
            if real_only:
                phi = float(rng.choice([90.0, 270.0, -90.0]))
            else:
                phi = float(rng.choice([0.0, 90.0, 180.0, rng.uniform(0, 360)]))
            kw = {}
            if with_jac and rng.rand() < 0.5:
                kw = {"order1": {"B1": {"alpha": float(rng.uniform(5, 60))}}}
                if rng.rand() < 0.2 and nvar < 4:
                    kw["order1"][f"a{nvar}"] = {"alpha": 1.0}
                    nvar += 1
            seq.append(epg.T(alpha, phi, **kw))
        elif r < 0.55:
            kw = {}
            if with_jac and rng.rand() < 0.7:
                kw = {"order1": ["T1", "T2"]}
            if real_only or rng.rand() < 0.5:
                seq.append(epg.E(rng.uniform(1, 20), T1, T2, **kw))
            else:
                seq.append(epg.E(rng.uniform(1, 20), T1, T2, g, **kw))
        elif r < 0.75:
            last_shift = int(rng.choice([1, 1, 1, -1, 2, -2]))
            seq.append(epg.S(last_shift))
        elif r < 0.8:
            seq.append(epg.D(rng.uniform(1, 10), rng.uniform(0.5e-3, 3e-3), k=(None if rng.rand() < 0.5 else prev_shift)))
        elif r < 0.83:
            seq.append(epg.SPOILER)
        elif r < 0.85 and not with_jac:
            seq.append(epg.RESET)
        elif r < 0.87 and not real_only:
            seq.append(epg.P(rng.uniform(1, 5), g))
        else:
            q = rng.rand()
            if q < 0.6:
                seq.append(epg.ADC)
            elif q < 0.8:
                seq.append(epg.Adc("Z0"))
            else:
                seq.append(epg.Adc(phase=float(rng.uniform(-180, 180))))
    seq.append(epg.ADC)
    opts = {"kvalue": float(rng.uniform(500, 3000))}
    if rng.rand() < 0.5:
        opts["max_nstate"] = int(rng.randint(1, 12))
    jac = (["B1", "T1", "T2"] + [f"a{i}" for i in range(nvar)]) if with_jac else None
    return seq, opts, jac
