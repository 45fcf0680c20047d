"""Which component of LM10 differs between the POT_LM10 / POT_ROTATED / POT_GENERAL kernels? (GPU box, debugging aid)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cogsworth_b200 as cb
from cogsworth_b200 import potential as P

ctx = cb.Context(devices=[0])
q = np.random.default_rng(5).normal(0, 9, (3, 4096))
base = P.LM10Potential()
for label, masses in (("all", (1, 1, 1)), ("disk only", (1, 0, 0)), ("bulge only", (0, 1, 0)), ("halo only", (0, 0, 1)),
                      ("disk+bulge", (1, 1, 0))):
    params = [tuple([p[0] * m] + list(p[1:])) for p, m in zip(base.params, masses)]
    pot = P.CompositePotential(names=base.names, types=base.types, params=params, rotations=base.rotations)
    out = {}
    for kind, env in (("lm10", None), ("rotated", "CW_NO_LM10_KIND"), ("general", "CW_NO_ROTATED_KIND")):
        for e in ("CW_NO_LM10_KIND", "CW_NO_ROTATED_KIND"):
            os.environ.pop(e, None)
        if env:
            os.environ[env] = "1"
        ctx.set_potential(pot)
        out[kind] = ctx.gradient(q)
    for kind in ("lm10", "rotated"):
        d = np.abs(out[kind] - out["general"]) / np.maximum(np.abs(out["general"]), 1e-300)
        print(f"{label:12s} {kind:8s} vs general: differing entries per row {[(int((d[c] > 0).sum())) for c in range(3)]}, "
              f"max rel {d.max():.2e}")
