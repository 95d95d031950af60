import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmad_b200 import Parameters, objectives as ob
from tests.golden.materials import objective_trees
RD = np.load("tests/golden/ref_def_types_rate.npz")
dev = torch.device("cuda:0")
for case in ("hill_rot.UNIAXIAL_STRESS", "hill_rot.PLANE_STRESS"):
    kind, dtn = case.split(".")
    for tag in ("scaled",):
        pre = f"{case}.obj_{tag}"
        for strategy, ctor in (("adjoint", ob.MPAdjointObjective), ("direct", ob.MPDirectObjective)):
            values, act, tr = objective_trees(kind, tag == "scaled")
            P = Parameters(values, act, tr)
            model = ob.SmallRateElasticPlastic(P, def_type=getattr(ob, dtn))
            obj = ctor(ob.Calibration(model, RD[f"{pre}.data"], RD[f"{pre}.weight"]), RD[f"{case}.F"], device=dev)
            r = obj.evaluate(RD[f"{pre}.x_canonical"])
            print(case, tag, strategy, r.J, RD[f"{pre}.J_{strategy}"], r.grad, RD[f"{pre}.grad_{strategy}"])
