"""Shared builders for the tests: the seeded problem the golden fixtures were made from."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pifu_b200 import config, synthetic as syn      # noqa: E402
from oracle import pifu_oracle as orc                # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def calibrated_problem(bias_std=0.01, saturated=False):
    """Exactly the tensors oracle/make_golden.py fed to the reference: un-calibrated
    problem, then the last fine conv rescaled from the pilot predictions (the golden file
    stores the first 4096 pilot values for a cross-check; the oracle recomputes all)."""
    prob = syn.make_problem(bias_std=bias_std)
    coarse, fine = oracle_states(prob)
    pilot = syn.random_points(20000, syn.SEED_PILOT, -1.0, 1.0)
    with torch.no_grad():
        p = orc.query_fine(fine, pilot, syn.default_calib())[0].numpy()
    syn.calibrate_last_layer(prob["fine"], 3, p)
    if saturated:
        syn.saturate(prob["fine"], 3)
    return prob, p


def oracle_states(prob, mlp_norm="none", mode="orthogonal"):
    oc = config.coarse_opt(mlp_norm=mlp_norm)
    of = config.fine_opt(mlp_norm=mlp_norm)
    sdc, sdf = dict(prob["coarse"]), dict(prob["fine"])
    if mlp_norm == "group":
        for sd, dims in ((sdc, oc.mlp_dim), (sdf, of.mlp_dim)):
            for i in range(len(dims) - 2):
                sd["norms.%d.weight" % i] = torch.ones(dims[i + 1])
                sd["norms.%d.bias" % i] = torch.zeros(dims[i + 1])
    coarse = orc.CoarseState(sdc, prob["feat_coarse"], oc, mode)
    fine = orc.FineState(sdf, prob["feat_fine"], of, coarse, mode)
    return coarse, fine
