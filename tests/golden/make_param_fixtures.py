"""Generates tests/golden/params_<Env>.npz from the reference's pretrained pickles.

Run HERE (the container that mounts /root/reference); the GPU box has no
/root/reference, so the parameter sets travel as these fixtures.  They are
parameter *data* (pretrained/<Env>/gcbf+/models/1000/{actor,cbf}.pkl), not code.
Usage: python tests/golden/make_param_fixtures.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.nn import flatten_params, load_ref_pickle  # noqa: E402

REF = "/root/reference/pretrained"
OUT = os.path.dirname(os.path.abspath(__file__))

for env in ["SingleIntegrator", "DoubleIntegrator", "DubinsCar", "LinearDrone"]:
    flat = {}
    for net in ["actor", "cbf"]:
        tree = load_ref_pickle(f"{REF}/{env}/gcbf+/models/1000/{net}.pkl")
        for k, v in flatten_params(tree).items():
            flat[f"{net}:{k}"] = np.asarray(v, dtype=np.float32)
    np.savez(os.path.join(OUT, f"params_{env}.npz"), **flat)
    print(env, sum(v.size for k, v in flat.items() if k.startswith("actor")),
          sum(v.size for k, v in flat.items() if k.startswith("cbf")))
