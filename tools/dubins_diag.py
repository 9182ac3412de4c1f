#!/usr/bin/env python
"""Where do the persistent kernel and the 5-launch path part ways for DubinsCar?  T = 1..3 steps, first differing record."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import product_algo, product_env
from gcbfplus_b200.trainer.rollout import RolloutEngine

env = product_env("DubinsCar", 12, 2.0, 4)
algo = product_algo(env, "DubinsCar")
g0 = env.reset(21, n_envs=2)
recs = {}
for persistent in (True, False):
    eng = RolloutEngine(env, 2, T=3, n_obs=4, persistent=persistent, use_cuda_graph=False)
    eng.set_params(algo.actor_params); eng.set_initial(g0.agent, g0.goal, g0.obstacle); eng.run(); torch.cuda.synchronize()
    recs[persistent] = {k: getattr(eng, k).clone().cpu().numpy() for k in ("agent", "hits", "actions", "rewards", "costs")}
    recs[persistent]["n_edges"] = eng.counters[:, 0].cpu().numpy()
a, b = recs[True], recs[False]
for t in range(4):
    if t < 3:
        da = np.abs(a["actions"][t] - b["actions"][t]); print("t", t, "actions max diff", da.max(), "at", np.unravel_index(da.argmax(), da.shape), "n differing", int((da > 0).sum()))
    ds = np.abs(a["agent"][t] - b["agent"][t]); print("t", t, "agent   max diff", ds.max(), "per comp", ds.reshape(-1, 4).max(0))
    fin = np.isfinite(a["hits"][t]) & (np.abs(a["hits"][t]) < 1e3)
    dh = np.abs(np.where(fin, a["hits"][t] - b["hits"][t], 0)); print("t", t, "hits    max diff", dh.max(), "n_edges", a["n_edges"][t], b["n_edges"][t])
# u_ref only: is the reference controller identical?  (actions - 2 pi): compare via env.u_ref on the same states
