"""Closed-loop rates of the reference's pretrained GCBF+ models on the B200 path, with the reference's own
scenario sampler (threefry keys of test.py: seed 1234, episode i -> split(split(PRNGKey(seed), 1000)[i])[0]).
Mirrors `python test.py --path pretrained/<Env>/gcbf+ --epi 5 --area-size <L> -n <N> --obs <O>` (README.md:104) but
reads the committed parameter fixtures (tests/golden/params_<Env>.npz), so it runs on a box without /root/reference.
Usage: python profiles/eval_rates.py > profiles/r01_eval_rates.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import product_algo  # noqa: E402
from gcbfplus_b200.env import make_env  # noqa: E402
from gcbfplus_b200.trainer.rollout import RolloutEngine  # noqa: E402
from gcbfplus_b200.trainer.utils import test_rates  # noqa: E402
from gcbfplus_b200.utils import jrandom as jr  # noqa: E402

CASES = [  # env, n, area, obs   (README: -n 16 --area-size 4 --obs 0; larger swarms at the same density)
    ("DoubleIntegrator", 16, 4.0, 0), ("DoubleIntegrator", 16, 4.0, 8), ("DoubleIntegrator", 64, 8.0, 8),
    ("DoubleIntegrator", 256, 16.0, 8), ("DoubleIntegrator", 512, 32.0, 8),
    ("SingleIntegrator", 16, 4.0, 0), ("DubinsCar", 16, 4.0, 0), ("LinearDrone", 16, 2.0, 0),
]
SEED, EPI = 1234, 5
print(f"# pretrained gcbf+ models, {EPI} episodes, T = 256, seed {SEED}; rates as test.py:184-198")
print(f"{'env':18s} {'n':>4s} {'area':>5s} {'obs':>3s} | {'safe %':>7s} {'finish %':>8s} {'success %':>9s} | {'reward':>8s} {'cost':>7s}")
for env_id, n, area, n_obs in CASES:
    env = make_env(env_id, n, area_size=area, num_obs=n_obs)
    algo = product_algo(env, env_id)
    keys = jr.split(jr.PRNGKey(SEED), 1_000)[:EPI]
    g0 = env.reset(jr.split(keys, 2)[:, 0])
    eng = RolloutEngine(env, EPI, T=env.max_episode_steps, n_obs=n_obs)
    eng.set_params(algo.actor_params)
    eng.set_initial(g0.agent, g0.goal, g0.obstacle)
    eng.run()
    ro = eng.result()
    rates, is_unsafe, is_finish = test_rates(env, ro)
    succ = (1 - is_unsafe) * is_finish
    print(f"{env_id:18s} {n:4d} {area:5.1f} {n_obs:3d} | {(1 - is_unsafe).mean() * 100:7.2f} {is_finish.mean() * 100:8.2f} "
          f"{succ.mean() * 100:9.2f} | {float(ro.rewards.sum(dim=1).mean()):8.3f} {float(ro.costs.sum(dim=1).mean()):7.3f}")
