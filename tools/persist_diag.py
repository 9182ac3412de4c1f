#!/usr/bin/env python
"""Persistent-rollout diagnostics on a GPU: cluster occupancy of the kernel and the in-kernel phase / cluster stamps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from helpers import product_algo, product_env  # noqa: E402


def main():
    from gcbfplus_b200 import _lib
    from gcbfplus_b200.trainer.rollout import RolloutEngine
    lib = _lib.load()
    for c in (1, 2, 4, 8, 16):
        print("max active clusters, cluster size", c, "->", lib.gcbf_rollout_persistent_max_clusters(c))
    E, N, T = int(os.environ.get("E", 16)), int(os.environ.get("N", 512)), 24
    env = product_env("DoubleIntegrator", N, (2 * N) ** 0.5, 8)
    algo = product_algo(env, "DoubleIntegrator")
    g0 = env.reset(1000, n_envs=E)
    eng = RolloutEngine(env, E, T=T, n_obs=8, use_cuda_graph=False, persistent=True)
    eng.phase_stamps = torch.zeros((T + 1) * 8 + 2 * E, dtype=torch.int64, device=env.device)
    eng.set_params(algo.actor_params)
    eng.set_initial(g0.agent, g0.goal, g0.obstacle)
    eng.run()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.run()
    ev1.record()
    torch.cuda.synchronize()
    st = eng.phase_stamps.cpu().numpy().astype(np.float64)
    ph = st[: (T + 1) * 8].reshape(T + 1, 8)
    cl = st[(T + 1) * 8:].reshape(E, 2)
    t0 = cl[:, 0].min()
    print("event time of the launch: %.1f us = %.2f us / step" % (ev0.elapsed_time(ev1) * 1e3, ev0.elapsed_time(ev1) * 1e3 / T))
    print("cluster start / end (us after the first start):")
    for e in range(E):
        print("  env %2d  %8.1f  %8.1f" % (e, (cl[e, 0] - t0) / 1e3, (cl[e, 1] - t0) / 1e3))
    d = (ph[2:, 1:] - ph[2:, :-1]).mean(axis=0) / 1e3
    print("phase us (E, A, U1, U2, tail, lidar+bits, fill):", np.round(d, 2), "sum", round(float(d.sum()), 2))
    print("build row (t = -1):", np.round((ph[0, 1:] - ph[0, :-1]) / 1e3, 2))


if __name__ == "__main__":
    main()
