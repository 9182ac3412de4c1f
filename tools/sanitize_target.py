#!/usr/bin/env python
"""Small, complete pass over the hot path for compute-sanitizer (memcheck / racecheck on every
kernel): reset -> graph build -> 3-step rollout (eager launches: every per-step kernel incl. the tcgen05 / TMA ones)
-> labels -> QP action labels -> one train step (3 forwards + backward + clip + AdamW) -> polyak.

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
    compute-sanitizer --tool synccheck python tools/sanitize_target.py

Sizes are small because the tools serialise and instrument every access; N = 40 still gives several row tiles'
worth of edges per GEMM (ragged last tile) and multi-CTA graph builds."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from helpers import product_algo, product_env  # noqa: E402


def main():
    env_id = sys.argv[1] if len(sys.argv) > 1 else "DoubleIntegrator"
    from gcbfplus_b200.algo.train import label_rollout, qp_labels, train_minibatch, update_tgt
    from gcbfplus_b200.trainer.rollout import RolloutEngine
    N, E, T = 40, 3, 3
    env = product_env(env_id, N, 2.5 if env_id != "LinearDrone" else 1.6, 4)
    env.edge_cap_per_agent = 48
    algo = product_algo(env, env_id)
    g0 = env.reset(3, n_envs=E)
    persistent = os.environ.get("GCBF_SANITIZE_PERSISTENT", "0") == "1"        # the single-launch rollout kernel instead
    eng = RolloutEngine(env, E, T=T, n_obs=4, use_cuda_graph=False, persistent=persistent)
    eng.set_params(algo.actor_params)
    eng.set_initial(g0.agent, g0.goal, g0.obstacle)
    eng.run()
    ro = eng.result()
    safe, unsafe = label_rollout(algo, ro)
    g = env.get_graph(ro.agent[:, 1].contiguous(), ro.goal, None, hits=ro.hits[:, 1].contiguous())
    u_qp = qp_labels(algo, g)
    train_minibatch(algo, g, safe[:, 1], unsafe[:, 1], u_qp, apply=True)
    update_tgt(algo, 0.5)
    h = algo.get_cbf(g)
    torch.cuda.synchronize()
    g.check_overflow()
    assert torch.isfinite(h).all() and torch.isfinite(algo.cbf_params.flat).all()
    print("sanitize target ok:", env_id, "launches:", env.lib.gcbf_launch_count())


if __name__ == "__main__":
    main()
