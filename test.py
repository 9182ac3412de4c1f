"""test.py -- evaluation CLI with the reference's flags (test.py:239-266): rollouts with a trained
(or --u-ref nominal) controller and safe / finish / success rates (test.py:184-198).
Video rendering is out of scope -> --no-video is implied; --cbf <agent id> computes the CBF
contour grids the reference hands to its renderer (test.py:125-131, trainer/utils.py:149-168) and saves them as
<path>/cbf_contours/epi<k>_agent<id>.npz (b_xs, b_ys, bb_h per time step).  --nojit-rollout is accepted: the
reference needs it to survive n >= 512 with dense graphs (env/base.py:191-259); the sparse rollout engine has no such limit."""
import argparse
import os

import numpy as np
import yaml

from gcbfplus_b200.algo import make_algo
from gcbfplus_b200.env import make_env
from gcbfplus_b200.trainer.rollout import RolloutEngine
from gcbfplus_b200.trainer.utils import cbf_contours, test_rates


def test(args):
    print(f"> Running test.py {args}")
    if args.cpu:
        raise SystemExit("--cpu: gcbfplus_b200 is the sm_100a CUDA path only (no CPU fallback by design)")
    np.random.seed(args.seed)
    config = None
    if not args.u_ref and args.path is not None:
        with open(os.path.join(args.path, "config.yaml"), "r") as f:
            config = yaml.load(f, Loader=yaml.UnsafeLoader)
    num_agents = config.num_agents if args.num_agents is None else args.num_agents
    env = make_env(env_id=config.env if args.env is None else args.env, num_agents=num_agents, num_obs=args.obs,
                   area_size=args.area_size, max_step=args.max_step, max_travel=args.max_travel)
    policy = "u_ref"
    algo = None
    if not args.u_ref:
        assert args.path is not None, "--path or --u-ref required"
        model_path = os.path.join(args.path, "models")
        step = max(int(m) for m in os.listdir(model_path) if m.isdigit()) if args.step is None else args.step
        print("step: ", step)
        algo = make_algo(
            algo=config.algo, env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
            action_dim=env.action_dim, n_agents=env.num_agents, gnn_layers=config.gnn_layers,
            batch_size=config.batch_size, buffer_size=config.buffer_size, horizon=config.horizon,
            lr_actor=config.lr_actor, lr_cbf=config.lr_cbf, alpha=config.alpha, eps=0.02, inner_epoch=8,
            loss_action_coef=config.loss_action_coef, loss_unsafe_coef=config.loss_unsafe_coef,
            loss_safe_coef=config.loss_safe_coef, loss_h_dot_coef=config.loss_h_dot_coef, max_grad_norm=2.0,
            seed=config.seed)
        algo.load(model_path, step)
        policy = "actor"
        path = args.path
    else:
        assert args.env is not None
        path = os.path.join(f"./logs/{args.env}/nominal")
        os.makedirs(path, exist_ok=True)
    n_epi = args.epi - args.offset
    eng = RolloutEngine(env, n_epi, T=env.max_episode_steps, policy=policy)
    if algo is not None:
        eng.set_params(algo.actor_params)
    # test.py:117-119,158: test_keys = split(PRNGKey(seed), 1000)[:epi][offset:]; episode i resets with
    # split(test_keys[i])[0].  All episodes run as one batch here.
    from gcbfplus_b200.utils import jrandom as jr
    test_keys = jr.split(jr.PRNGKey(args.seed), 1_000)[: args.epi][args.offset:]
    g0 = env.reset(jr.split(test_keys, 2)[:, 0])
    eng.set_initial(g0.agent, g0.goal, g0.obstacle)
    eng.run()
    ro = eng.result()
    rates, is_unsafe, is_finish = test_rates(env, ro)
    rewards = ro.rewards.sum(dim=1).cpu().numpy()
    costs = ro.costs.sum(dim=1).cpu().numpy()
    for i in range(n_epi):
        print(f"epi: {i}, reward: {rewards[i]:.3f}, cost: {costs[i]:.3f}, safe rate: {rates[i, 0] * 100:.3f}%,"
              f"finish rate: {rates[i, 1] * 100:.3f}%, success rate: {rates[i, 2] * 100:.3f}%")
    safe_mean, safe_std = (1 - is_unsafe).mean(), (1 - is_unsafe).std()
    finish_mean, finish_std = is_finish.mean(), is_finish.std()
    succ = (1 - is_unsafe) * is_finish
    print(f"reward: {np.mean(rewards):.3f}, min/max reward: {np.min(rewards):.3f}/{np.max(rewards):.3f}, "
          f"cost: {np.mean(costs):.3f}, min/max cost: {np.min(costs):.3f}/{np.max(costs):.3f}, "
          f"safe_rate: {safe_mean * 100:.3f}%, finish_rate: {finish_mean * 100:.3f}%, "
          f"success_rate: {succ.mean() * 100:.3f}%")
    if args.log:
        with open(os.path.join(path, "test_log.csv"), "a") as f:
            f.write(f"{env.num_agents},{args.epi},{env.max_episode_steps},{env.area_size},{env.params['n_obs']},"
                    f"{safe_mean * 100:.3f},{safe_std * 100:.3f},{finish_mean * 100:.3f},{finish_std * 100:.3f},"
                    f"{succ.mean() * 100:.3f},{succ.std() * 100:.3f}\n")
    if args.cbf is not None:
        assert algo is not None, "--cbf needs a trained CBF (--path)"
        out_dir = os.path.join(path, "cbf_contours")
        os.makedirs(out_dir, exist_ok=True)
        for i in range(n_epi):
            b_x, b_y, bb_h = cbf_contours(algo, env, ro, i, args.cbf)
            f_out = os.path.join(out_dir, f"epi{i + args.offset:02}_agent{args.cbf}.npz")
            np.savez_compressed(f_out, b_xs=b_x, b_ys=b_y, bb_h=bb_h, agent_id=args.cbf)
            print(f"cbf contour grid: {f_out} {bb_h.shape}")
    if not args.no_video:
        print("video rendering is out of scope of the B200 hot path (SURVEY.md section 2, row 17); skipped")


# the reference's command line (test.py:239-266), table-driven like train.py
FLAGS = [
    (("-n", "--num-agents"), int, None), (("--obs",), int, 0), (("--area-size",), float, "required"),
    (("--max-step",), int, None), (("--path",), str, None), (("--n-rays",), int, 32), (("--alpha",), float, 1.0),
    (("--max-travel",), float, None), (("--cbf",), int, None), (("--seed",), int, 1234), (("--debug",), "flag", False),
    (("--cpu",), "flag", False), (("--u-ref",), "flag", False), (("--env",), str, None), (("--algo",), str, None),
    (("--step",), int, None), (("--epi",), int, 5), (("--offset",), int, 0), (("--no-video",), "flag", False),
    (("--nojit-rollout",), "flag", False), (("--log",), "flag", False), (("--dpi",), int, 100),
]


def main():
    from train import build_parser
    test(build_parser(FLAGS).parse_args())


if __name__ == "__main__":
    main()
