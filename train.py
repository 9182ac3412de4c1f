"""train.py -- same flags as the reference's train.py:115-151 (+ --cpu which is rejected: the
product path has no CPU fallback; the CPU restatement lives in oracle/ for tests only)."""
import argparse
import datetime
import os

import numpy as np
import yaml

from gcbfplus_b200.algo import make_algo
from gcbfplus_b200.env import make_env
from gcbfplus_b200.trainer.trainer import Trainer


def train(args):
    print(f"> Running train.py {args}")
    if args.cpu:
        raise SystemExit("--cpu: gcbfplus_b200 is the sm_100a CUDA path only (no CPU fallback by design)")
    os.environ.setdefault("WANDB_MODE", "offline")
    np.random.seed(args.seed)
    # one process per GPU under torchrun (python -m torch.distributed.run --nproc-per-node N train.py ...):
    # environments are sharded over the ranks, gradients all-reduced once per optimizer step
    import torch
    from gcbfplus_b200 import dist as gdist
    rank, local_rank, world = gdist.init_from_env()
    device = torch.device("cuda", local_rank)
    if args.debug or rank != 0:
        os.environ["WANDB_MODE"] = "disabled"
    env = make_env(env_id=args.env, num_agents=args.num_agents, num_obs=args.obs, n_rays=args.n_rays,
                   area_size=args.area_size, device=device)
    env_test = make_env(env_id=args.env, num_agents=args.num_agents, num_obs=args.obs, n_rays=args.n_rays,
                        area_size=args.area_size, device=device)
    algo = make_algo(
        algo=args.algo, env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
        action_dim=env.action_dim, n_agents=env.num_agents, gnn_layers=args.gnn_layers, batch_size=256,
        buffer_size=args.buffer_size, horizon=args.horizon, lr_actor=args.lr_actor, lr_cbf=args.lr_cbf,
        alpha=args.alpha, eps=0.02, inner_epoch=8, loss_action_coef=args.loss_action_coef,
        loss_unsafe_coef=args.loss_unsafe_coef, loss_safe_coef=args.loss_safe_coef,
        loss_h_dot_coef=args.loss_h_dot_coef, max_grad_norm=2.0, seed=args.seed)
    start_time = datetime.datetime.now().strftime("%Y%m%d%H%M%S")
    log_dir = f"{args.log_dir}/{args.env}/{args.algo}/seed{args.seed}_{start_time}"
    if rank == 0:
        os.makedirs(log_dir, exist_ok=True)
    run_name = f"{args.algo}_{args.env}_{start_time}" if args.name is None else args.name
    train_params = {"run_name": run_name, "training_steps": args.steps, "eval_interval": args.eval_interval,
                    "eval_epi": args.eval_epi, "save_interval": args.save_interval}
    trainer = Trainer(env=env, env_test=env_test, algo=algo, log_dir=log_dir, n_env_train=args.n_env_train,
                      n_env_test=args.n_env_test, seed=args.seed, params=train_params,
                      save_log=not args.debug and rank == 0)
    if not args.debug and rank == 0:
        with open(f"{log_dir}/config.yaml", "w") as f:
            yaml.dump(args, f)
            yaml.dump(algo.config, f)
    trainer.train()


def main():
    parser = argparse.ArgumentParser()
    # custom arguments (train.py:119-128)
    parser.add_argument("-n", "--num-agents", type=int, default=8)
    parser.add_argument("--algo", type=str, default="gcbf+")
    parser.add_argument("--env", type=str, default="SimpleCar")
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--steps", type=int, default=1000)
    parser.add_argument("--name", type=str, default=None)
    parser.add_argument("--debug", action="store_true", default=False)
    parser.add_argument("--obs", type=int, default=None)
    parser.add_argument("--n-rays", type=int, default=32)
    parser.add_argument("--area-size", type=float, required=True)
    # gcbf / gcbf+ arguments (train.py:131-140)
    parser.add_argument("--gnn-layers", type=int, default=1)
    parser.add_argument("--alpha", type=float, default=1.0)
    parser.add_argument("--horizon", type=int, default=32)
    parser.add_argument("--lr-actor", type=float, default=3e-5)
    parser.add_argument("--lr-cbf", type=float, default=3e-5)
    parser.add_argument("--loss-action-coef", type=float, default=0.0001)
    parser.add_argument("--loss-unsafe-coef", type=float, default=1.0)
    parser.add_argument("--loss-safe-coef", type=float, default=1.0)
    parser.add_argument("--loss-h-dot-coef", type=float, default=0.01)
    parser.add_argument("--buffer-size", type=int, default=512)
    # default arguments (train.py:143-148)
    parser.add_argument("--n-env-train", type=int, default=16)
    parser.add_argument("--n-env-test", type=int, default=32)
    parser.add_argument("--log-dir", type=str, default="./logs")
    parser.add_argument("--eval-interval", type=int, default=1)
    parser.add_argument("--eval-epi", type=int, default=1)
    parser.add_argument("--save-interval", type=int, default=10)
    parser.add_argument("--cpu", action="store_true", default=False)
    args = parser.parse_args()
    train(args)


if __name__ == "__main__":
    main()
