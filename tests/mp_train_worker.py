"""Worker for tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU):
the sharded train step (denominator all-reduce + one packed gradient all-reduce over NCCL) must
reproduce the single-GPU full-minibatch gradient, and every rank must end with identical params."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _say(rank, msg):
    print(f"[rank {rank}] {msg}", flush=True)


def main(out_path):
    from helpers import product_algo, product_env, product_obstacles, random_scene
    from gcbfplus_b200 import dist as gd
    from gcbfplus_b200.algo import train as T
    rank, local_rank, world = gd.init_from_env()
    torch.cuda.set_device(local_rank)
    env_id, N, B, area, n_obs = "DoubleIntegrator", 16, 8, 2.0, 4
    agent, goal, obs = random_scene(env_id, N, B, area, n_obs, seed=5)
    env = product_env(env_id, N, area, n_obs, device=f"cuda:{local_rank}")
    env.edge_cap_per_agent = 48
    algo = product_algo(env, env_id)
    algo.loss_action_coef, algo.loss_h_dot_coef, algo.lr_cbf, algo.lr_actor = 0.05, 0.3, 1e-3, 1e-3
    pobs = product_obstacles(env_id, obs, device=env.device)
    dev = env.device
    full = env.get_graph(torch.from_numpy(agent).to(dev), torch.from_numpy(goal).to(dev), pobs)
    rng = np.random.default_rng(1)
    unsafe = env.unsafe_mask(full)
    safe = (~unsafe) & torch.from_numpy(rng.uniform(size=(B, N)) < 0.6).to(dev)
    u_qp = env.u_ref(full) + 0.1
    _say(rank, "scene ready")
    # --- reference: full minibatch on this GPU, collectives disabled
    T._dist = lambda: None
    ts = T.train_minibatch(algo, full, safe, unsafe, u_qp, apply=False)
    ref = ts.packed.clone()
    # --- sharded: this rank's graphs only, collectives on
    del T._dist
    import importlib
    importlib.reload(T)
    lo, hi = gd.shard_bounds(B, rank, world)
    sel = slice(lo, hi)
    shard = env.get_graph(full.agent[sel], full.goal[sel], pobs.select(sel), hits=full.hits[sel].contiguous())
    algo._trainer_state = None
    ts2 = T.train_minibatch(algo, shard, safe[sel], unsafe[sel], u_qp[sel], apply=True)
    torch.cuda.synchronize()
    n = ts2.n_cbf + ts2.n_act
    gmax = float(ref[:n].abs().max())
    err = float((ts2.packed[:n] - ref[:n]).abs().max())
    stats_err = float((ts2.packed[n:n + 10] - ref[n:n + 10]).abs().max())
    # params identical on all ranks after the update
    p = torch.cat([algo.cbf_params.flat, algo.actor_net_params.flat])
    pmax, pmin = p.clone(), p.clone()
    torch.distributed.all_reduce(pmax, op=torch.distributed.ReduceOp.MAX)
    torch.distributed.all_reduce(pmin, op=torch.distributed.ReduceOp.MIN)
    same = bool(torch.equal(pmax, pmin))
    _say(rank, f"sharded eager step done (grad err {err:.2e})")
    # --- the captured optimizer step (algo/train.py MinibatchRunner): gather + graph build + train step + the packed
    # all-reduce INSIDE one CUDA graph, label counts all-reduced once for the "epoch"; 4 steps (eager warm-up, capture,
    # 2 replays) must leave identical parameters on every rank
    batch = {"agent": full.agent.contiguous(), "goal": full.goal.contiguous(), "hits": full.hits.contiguous(),
             "safe": safe.to(torch.uint8).contiguous(), "unsafe": unsafe.to(torch.uint8).contiguous()}
    mb = B // world
    per_graph = full.row_deg.reshape(B, N).sum(dim=1)
    runner = T.MinibatchRunner(algo, batch, mb, int(per_graph.max().item()) * mb, u_qp.contiguous())
    # every rank trains on its shard [lo, hi) of the global minibatch; global counts = sum over the ranks' shards
    den = T._minibatch_counts(batch, torch.arange(B, device=dev), np.array([lo, hi]))
    torch.distributed.all_reduce(den)
    before = algo.cbf_params.flat.clone()
    for i in range(4):
        runner.run(torch.arange(lo, hi, device=dev), den[0])
        torch.cuda.synchronize()
        _say(rank, f"captured-step runner call {i} done")
    runner.graph.check_overflow()
    p2 = torch.cat([algo.cbf_params.flat, algo.actor_net_params.flat])
    p2max, p2min = p2.clone(), p2.clone()
    torch.distributed.all_reduce(p2max, op=torch.distributed.ReduceOp.MAX)
    torch.distributed.all_reduce(p2min, op=torch.distributed.ReduceOp.MIN)
    graph_ok = bool(torch.equal(p2max, p2min)) and bool(torch.isfinite(p2).all()) and not torch.equal(before, algo.cbf_params.flat)
    graph_ok = graph_ok and runner.cuda_graph is not None
    if rank == 0:
        json.dump({"world": world, "grad_err": err, "grad_max": gmax, "stats_err": stats_err, "params_identical": same,
                   "captured_step_ok": graph_ok}, open(out_path, "w"))
    # a captured CUDA graph that contains NCCL kernels must be gone before the communicator is torn down
    # (destroy_process_group hung with the runner alive: measured on 2 GPUs)
    del runner
    import gc
    gc.collect()
    torch.cuda.synchronize()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
