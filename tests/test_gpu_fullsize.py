"""GPU parity at BASELINE.json's FULL sizes (configs 3-5): DoubleIntegrator n=512 x 16 envs (8 obstacles, 32 rays),
DubinsCar n=256 x 32 envs (16 obstacles), LinearDrone n=1024 x 64 envs (514 rays -> 16 hits).

The dense reference formulation is O(N^2) per graph, so at these sizes the oracle is used where it stays in
seconds (one environment of the batch: LiDAR + index sets bit-exact, h / pi / next state within the single-step
tolerance on the oracle's SPARSIFIED graph -- dense == sparse is proven at small N in test_oracle.py /
test_gpu_gnn.py) and the whole batch is checked through size-independent properties:
  * neighbour lists == brute-force fp32 radius test for every agent of every env (symmetric, ascending, no self);
  * every row is [goal | agents ascending | active hits ascending], degrees sum to the edge counter;
  * batch invariance: an env run alone gives the bits it gives inside the batch (what env-sharding relies on);
  * rollout: CUDA-graph replay == eager bits; the two halves of the env batch run separately == the full batch;
  * train step: gradients of a minibatch == sum of the gradients of its two halves under the global
    denominators (the identity behind the single all-reduce), h_dot loss included.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import (edge_sets_oracle, edge_sets_product, oracle_env, oracle_obstacles, oracle_params, product_algo,
                     product_env, product_obstacles, random_scene)

pytestmark = pytest.mark.gpu

# env, N, E, area (density-preserving: sqrt(2N) in 2-D, N^(1/3) in 3-D), n_obs
FULL = [("DoubleIntegrator", 512, 16, 32.0, 8), ("DubinsCar", 256, 32, 22.63, 16), ("LinearDrone", 1024, 64, 10.08, 4)]
TOL = {"simt": 1e-5, "tc": 3e-5}


def _scene(env_id, N, E, area, n_obs, seed=3):
    agent, goal, obs = random_scene(env_id, N, E, area, n_obs, seed)
    env = product_env(env_id, N, area, n_obs)
    pobs = product_obstacles(env_id, obs)
    graph = env.get_graph(torch.from_numpy(agent).cuda(), torch.from_numpy(goal).cuda(), pobs)
    torch.cuda.synchronize()
    graph.check_overflow()
    return env, graph, agent, goal, pobs


@pytest.mark.parametrize("env_id,N,E,area,n_obs", FULL)
def test_graph_build_full_size(env_id, N, E, area, n_obs):
    env, graph, agent, goal, pobs = _scene(env_id, N, E, area, n_obs)
    pd = env.pos_dim
    rs, rd = graph.row_start.cpu().numpy(), graph.row_deg.cpu().numpy()
    src, recv = graph.edge_src.cpu().numpy(), graph.edge_recv.cpu().numpy()
    assert int(rd.sum()) == graph.n_edge
    f = np.float32
    R = f(env._params["comm_radius"])
    n_pairs = 0
    for e in range(E):
        p = agent[e, :, :pd].astype(f)
        acc = np.zeros((N, N), dtype=f)
        for c in range(pd):                      # same fp32 summation order as the kernel / jnp.linalg.norm
            dlt = p[:, None, c] - p[None, :, c]
            acc = acc + dlt * dlt
        nbr = (np.sqrt(acc) < R) & ~np.eye(N, dtype=bool)
        assert (nbr == nbr.T).all()
        for i in range(N):
            a = e * N + i
            codes = src[rs[a]: rs[a] + rd[a]]
            assert (recv[rs[a]: rs[a] + rd[a]] == a).all()
            assert codes[0] == -1                                    # goal edge first
            ag = codes[codes >= 0] - e * N
            assert (np.diff(ag) > 0).all() and np.array_equal(ag, np.nonzero(nbr[i])[0])
            ht = codes[codes <= -2]
            assert (np.diff(ht) < 0).all() and len(ag) + len(ht) + 1 == len(codes)     # hit k = -2-k ascending in k
            assert np.array_equal(codes, np.concatenate([[-1], ag + e * N, ht]))
            n_pairs += len(ag)
    assert n_pairs > 0
    # one environment against the oracle: hit points bit-exact, index sets identical
    packed = pobs.packed.cpu().numpy()
    oenv = oracle_env(env_id, N, area, n_obs)
    for e in (E - 1,):
        og_dense = oenv.get_graph(torch.from_numpy(agent[e]), torch.from_numpy(goal[e]), oracle_obstacles(packed[e]))
        ohits = og_dense.states[2 * N:-1, :pd].reshape(N, env.n_hits, pd).numpy()
        np.testing.assert_array_equal(graph.hits[e].cpu().numpy(), ohits)
        assert edge_sets_product(graph, e, N) == edge_sets_oracle(oenv.sparsify(og_dense), N, env.n_hits)
    # batch invariance
    solo = env.get_graph(graph.agent[E // 2: E // 2 + 1].contiguous(), graph.goal[E // 2: E // 2 + 1].contiguous(),
                         pobs.select(slice(E // 2, E // 2 + 1)))
    torch.cuda.synchronize()
    assert torch.equal(solo.hits[0], graph.hits[E // 2]) or \
        (torch.isnan(solo.hits[0]) == torch.isnan(graph.hits[E // 2])).all()
    assert edge_sets_product(solo, 0, N) == edge_sets_product(graph, E // 2, N)


@pytest.mark.parametrize("env_id,N,E,area,n_obs", FULL)
def test_forward_and_step_full_size(env_id, N, E, area, n_obs, gemm_path):
    from oracle.algo import act, get_cbf
    env, graph, agent, goal, pobs = _scene(env_id, N, E, area, n_obs)
    algo = product_algo(env, env_id)
    h = algo.get_cbf(graph)
    a = algo.act(graph)
    nxt = env.step(graph, a)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(h).all()) and bool(torch.isfinite(a).all())
    e = E - 1
    oenv = oracle_env(env_id, N, area, n_obs)
    ap, cp = oracle_params(env_id)
    packed = pobs.packed.cpu().numpy()
    with torch.no_grad():
        og = oenv.sparsify(oenv.get_graph(torch.from_numpy(agent[e]), torch.from_numpy(goal[e]),
                                          oracle_obstacles(packed[e])))
        np.testing.assert_allclose(h[e].cpu().numpy(), get_cbf(cp, og).numpy(), atol=TOL[gemm_path], rtol=0)
        np.testing.assert_allclose(a[e].cpu().numpy(), act(oenv, ap, og).numpy(), atol=2 * TOL[gemm_path] + 1e-5, rtol=0)
        og2, r, c = oenv.step(og, torch.from_numpy(a[e].cpu().numpy()))
        np.testing.assert_allclose(nxt.graph.agent[e].cpu().numpy(), og2.agent.numpy(), atol=1e-6, rtol=0)
        np.testing.assert_allclose(nxt.cost[e].item(), c.item(), atol=1e-6)
    # batch invariance of the network outputs
    solo = env.get_graph(graph.agent[2:3].contiguous(), graph.goal[2:3].contiguous(), pobs.select(slice(2, 3)))
    assert torch.equal(algo.get_cbf(solo)[0], h[2])
    assert torch.equal(algo.act(solo)[0], a[2])


@pytest.mark.parametrize("env_id,N,E,area,n_obs", FULL[:2])
def test_rollout_full_size_sharding_invariance(env_id, N, E, area, n_obs):
    """The env batch cut in two (what two ranks would run) reproduces the full batch bit for bit, and the captured
    CUDA graph replays the eager launch sequence exactly."""
    from gcbfplus_b200.trainer.rollout import RolloutEngine
    T = 24
    env, graph, agent, goal, pobs = _scene(env_id, N, E, area, n_obs)
    algo = product_algo(env, env_id)

    def run(lo, hi, use_graph=True):
        eng = RolloutEngine(env, hi - lo, T=T, n_obs=n_obs, use_cuda_graph=use_graph)
        eng.set_params(algo.actor_params)
        eng.set_initial(graph.agent[lo:hi].contiguous(), graph.goal[lo:hi].contiguous(), pobs.select(slice(lo, hi)))
        eng.run()
        torch.cuda.synchronize()
        return eng.agent.clone(), eng.rewards.clone(), eng.costs.clone()
    full = run(0, E)
    eager = run(0, E, use_graph=False)
    assert torch.equal(full[0], eager[0]) and torch.equal(full[1], eager[1])
    a, b = run(0, E // 2), run(E // 2, E)
    assert torch.equal(torch.cat([a[0], b[0]], dim=1), full[0])
    assert torch.equal(torch.cat([a[1], b[1]], dim=1), full[1])
    assert torch.equal(torch.cat([a[2], b[2]], dim=1), full[2])
    assert bool(torch.isfinite(full[0]).all())
    assert float((full[0][T] - full[0][0]).abs().max()) > 1e-2       # the swarm moved


def test_train_step_full_size_gradient_additivity(gemm_path):
    """Config 3 (DoubleIntegrator n=512, full train step with the h_dot loss): grad(minibatch) == grad(half A) +
    grad(half B) under the global denominators -- the identity that makes ONE gradient all-reduce exact."""
    from gcbfplus_b200 import _lib
    from gcbfplus_b200.algo.train import TrainState, qp_labels, train_minibatch
    env_id, N, B, area, n_obs = "DoubleIntegrator", 512, 32, 32.0, 8
    env, graph, agent, goal, pobs = _scene(env_id, N, B, area, n_obs, seed=9)
    algo = product_algo(env, env_id)
    unsafe = env.unsafe_mask(graph).bool().to(torch.uint8)
    safe = (env.safe_mask(graph).bool() & ~unsafe.bool()).to(torch.uint8)
    assert int(unsafe.sum()) > 0 and int(safe.sum()) > 0
    u_qp = qp_labels(algo, graph)

    def grads(lo, hi, denoms=None):
        g = env.get_graph(graph.agent[lo:hi].contiguous(), graph.goal[lo:hi].contiguous(), None,
                          hits=graph.hits[lo:hi].contiguous())
        algo._trainer_state = None
        ts = _train(algo, g, safe[lo:hi], unsafe[lo:hi], u_qp[lo:hi], denoms)
        torch.cuda.synchronize()
        return ts.packed.clone(), ts.denoms.clone()

    def _train(algo, g, sm, um, uq, denoms):
        if denoms is None:
            return train_minibatch(algo, g, sm, um, uq, apply=False)
        # same call with the denominators of the FULL minibatch injected (what the count all-reduce produces)
        env_ = algo._env
        algo._trainer_state = TrainState(algo)
        ts = algo._trainer_state
        Bn = g.n_graphs
        d = env_.desc(Bn, 0, edge_cap=g.edge_recv.numel())
        n = env_.lib.gcbf_train_workspace_floats(C.byref(d))
        ts.ws = torch.empty(int(n), dtype=torch.float32, device=env_.device)
        ts.denoms.copy_(denoms)
        hp = (C.c_float * 7)(algo.alpha, algo.eps, algo.loss_action_coef, algo.loss_unsafe_coef, algo.loss_safe_coef,
                             algo.loss_h_dot_coef, 1.0 if _lib.USE_TC else 0.0)
        sm = sm.reshape(-1).contiguous()
        um = um.reshape(-1).contiguous()
        uq = uq.reshape(Bn * N, -1).contiguous()
        rc = env_.lib.gcbf_train_step(C.byref(d), hp, _lib.ptr(algo.cbf_params.flat), _lib.ptr(algo.actor_net_params.flat),
                                      _lib.ptr(g.agent), _lib.ptr(g.goal), _lib.ptr(g.hits), _lib.ptr(g.row_start),
                                      _lib.ptr(g.row_deg), _lib.ptr(g.edge_recv), _lib.ptr(g.edge_src),
                                      _lib.ptr(g.counters), _lib.ptr(sm), _lib.ptr(um), _lib.ptr(uq), _lib.ptr(ts.denoms),
                                      _lib.ptr(ts.grad_cbf), _lib.ptr(ts.grad_act), _lib.ptr(ts.stats), _lib.ptr(ts.ws),
                                      ts.ws.numel(), env_._stream())
        _lib.check(rc, "gcbf_train_step")
        return ts
    full, den = grads(0, B)
    ga, _ = grads(0, B // 2, den)
    gb, _ = grads(B // 2, B, den)
    summed = ga + gb
    scale = float(full.abs().max())
    assert scale > 0
    err = float((summed - full).abs().max())
    assert err <= 2e-5 * scale + 1e-7, (err, scale)     # fp32 summation order of the dW reductions differs
    assert bool(torch.isfinite(full).all())
