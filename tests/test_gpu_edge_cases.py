"""GPU edge cases through the C ABI: single-agent graphs (no agent-agent edge, no neighbour), zero obstacles,
N not a multiple of the warp / CTA tiles, one graph, and the QP label of a scene where no CBF constraint binds."""
import numpy as np
import pytest
import torch

from helpers import (edge_sets_oracle, edge_sets_product, oracle_env, oracle_obstacles, oracle_params, product_algo,
                     product_env, product_obstacles, random_scene)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env_id,N,G,n_obs", [("DoubleIntegrator", 1, 3, 2), ("SingleIntegrator", 1, 1, 0),
                                              ("LinearDrone", 1, 2, 1), ("DubinsCar", 3, 1, 0),
                                              ("DoubleIntegrator", 33, 1, 0)])
def test_tiny_and_ragged_swarms(env_id, N, G, n_obs, gemm_path):
    from oracle.algo import act, get_cbf
    area = 1.5
    agent, goal, obs = random_scene(env_id, N, G, area, max(n_obs, 1), seed=13)
    env = product_env(env_id, N, area, n_obs)
    env.edge_cap_per_agent = 64
    pobs = product_obstacles(env_id, {k: v[:, :n_obs] for k, v in obs.items()}) if n_obs > 0 else None
    graph = env.get_graph(torch.from_numpy(agent).cuda(), torch.from_numpy(goal).cuda(), pobs)
    algo = product_algo(env, env_id)
    h = algo.get_cbf(graph)
    a = algo.act(graph)
    nxt = env.step(graph, a)
    u_qp = algo.get_b_u_qp(graph)
    torch.cuda.synchronize()
    graph.check_overflow()
    oenv = oracle_env(env_id, N, area, n_obs)
    ap, cp = oracle_params(env_id)
    tol = 1e-5 if gemm_path == "simt" else 3e-5
    with torch.no_grad():
        for g in range(G):
            oobs = oracle_obstacles(pobs.packed.cpu().numpy()[g]) if n_obs > 0 else None
            og = oenv.sparsify(oenv.get_graph(torch.from_numpy(agent[g]), torch.from_numpy(goal[g]), oobs))
            assert edge_sets_product(graph, g, N) == edge_sets_oracle(og, N, env.n_hits)
            np.testing.assert_allclose(h[g].cpu().numpy(), get_cbf(cp, og).numpy(), atol=tol, rtol=0)
            np.testing.assert_allclose(a[g].cpu().numpy(), act(oenv, ap, og).numpy(), atol=2 * tol + 1e-5, rtol=0)
            og2, r, c = oenv.step(og, torch.from_numpy(a[g].cpu().numpy()))
            np.testing.assert_allclose(nxt.graph.agent[g].cpu().numpy(), og2.agent.numpy(), atol=1e-6, rtol=0)
    assert bool(torch.isfinite(u_qp).all()) and u_qp.shape == (G, N, env.action_dim)


def test_qp_label_equals_clipped_u_ref_when_no_constraint_binds():
    """Agents far apart, no obstacles, at rest next to their goals: h > 0, Lf_h = 0 and Lg_h u_ref ~ 1e-3, so every
    CBF row -Lg_h u - r <= 0.1 alpha h is slack at u = u_ref; the unique minimiser is clip(u_ref) and the multipliers
    are exactly zero (gcbf_plus.py:329-339 with an inactive C x <= b)."""
    from gcbfplus_b200.algo.train import batch_u_ref, qp_labels
    env_id, N, G, area = "DoubleIntegrator", 6, 4, 12.0
    rng = np.random.Generator(np.random.PCG64(0))
    agent = np.zeros((G, N, 4), np.float32)
    goal = np.zeros((G, N, 4), np.float32)
    grid = np.stack(np.meshgrid(np.arange(3), np.arange(2)), -1).reshape(-1, 2).astype(np.float32) * 4.0 + 1.0
    agent[..., :2] = grid[None]
    goal[..., :2] = grid[None] + rng.uniform(-1e-4, 1e-4, size=(G, N, 2)).astype(np.float32)   # u_ref ~ 1e-4
    env = product_env(env_id, N, area, 0)
    algo = product_algo(env, env_id)
    graph = env.get_graph(torch.from_numpy(agent).cuda(), torch.from_numpy(goal).cuda(), None)
    u, aux, iters = qp_labels(algo, graph, params=algo.cbf_params, with_aux=True)
    h = algo.get_cbf(graph)
    torch.cuda.synchronize()
    assert float(h.min()) > 0.05
    u_ref = batch_u_ref(algo, {"agent": graph.agent, "goal": graph.goal}).clamp(-1, 1)
    assert float(aux[..., 0].abs().max()) == 0.0 and float(aux[..., 1].abs().max()) == 0.0
    assert torch.equal(u, u_ref)
    assert int(iters.max()) <= 2


@pytest.mark.parametrize("env_id", ["DoubleIntegrator", "LinearDrone"])
def test_inside_obstacles_is_independent_of_agent_collisions(env_id):
    """env.step(get_eval_info=True)['inside_obstacles'] = inside_obstacles(agent_pos, obstacles, r) alone
    (double_integrator.py:172-175): an agent that sits in an obstacle AND touches another agent is still inside
    (the mask used to be collision & ~agent_collision)."""
    from oracle.geometry import inside_obstacles
    N, G, area, n_obs = 12, 3, 1.2, 3
    agent, goal, obs = random_scene(env_id, N, G, area, n_obs, seed=5)
    pd = 3 if env_id == "LinearDrone" else 2
    agent[:, 0, :pd] = obs["center"][:, 0]                        # agent 0: at the centre of obstacle 0 ...
    agent[:, 1, :pd] = obs["center"][:, 0] + 0.01                 # ... and 1 cm from agent 1 (colliding pair)
    env = product_env(env_id, N, area, n_obs)
    pobs = product_obstacles(env_id, obs)
    graph = env.get_graph(torch.from_numpy(agent).cuda(), torch.from_numpy(goal).cuda(), pobs)
    got = env.inside_obstacles(graph).cpu().numpy()
    info = env.step(graph, env.u_ref(graph), get_eval_info=True).info["inside_obstacles"].cpu().numpy()
    torch.cuda.synchronize()
    assert got[:, 0].all() and got[:, 1].all()
    packed = pobs.packed.cpu().numpy()
    for g in range(G):
        want = inside_obstacles(torch.from_numpy(agent[g, :, :pd]), oracle_obstacles(packed[g]), r=0.05).numpy()
        np.testing.assert_array_equal(got[g], want)
        np.testing.assert_array_equal(info[g], want)
