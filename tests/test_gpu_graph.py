"""GPU parity: graph build (LiDAR + radius neighbour lists) and label masks vs the CPU oracle.
Bar: bit-exact hit points, identical neighbour / active-hit index sets, identical masks
(geometry.cu is compiled with -fmad=false; same op order as oracle/geometry.py)."""
import numpy as np
import pytest
import torch

from helpers import (ENVS, edge_sets_oracle, edge_sets_product, oracle_env, oracle_obstacles, product_env,
                     product_obstacles, random_scene)

pytestmark = pytest.mark.gpu

CASES = [  # env, N, G, area, n_obs, n_rays
    ("SingleIntegrator", 8, 3, 2.0, 4, None),
    ("DoubleIntegrator", 8, 4, 2.0, 8, None),
    ("DoubleIntegrator", 61, 2, 4.0, 8, None),      # N not a multiple of the CTA tile
    ("DoubleIntegrator", 16, 2, 2.5, 0, None),      # zero obstacles
    ("DoubleIntegrator", 16, 2, 2.5, 6, 16),        # fewer rays than a warp
    ("DubinsCar", 12, 3, 2.5, 6, None),
    ("LinearDrone", 10, 2, 1.5, 4, None),
    ("LinearDrone", 33, 2, 2.0, 0, None),
]


def _build(env_id, N, G, area, n_obs, n_rays, seed):
    agent, goal, obs = random_scene(env_id, N, G, area, n_obs, seed)
    env = product_env(env_id, N, area, n_obs, n_rays)
    env.edge_cap_per_agent = 64
    pobs = product_obstacles(env_id, obs)
    graph = env.get_graph(torch.from_numpy(agent).cuda(), torch.from_numpy(goal).cuda(), pobs)
    torch.cuda.synchronize()
    graph.check_overflow()
    return env, graph, agent, goal, pobs


@pytest.mark.parametrize("env_id,N,G,area,n_obs,n_rays", CASES)
def test_graph_build_matches_oracle(env_id, N, G, area, n_obs, n_rays):
    env, graph, agent, goal, pobs = _build(env_id, N, G, area, n_obs, n_rays, seed=1)
    oenv = oracle_env(env_id, N, area, n_obs, n_rays)
    packed = pobs.packed.cpu().numpy()
    hits = graph.hits.cpu().numpy()
    n_edges = 0
    for g in range(G):
        oobs = oracle_obstacles(packed[g])
        og_dense = oenv.get_graph(torch.from_numpy(agent[g]), torch.from_numpy(goal[g]), oobs)
        og = oenv.sparsify(og_dense)
        ohits = og_dense.states[2 * N:-1, : env.pos_dim].reshape(N, env.n_hits, env.pos_dim).numpy()
        np.testing.assert_array_equal(hits[g], ohits)          # bit-exact (NaN == NaN allowed)
        assert edge_sets_product(graph, g, N) == edge_sets_oracle(og, N, env.n_hits)
        n_edges += og.edges.shape[0]
    assert graph.n_edge == n_edges


@pytest.mark.parametrize("env_id", ENVS)
def test_masks_match_oracle(env_id):
    N, G, area, n_obs = (24, 3, 1.6, 6) if env_id != "LinearDrone" else (24, 3, 1.2, 4)
    env, graph, agent, goal, pobs = _build(env_id, N, G, area, n_obs, None, seed=5)
    oenv = oracle_env(env_id, N, area, n_obs)
    packed = pobs.packed.cpu().numpy()
    got = {k: getattr(env, k + "_mask")(graph).cpu().numpy() for k in ("unsafe", "collision", "finish", "safe")}
    n_unsafe = 0
    for g in range(G):
        og = oenv.get_graph(torch.from_numpy(agent[g]), torch.from_numpy(goal[g]), oracle_obstacles(packed[g]))
        for k in got:
            want = getattr(oenv, k + "_mask")(og).numpy()
            np.testing.assert_array_equal(got[k][g], want, err_msg=f"{k} mask, graph {g}")
        n_unsafe += int(oenv.unsafe_mask(og).sum())
    assert n_unsafe > 0, "scene too sparse to exercise the unsafe labels"


def test_topology_only_rebuild_and_overflow():
    env, graph, agent, goal, pobs = _build("DoubleIntegrator", 32, 2, 2.0, 4, None, seed=3)
    g2 = env.get_graph(graph.agent, graph.goal, pobs, hits=graph.hits.clone())
    torch.cuda.synchronize()
    for g in range(2):
        assert edge_sets_product(graph, g, 32) == edge_sets_product(g2, g, 32)
    env.edge_cap_per_agent = 1   # goal edge only -> must overflow and be reported, not silently dropped
    g3 = env.get_graph(graph.agent, graph.goal, pobs)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="overflow"):
        g3.check_overflow()


def test_safe_horizon_matches_oracle():
    import ctypes as C
    from gcbfplus_b200 import _lib
    from oracle.algo import safe_mask_horizon
    rng = np.random.default_rng(0)
    for (B, T, N, H) in [(3, 40, 5, 8), (2, 16, 4, 32), (1, 64, 3, 0)]:
        unsafe = (rng.uniform(size=(B, T, N)) < 0.07)
        u = torch.from_numpy(unsafe.astype(np.uint8)).cuda()
        s = torch.empty_like(u)
        _lib.check(_lib.load().gcbf_safe_horizon(_lib.ptr(u), _lib.ptr(s), B, T, N, H,
                                                 torch.cuda.current_stream().cuda_stream), "safe_horizon")
        want = np.stack([safe_mask_horizon(unsafe[b], H) for b in range(B)])
        np.testing.assert_array_equal(s.cpu().numpy().astype(bool), want)
