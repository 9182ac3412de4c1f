"""GPU parity: device reset (gcbf_reset_positions, one warp per environment, threefry key chain) against the host
NumPy restatement (env/base.py `_sample_agents_goals`, itself bit-exact against the scalar oracle in
tests/test_oracle.py).  Bar: bit-exact start / goal states and therefore identical graphs."""
import numpy as np
import pytest
import torch

from helpers import product_env

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env_id,N,E,area,n_obs,max_travel", [
    ("SingleIntegrator", 8, 5, 1.6, 3, None), ("DoubleIntegrator", 24, 4, 2.6, 8, None),
    ("DoubleIntegrator", 6, 3, 3.0, 4, 1.0), ("DubinsCar", 9, 3, 2.2, 4, None), ("LinearDrone", 12, 4, 1.4, 4, None),
    ("DoubleIntegrator", 512, 2, 32.0, 8, None)])
@pytest.mark.parametrize("partitionable", [False, True])
def test_device_reset_matches_host_sampler(env_id, N, E, area, n_obs, max_travel, partitionable, monkeypatch):
    """Both threefry stream layouts (jax_threefry_partitionable off / on, utils/jrandom.py)."""
    from gcbfplus_b200.env import make_env
    from gcbfplus_b200.utils import jrandom as jr
    monkeypatch.setattr(jr, "PARTITIONABLE", partitionable)
    keys = jr.split(jr.PRNGKey(17), E)
    out = []
    for host in (False, True):
        env = make_env(env_id, N, area_size=area, num_obs=n_obs, max_travel=max_travel)
        env.host_reset = host
        g = env.reset(keys)
        torch.cuda.synchronize()
        out.append((g.agent.cpu().numpy(), g.goal.cpu().numpy(), g.obstacle.packed.cpu().numpy(), g.hits.cpu().numpy()))
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)
    agent = out[0][0]
    pd = 3 if env_id == "LinearDrone" else 2
    for e in range(E):       # accepted samples respect the 4 r spacing
        d = np.linalg.norm(agent[e][:, None, :pd] - agent[e][None, :, :pd], axis=-1) + np.eye(N) * 10
        assert d.min() > 4 * 0.05
