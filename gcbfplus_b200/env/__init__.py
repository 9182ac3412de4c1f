"""gcbfplus.env surface (gcbfplus/env/__init__.py:1-46)."""
from typing import Optional

from .base import MultiAgentEnv, RolloutResult, StepResult
from .double_integrator import DoubleIntegrator
from .dubins_car import DubinsCar
from .linear_drone import LinearDrone
from .single_integrator import SingleIntegrator


class CrazyFlie:  # gcbfplus/env/crazyflie.py -- out of scope
    PARAMS: dict = {}

    def __init__(self, *a, **k):
        raise NotImplementedError("CrazyFlie is outside the B200 hot-path scope (SURVEY.md section 2, row 16)")


ENV = {
    "SingleIntegrator": SingleIntegrator,
    "DoubleIntegrator": DoubleIntegrator,
    "LinearDrone": LinearDrone,
    "DubinsCar": DubinsCar,
    "CrazyFlie": CrazyFlie,
}

DEFAULT_MAX_STEP = 256


def make_env(env_id: str, num_agents: int, area_size: float = None, max_step: int = None,
             max_travel: Optional[float] = None, num_obs: Optional[int] = None, n_rays: Optional[int] = None,
             device: str = "cuda") -> MultiAgentEnv:
    """gcbfplus/env/__init__.py:23-46 (same kwargs; dt fixed to 0.03).  Unlike the reference the
    class-level PARAMS dict is copied, not mutated (documented deviation)."""
    assert env_id in ENV.keys(), f"Environment {env_id} not implemented."
    params = dict(ENV[env_id].PARAMS)
    max_step = DEFAULT_MAX_STEP if max_step is None else max_step
    if num_obs is not None:
        params["n_obs"] = num_obs
    if n_rays is not None:
        params["n_rays"] = n_rays
    return ENV[env_id](num_agents=num_agents, area_size=area_size, max_step=max_step, max_travel=max_travel, dt=0.03,
                       params=params, device=device)
