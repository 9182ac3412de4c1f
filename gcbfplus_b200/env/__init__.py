"""Environment registry and factory with the reference's names (gcbfplus/env/__init__.py): `ENV`, `DEFAULT_MAX_STEP`,
`make_env(env_id, num_agents, area_size, max_step, max_travel, num_obs, n_rays)` plus a `device` argument."""
from typing import Optional

from .base import MultiAgentEnv, RolloutResult, StepResult
from .double_integrator import DoubleIntegrator
from .dubins_car import DubinsCar
from .linear_drone import LinearDrone
from .single_integrator import SingleIntegrator

DEFAULT_MAX_STEP = 256     # episode length T of train.py / test.py
_DT = 0.03                 # integration step every environment is created with


class CrazyFlie:
    """Placeholder: the CrazyFlie dynamics need the `control` package and are outside the hot path."""
    PARAMS: dict = {}

    def __init__(self, *unused, **also_unused):
        raise NotImplementedError("CrazyFlie is outside the B200 hot-path scope (SURVEY.md section 2, row 16)")


ENV = {cls.__name__: cls for cls in (SingleIntegrator, DoubleIntegrator, LinearDrone, DubinsCar, CrazyFlie)}


def make_env(env_id: str, num_agents: int, area_size: float = None, max_step: int = None,
             max_travel: Optional[float] = None, num_obs: Optional[int] = None, n_rays: Optional[int] = None,
             device: str = "cuda") -> MultiAgentEnv:
    """Build `env_id` with per-call overrides of the obstacle count and ray count.  The class-level PARAMS dict is
    copied, never mutated (the reference mutates it, gcbfplus/env/__init__.py:36-40: documented deviation)."""
    if env_id not in ENV:
        raise AssertionError(f"Environment {env_id} not implemented.")
    cls = ENV[env_id]
    overrides = {"n_obs": num_obs, "n_rays": n_rays}
    params = {**cls.PARAMS, **{k: v for k, v in overrides.items() if v is not None}}
    return cls(num_agents=num_agents, area_size=area_size, max_step=max_step or DEFAULT_MAX_STEP,
               max_travel=max_travel, dt=_DT, params=params, device=device)
