"""SingleIntegrator -- gcbfplus/env/single_integrator.py (state [x, y], action [vx, vy])."""
import numpy as np
import torch

from .base import MultiAgentEnv, lqr


class SingleIntegrator(MultiAgentEnv):
    AGENT, GOAL, OBS = 0, 1, 2
    ENV_ID = "SingleIntegrator"
    DIMS = (2, 2, 2, 2)
    PARAMS = {"car_radius": 0.05, "comm_radius": 0.5, "n_rays": 32, "obs_len_range": [0.1, 0.6], "n_obs": 8}

    def _setup_dynamics(self) -> None:
        """single_integrator.py:53-59."""
        sd = self.state_dim
        A = np.zeros((sd, sd), dtype=np.float32) * self._dt + np.eye(sd)
        B = np.array([[1.0, 0.0], [0.0, 1.0]]) * self._dt
        self._K = lqr(A, B, np.eye(sd) * 2, np.eye(self.action_dim))

    def _thresholds(self) -> dict:
        r = self.radius  # single_integrator.py:323-358
        return dict(unsafe_agent=r * 2, unsafe_obs=r, safe_agent=r * 2.5, safe_obs=r * 1.5)

    def state_lim(self, state=None):
        up = torch.ones(2) * float("inf")
        return -up, up

    def action_lim(self):
        up = torch.ones(2)
        return -up, up

    def control_affine_dyn(self, state: torch.Tensor):
        """single_integrator.py:231-238."""
        f = torch.zeros_like(state)
        g = torch.eye(state.shape[-1], device=state.device).expand(*state.shape[:-1], -1, -1)
        return f, g
