"""DoubleIntegrator -- gcbfplus/env/double_integrator.py (state [x, y, vx, vy], action [fx, fy])."""
import numpy as np
import torch

from .base import MultiAgentEnv, lqr


class DoubleIntegrator(MultiAgentEnv):
    AGENT, GOAL, OBS = 0, 1, 2
    ENV_ID = "DoubleIntegrator"
    DIMS = (4, 4, 2, 2)
    PARAMS = {"car_radius": 0.05, "comm_radius": 0.5, "n_rays": 32, "obs_len_range": [0.1, 0.5], "n_obs": 8,
              "m": 0.1}

    def _setup_dynamics(self) -> None:
        """double_integrator.py:53-66."""
        sd = self.state_dim
        A = np.zeros((sd, sd), dtype=np.float32)
        A[0, 2] = 1.0
        A[1, 3] = 1.0
        A = A * self._dt + np.eye(sd)
        m = self._params["m"]
        B = np.array([[0.0, 0.0], [0.0, 0.0], [1.0 / m, 0.0], [0.0, 1.0 / m]]) * self._dt
        self._K = lqr(A, B, np.eye(sd) * 5, np.eye(self.action_dim))

    def state_lim(self, state=None):
        up = torch.tensor([float("inf"), float("inf"), 0.5, 0.5])
        return -up, up

    def action_lim(self):
        up = torch.ones(2)
        return -up, up

    def control_affine_dyn(self, state: torch.Tensor):
        """double_integrator.py:266-273."""
        f = torch.cat([state[..., 2:], torch.zeros_like(state[..., :2])], dim=-1)
        g = torch.cat([torch.zeros(2, 2), torch.eye(2) / self._params["m"]], dim=0).to(state.device)
        return f, g.expand(*state.shape[:-1], -1, -1)
