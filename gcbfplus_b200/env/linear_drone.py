"""LinearDrone -- gcbfplus/env/linear_drone.py (state [x,y,z,vx,vy,vz], action [ax,ay,az])."""
import numpy as np
import scipy.linalg
import torch

from .base import MultiAgentEnv, lqr


class LinearDrone(MultiAgentEnv):
    AGENT, GOAL, OBS = 0, 1, 2
    ENV_ID = "LinearDrone"
    DIMS = (6, 6, 3, 3)
    PARAMS = {"drone_radius": 0.05, "comm_radius": 0.5, "n_rays": 32, "obs_len_range": [0.15, 0.3], "n_obs": 4}

    def _setup_dynamics(self) -> None:
        """linear_drone.py:55-73."""
        sd, nu = self.state_dim, self.action_dim
        A = np.zeros((sd, sd))
        A[0, 3] = A[1, 4] = A[2, 5] = 1.0
        A[3, 3] = A[4, 4] = -1.1
        A[5, 5] = -6.0
        B = np.zeros((sd, nu))
        B[3, 0] = B[4, 1] = B[5, 2] = 10.0
        self._A, self._B = A, B
        self._K = lqr(scipy.linalg.expm(A * self._dt), B, np.diag([5e1, 5e1, 5e1, 1.0, 1.0, 1.0]), np.eye(nu))
        self.n_rays = 16  # consider top k rays (linear_drone.py:73)

    @property
    def n_hits(self) -> int:
        return 16

    def _thresholds(self) -> dict:
        r = self.radius  # linear_drone.py:346-383
        return dict(unsafe_agent=r * 2.5, unsafe_obs=r * 1.5, safe_agent=r * 4, safe_obs=r * 2)

    def state_lim(self, state=None):
        inf = float("inf")
        up = torch.tensor([inf, inf, inf, 0.5, 0.5, 0.5])
        return -up, up

    def action_lim(self):
        up = torch.ones(3)
        return -up, up

    def control_affine_dyn(self, state: torch.Tensor):
        """linear_drone.py:255-262."""
        A = torch.tensor(self._A, dtype=state.dtype, device=state.device)
        B = torch.tensor(self._B, dtype=state.dtype, device=state.device)
        return state @ A.T, B.expand(*state.shape[:-1], -1, -1)
