"""Obstacle containers (host side) -- Rectangle / Sphere of gcbfplus/env/obstacle.py:25-51,
225-232, batched over graphs and packed for the CUDA kernels (include/gcbf_b200.h):
Rectangle -> [G, O, 16] = cx, cy, w/2, h/2, cos, sin, p0x, p0y, ..., p3x, p3y, 0, 0
Sphere    -> [G, O, 4]  = cx, cy, cz, radius.
Creation is setup work (NumPy fp32 on the host), not the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Rectangle:
    packed: torch.Tensor          # [G, O, 16] fp32 (device)
    width: np.ndarray             # [G, O] host copies of the reference's fields
    height: np.ndarray
    theta: np.ndarray

    @staticmethod
    def create(center, width, height, theta, device="cuda") -> "Rectangle":
        """gcbfplus/env/obstacle.py:34-51: points = (rot @ bbox + center).T with corners
        (+w/2,+h/2), (-w/2,+h/2), (-w/2,-h/2), (+w/2,-h/2).  Inputs [G, O, ...]."""
        f = np.float32
        center = np.asarray(center, dtype=f)
        width = np.asarray(width, dtype=f)
        height = np.asarray(height, dtype=f)
        theta = np.asarray(theta, dtype=f)
        G, O = width.shape
        c, s = np.cos(theta).astype(f), np.sin(theta).astype(f)
        hw, hh = width / f(2), height / f(2)
        bx = np.stack([hw, -hw, -hw, hw], axis=-1)
        by = np.stack([hh, hh, -hh, -hh], axis=-1)
        px = (c[..., None] * bx + (-s)[..., None] * by) + center[..., 0:1]
        py = (s[..., None] * bx + c[..., None] * by) + center[..., 1:2]
        packed = np.zeros((G, O, 16), dtype=f)
        packed[..., 0:2] = center
        packed[..., 2], packed[..., 3], packed[..., 4], packed[..., 5] = hw, hh, c, s
        packed[..., 6:14:2] = px
        packed[..., 7:14:2] = py
        return Rectangle(torch.from_numpy(packed).to(device), width, height, theta)

    @property
    def center(self) -> torch.Tensor:
        return self.packed[..., 0:2]

    @property
    def points(self) -> torch.Tensor:
        return self.packed[..., 6:14].reshape(*self.packed.shape[:2], 4, 2)

    @property
    def n_obs(self) -> int:
        return int(self.packed.shape[1])

    def select(self, idx) -> "Rectangle":
        return Rectangle(self.packed[idx].contiguous(), self.width[idx], self.height[idx], self.theta[idx])

    def repeat(self, n: int) -> "Rectangle":
        """Tile the graph-batch dim n times ([G,...] -> [n*G,...], graph-major order t*G+g)."""
        return Rectangle(self.packed.repeat(n, 1, 1), np.tile(self.width, (n, 1)), np.tile(self.height, (n, 1)),
                         np.tile(self.theta, (n, 1)))


@dataclass
class Sphere:
    packed: torch.Tensor          # [G, O, 4] fp32 (device)

    @staticmethod
    def create(center, radius, device="cuda") -> "Sphere":
        """gcbfplus/env/obstacle.py:230-232.  center [G,O,3], radius [G,O]."""
        f = np.float32
        center = np.asarray(center, dtype=f)
        radius = np.asarray(radius, dtype=f)
        packed = np.concatenate([center, radius[..., None]], axis=-1).astype(f)
        return Sphere(torch.from_numpy(packed).to(device))

    @property
    def center(self) -> torch.Tensor:
        return self.packed[..., 0:3]

    @property
    def radius(self) -> torch.Tensor:
        return self.packed[..., 3]

    @property
    def n_obs(self) -> int:
        return int(self.packed.shape[1])

    def select(self, idx) -> "Sphere":
        return Sphere(self.packed[idx].contiguous())

    def repeat(self, n: int) -> "Sphere":
        return Sphere(self.packed.repeat(n, 1, 1))
