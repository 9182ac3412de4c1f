"""Oracle: obstacles, LiDAR ray casting, inside tests (TEST INFRASTRUCTURE ONLY).

Restates gcbfplus/env/obstacle.py (Rectangle :25-96, Sphere :225-270) and
gcbfplus/env/utils.py (get_lidar :49-79, inside_obstacles :82-107,
raytracing :110-131) with torch-CPU tensors.  Every arithmetic step is a
separate torch op (one IEEE rounding per op, no FMA contraction) in the order
the reference writes it, so that the CUDA geometry kernels -- compiled with
``-fmad=false`` -- can be compared bit-for-bit on identical inputs.

Two deliberate, documented conventions (shared with the product):
  * trig of the *fixed* ray angles and of the obstacle angle theta is
    evaluated once on the host (NumPy float32) and passed around as tables /
    obstacle fields, instead of being re-evaluated per call (same values the
    reference recomputes every call; avoids libm-vs-CUDA ulp noise);
  * python-float constants are rounded to the working dtype exactly where JAX's
    weak typing would round them (e.g. ``comm_radius - 1e-1`` is computed in
    double, then cast).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

NO_HIT = 1e6  # env/obstacle.py:94, env/utils.py:125


def exact_sqrt(x: torch.Tensor) -> torch.Tensor:
    """Correctly rounded sqrt.  torch.sqrt on CPU fp32 is NOT correctly rounded (~0.6% of
    inputs are 1 ulp off, measured in this image), NumPy's is (hardware sqrtps) -- and so are
    XLA-CPU's and CUDA's sqrtf.  Falls back to torch.sqrt when autograd is needed."""
    if x.requires_grad:
        return torch.sqrt(x)
    return torch.from_numpy(np.sqrt(x.detach().numpy()))


# --------------------------------------------------------------------------- obstacles
@dataclass
class Rectangle:
    """Stacked 2-D rectangles, leading dim O (obstacle.py:25-51)."""
    center: torch.Tensor   # [O,2]
    width: torch.Tensor    # [O]
    height: torch.Tensor   # [O]
    theta: torch.Tensor    # [O]
    points: torch.Tensor   # [O,4,2]
    cos: torch.Tensor      # [O]  cos(theta), host-evaluated
    sin: torch.Tensor      # [O]

    @property
    def n(self) -> int:
        return int(self.center.shape[0])

    @staticmethod
    def create(center, width, height, theta, dtype=torch.float32) -> "Rectangle":
        """obstacle.py:34-51.  points = (rot @ bbox + center).T, corner order
        (+w,+h), (-w,+h), (-w,-h), (+w,-h)."""
        npdt = np.float32 if dtype == torch.float32 else np.float64
        center = np.asarray(center, dtype=npdt).reshape(-1, 2)
        width = np.asarray(width, dtype=npdt).reshape(-1)
        height = np.asarray(height, dtype=npdt).reshape(-1)
        theta = np.asarray(theta, dtype=npdt).reshape(-1)
        c, s = np.cos(theta).astype(npdt), np.sin(theta).astype(npdt)
        hw, hh = width / npdt(2), height / npdt(2)
        bx = np.stack([hw, -hw, -hw, hw], axis=1)   # [O,4]
        by = np.stack([hh, hh, -hh, -hh], axis=1)
        px = (c[:, None] * bx + (-s)[:, None] * by) + center[:, 0:1]
        py = (s[:, None] * bx + c[:, None] * by) + center[:, 1:2]
        points = np.stack([px, py], axis=-1).astype(npdt)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        return Rectangle(t(center), t(width), t(height), t(theta), t(points), t(c), t(s))

    def to(self, dtype) -> "Rectangle":
        return Rectangle(*(getattr(self, f).to(dtype) for f in
                           ("center", "width", "height", "theta", "points", "cos", "sin")))

    def inside(self, point: torch.Tensor, r: float = 0.0) -> torch.Tensor:
        """obstacle.py:53-63.  point [...,2] -> bool [..., O]."""
        rr = torch.tensor(r, dtype=point.dtype)
        rel_x = point[..., None, 0] - self.center[:, 0]
        rel_y = point[..., None, 1] - self.center[:, 1]
        rel_xx = torch.abs(rel_x * self.cos + rel_y * self.sin) - self.width / 2
        rel_yy = torch.abs(rel_x * self.sin - rel_y * self.cos) - self.height / 2
        is_in_down = (rel_xx < rr) & (rel_yy < 0)
        is_in_up = (rel_xx < 0) & (rel_yy < rr)
        is_out_corner = (rel_xx > 0) & (rel_yy > 0)
        is_in_circle = exact_sqrt(rel_xx * rel_xx + rel_yy * rel_yy) < rr
        return (is_in_down | is_in_up) | (is_out_corner & is_in_circle)

    def raytracing(self, start: torch.Tensor, end: torch.Tensor) -> torch.Tensor:
        """obstacle.py:65-96.  start/end [...,2] -> alpha [..., O] (min over 4 edges)."""
        x1, y1 = start[..., None, None, 0], start[..., None, None, 1]
        x2, y2 = end[..., None, None, 0], end[..., None, None, 1]
        x3, y3 = self.points[:, :, 0], self.points[:, :, 1]            # [O,4]
        prev = [3, 0, 1, 2]                                            # points[[-1,0,1,2]]
        x4, y4 = self.points[:, prev, 0], self.points[:, prev, 1]
        det = (x1 - x2) * (y4 - y3) - (y1 - y2) * (x4 - x3)
        det = torch.sign(det) * torch.clamp(torch.abs(det), 1e-7, 1e7)
        alphas = ((y4 - y3) * (x1 - x3) - (x4 - x3) * (y1 - y3)) / det
        betas = (-(y1 - y2) * (x1 - x3) + (x1 - x2) * (y1 - y3)) / det
        valids = ((alphas <= 1) & (alphas >= 0)) & ((betas <= 1) & (betas >= 0))
        v = valids.to(alphas.dtype)
        alphas = v * alphas + (1 - v) * NO_HIT
        return _nanmin(alphas, dim=-1)


@dataclass
class Sphere:
    """Stacked spheres, leading dim O (obstacle.py:225-232)."""
    center: torch.Tensor   # [O,3]
    radius: torch.Tensor   # [O]

    @property
    def n(self) -> int:
        return int(self.center.shape[0])

    @staticmethod
    def create(center, radius, dtype=torch.float32) -> "Sphere":
        c = torch.as_tensor(np.asarray(center), dtype=dtype).reshape(-1, 3)
        r = torch.as_tensor(np.asarray(radius), dtype=dtype).reshape(-1)
        return Sphere(c, r)

    def to(self, dtype) -> "Sphere":
        return Sphere(self.center.to(dtype), self.radius.to(dtype))

    def inside(self, point: torch.Tensor, r: float = 0.0) -> torch.Tensor:
        """obstacle.py:234-235: ||p - c|| <= radius + r.  point [...,3] -> [..., O]."""
        d = point[..., None, :] - self.center
        nrm = exact_sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
        return nrm <= self.radius + torch.tensor(r, dtype=point.dtype)

    def raytracing(self, start: torch.Tensor, end: torch.Tensor) -> torch.Tensor:
        """obstacle.py:237-270.  start/end [...,3] -> alpha [..., O]."""
        x1, y1, z1 = (start[..., None, k] for k in range(3))
        x2, y2, z2 = (end[..., None, k] for k in range(3))
        xc, yc, zc = (self.center[:, k] for k in range(3))
        r = self.radius
        dx, dy, dz = x2 - x1, y2 - y1, z2 - z1
        rmax = exact_sqrt(dx * dx + dy * dy + dz * dz)
        A = rmax * rmax
        B = 2 * (dx * (x1 - xc) + dy * (y1 - yc) + dz * (z1 - zc))
        C = (x1 - xc) * (x1 - xc) + (y1 - yc) * (y1 - yc) + (z1 - zc) * (z1 - zc) - r * r
        delta = B * B - 4 * A * C
        valid1 = (delta >= 0).to(delta.dtype)
        sq = exact_sqrt(delta * valid1)
        alpha1 = (-B - sq) / (2 * A) * valid1 + (1 - valid1)
        alpha2 = (-B + sq) / (2 * A) * valid1 + (1 - valid1)
        a1 = (alpha1 >= 0).to(delta.dtype) * alpha1 + (alpha1 < 0).to(delta.dtype) * 1
        a2 = (alpha2 >= 0).to(delta.dtype) * alpha2 + (alpha2 < 0).to(delta.dtype) * 1
        alphas = torch.minimum(a1, a2)
        alphas = torch.clamp(alphas, 0, 1)
        return valid1 * alphas + (1 - valid1) * NO_HIT


def _nanmin(x: torch.Tensor, dim: int) -> torch.Tensor:
    """jnp.min semantics: NaN-propagating minimum."""
    nan = torch.isnan(x)
    m = torch.min(torch.where(nan, torch.full_like(x, float("inf")), x), dim=dim).values
    return torch.where(nan.any(dim=dim), torch.full_like(m, float("nan")), m)


# --------------------------------------------------------------------------- ray tables
def ray_table_2d(num_beams: int, sense_range: float, dtype=torch.float32) -> torch.Tensor:
    """env/utils.py:51-56: thetas = linspace(-pi, pi - 2pi/n, n); returns
    [n,2] = (cos(theta)*range, sin(theta)*range), host-evaluated in `dtype`."""
    npdt = np.float32 if dtype == torch.float32 else np.float64
    thetas = np.linspace(-np.pi, np.pi - 2 * np.pi / num_beams, num_beams).astype(npdt)
    rng = npdt(sense_range)
    d = np.stack([np.cos(thetas).astype(npdt) * rng, np.sin(thetas).astype(npdt) * rng], axis=-1)
    return torch.from_numpy(d.astype(npdt))


def ray_table_3d(num_beams: int, sense_range: float, dtype=torch.float32) -> torch.Tensor:
    """env/utils.py:57-74: (n/2) thetas x n phis (theta-major) + the two poles.
    Returns [(n/2)*n + 2, 3] direction*range vectors."""
    npdt = np.float32 if dtype == torch.float32 else np.float64
    thetas = np.linspace(-np.pi / 2 + 2 * np.pi / num_beams, np.pi / 2 - 2 * np.pi / num_beams,
                         num_beams // 2).astype(npdt)
    phis = np.linspace(-np.pi, np.pi - 2 * np.pi / num_beams, num_beams).astype(npdt)
    rng = npdt(sense_range)
    ct, st = np.cos(thetas).astype(npdt), np.sin(thetas).astype(npdt)
    cp, sp = np.cos(phis).astype(npdt), np.sin(phis).astype(npdt)
    dx = (ct[:, None] * cp[None, :]) * rng
    dy = (ct[:, None] * sp[None, :]) * rng
    dz = np.broadcast_to((st * rng)[:, None], dx.shape)
    d = np.stack([dx, dy, dz], axis=-1).reshape(-1, 3)
    poles = np.array([[0, 0, rng], [0, 0, -rng]], dtype=npdt)
    return torch.from_numpy(np.concatenate([d, poles], axis=0).astype(npdt))


# --------------------------------------------------------------------------- lidar
def inside_obstacles(points: torch.Tensor, obstacles, r: float = 0.0) -> torch.Tensor:
    """env/utils.py:82-107.  points [n,dim] -> bool [n] (any obstacle)."""
    if obstacles is None or obstacles.n == 0:
        return torch.zeros(points.shape[:-1], dtype=torch.bool)
    return obstacles.inside(points, r).any(dim=-1)


def raytracing(starts: torch.Tensor, ends: torch.Tensor, obstacles, max_returns: int):
    """env/utils.py:110-131, batched over leading dims: starts/ends [..., n_rays, dim].
    Returns (hit points [..., min(max_returns,n_rays), dim], sorted alphas, order)."""
    if obstacles is None or obstacles.n == 0:
        alphas = torch.ones(starts.shape[:-1], dtype=starts.dtype) * NO_HIT
    else:
        is_in = inside_obstacles(starts, obstacles)                       # r = 0
        alphas = _nanmin(obstacles.raytracing(starts, ends), dim=-1)      # min over obstacles
        alphas = alphas * (1 - is_in.to(alphas.dtype))
    order = torch.argsort(alphas, dim=-1, stable=True)[..., :max_returns]  # NaNs sort last
    hitting = starts + (ends - starts) * alphas[..., None]
    hit_sorted = torch.gather(hitting, -2, order[..., None].expand(*order.shape, starts.shape[-1]))
    return hit_sorted, torch.gather(alphas, -1, order), order


def get_lidar(start_point: torch.Tensor, obstacles, ray_table: torch.Tensor, max_returns: int = 32):
    """env/utils.py:49-79.  start_point [..., dim] -> hits [..., R, dim];
    ray_table from ray_table_2d/3d."""
    starts = start_point[..., None, :].expand(*start_point.shape[:-1], ray_table.shape[0],
                                              start_point.shape[-1]).contiguous()
    ends = starts + ray_table
    return raytracing(starts, ends, obstacles, max_returns)[0]


def get_lidar_all(agent_pos: torch.Tensor, obstacles, ray_table: torch.Tensor, max_returns: int = 32):
    """vmap of get_lidar over agents: [N,dim] -> [N, R, dim]."""
    return get_lidar(agent_pos, obstacles, ray_table, max_returns)
