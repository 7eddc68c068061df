"""Ray / pose math of utils/ray_utils.py (torch ops).  Inside LocalTensorfs.forward the fused kernel
generates rays itself; these functions serve callers that need the directions on their own (flow
and depth losses in train.py) and the unit tests.

    contract                    ray_utils.py:9-12
    get_ray_directions_lean     ray_utils.py:14-24
    get_ray_directions_360      ray_utils.py:26-37
    get_rays_lean               ray_utils.py:39-53
"""
import math

import torch


def contract(x):
    """L-infinity scene contraction: identity inside the unit cube, (2 - 1/n) * x/n outside."""
    n = x.abs().amax(dim=-1, keepdim=True).clamp(min=1e-6)
    return torch.where(n <= 1, x, ((2 * n - 1) / (n ** 2)) * x)


def get_ray_directions_lean(i, j, focal, center):
    x = (i.float() + 0.5 - center[0]) / focal
    y = -(j.float() + 0.5 - center[1]) / focal
    return torch.stack([x, y, -torch.ones_like(x)], dim=-1)


def sphere2xyz(r, theta, phi):
    return torch.stack([r * phi.cos() * theta.sin(), r * phi.sin(), r * phi.cos() * theta.cos()], dim=-1)


def get_ray_directions_360(i, j, W, H):
    phi = (j.float() + 0.5) * math.pi / H - math.pi / 2.0
    theta = (i.float() + 0.5) * 2.0 * math.pi / W + math.pi
    return sphere2xyz(torch.ones_like(theta), theta, phi)


def get_rays_lean(directions, c2w):
    rays_o = c2w[:, :3, 3]
    rays_d = torch.bmm(c2w[:, :3, :3], directions[..., None])[..., 0]
    return rays_o, rays_d
