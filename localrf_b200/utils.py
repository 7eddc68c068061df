"""Small pose / schedule helpers on the render path's boundary.

    sixD_to_mtx, mtx_to_sixD   utils/utils.py:381-392
    N_to_reso                  utils/utils.py:200-203
    TVLoss                     utils/utils.py:293-312
"""
import torch


def sixD_to_mtx(r):
    """[..., 3, 2] continuous 6-D rotation -> [..., 3, 3] by Gram-Schmidt (columns b1,b2,b3).

    The reference calls torch.cross without `dim`, which for a batch of exactly 3 views crosses
    over the batch axis; the intended per-view cross product (what every other batch size gets)
    is used here.
    """
    a1, a2 = r[..., 0], r[..., 1]
    b1 = a1 / torch.norm(a1, dim=-1, keepdim=True)
    b2 = a2 - torch.sum(b1 * a2, dim=-1, keepdim=True) * b1
    b2 = b2 / torch.norm(b2, dim=-1, keepdim=True)
    b3 = torch.linalg.cross(b1, b2, dim=-1)
    return torch.stack([b1, b2, b3], dim=-1)


def mtx_to_sixD(r):
    return torch.stack([r[..., 0], r[..., 1]], dim=-1)


def N_to_reso(n_voxels, bbox):
    lo, hi = bbox
    voxel = ((hi - lo).prod() / n_voxels).pow(1 / 3)
    return ((hi - lo) / voxel).long().tolist()


class TVLoss(torch.nn.Module):
    """Total-variation regulariser over the last two axes (mean of squared neighbour differences)."""

    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        tv = 0
        if x.shape[2] > 1:
            tv = tv + (x[:, :, 1:, :] - x[:, :, :-1, :]).pow(2).mean()
        if x.shape[3] > 1:
            tv = tv + (x[:, :, :, 1:] - x[:, :, :, :-1]).pow(2).mean()
        return self.TVLoss_weight * 2 * tv
