"""Ray-batch data parallelism over the GPUs of one node.

The render path has no cross-ray term, so the field is replicated on every rank, each rank renders
a contiguous shard of the batch's rays and ONE all-gather of the rendered [rays, 4] pixels
(rgb + depth) closes the step.  No collective touches the data path itself.  The reference has no
distributed code at all (its only multi-GPU mode is one scene per process, scripts/train_all.sh).

Backend: NCCL over NVLink 5 / NVSwitch on GPUs, gloo in the CPU tests of this host logic.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_rays, rank, world, align=8):
    """Contiguous shard [lo, hi) of rank `rank`; boundaries aligned to `align` rays (the kernel's
    ray-tile size) so that every rank's launch covers whole tiles; the union is exactly
    [0, n_rays) and shards differ by at most one aligned block."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    blocks = (n_rays + align - 1) // align
    base, extra = divmod(blocks, world)
    lo_b = rank * base + min(rank, extra)
    hi_b = lo_b + base + (1 if rank < extra else 0)
    return min(lo_b * align, n_rays), min(hi_b * align, n_rays)


def shard_views(view_ids, n_rays, rank, world):
    """Shard of a training-style batch (rays grouped by view): whole views per rank, so the
    kernel's view indexing (ray // rays_per_view) stays valid inside the shard."""
    n_views = view_ids.shape[0]
    per_view = n_rays // n_views
    v_lo, v_hi = shard_bounds(n_views, rank, world, align=1)
    return v_lo, v_hi, v_lo * per_view, v_hi * per_view


def gather_pixels(rgb, depth, n_total, group=None):
    """One all-gather of the rendered pixels: local [n_local,3] + [n_local] -> [n_total,3], [n_total].

    Shards may differ in length by one aligned block, so every rank pads to the largest shard."""
    world = dist.get_world_size(group)
    if world == 1:
        return rgb, depth
    rank = dist.get_rank(group)
    bounds = [shard_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in bounds)
    pix = torch.zeros(width, 4, dtype=rgb.dtype, device=rgb.device)
    n_local = bounds[rank][1] - bounds[rank][0]
    pix[:n_local, :3] = rgb
    pix[:n_local, 3] = depth
    out = torch.empty(world * width, 4, dtype=rgb.dtype, device=rgb.device)
    dist.all_gather_into_tensor(out, pix, group=group)
    out = out.view(world, width, 4)
    parts = [out[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)]
    full = torch.cat(parts, dim=0)
    return full[:, :3].contiguous(), full[:, 3].contiguous()


def render_sharded(local_tensorfs, ray_ids, view_ids, W, H, group=None, **kw):
    """LocalTensorfs.forward with the rays of an eval batch (single view) sharded over the ranks;
    every rank returns the full (rgb, depth)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = ray_ids.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    rgb, depth, _, _ = local_tensorfs(ray_ids[lo:hi], view_ids, W, H, **kw)
    return gather_pixels(rgb, depth, n, group)
