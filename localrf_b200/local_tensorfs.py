"""Host-side mirror of the reference's scene model (local_tensorfs.py:31-499).

`LocalTensorfs` keeps the reference's constructor kwargs, attributes, `state_dict` key set
(`blending_weights, init_focal, focal_offset, center_rel, r_c2w.i, t_c2w.i, exposure.i, world2rf.k,
tensorfs.k.*`) and method signatures, so the reference's train.py / renderer.py can import it
unchanged.  `forward` is the B200 path: every active field is ONE fused kernel launch over the
whole ray batch (ray generation, march, shading, blend accumulation; exposure + clamp ride on the
last field's launch).  Differences that are deliberate:

  * fields are never parked on the CPU (180 GB of HBM holds every field of a scene); the
    reference's CPU<->GPU shuffle (local_tensorfs.py:132,432-434,476-479) has no equivalent;
  * `chunk` does not bound memory any more (the kernel materialises nothing per sample); it is
    honoured only in train mode, where each chunk draws its own jitter like the reference;
  * `sixD_to_mtx` uses the per-view cross product for every batch size (see utils.py).
"""
import ctypes as C
import math
import re

import torch

from . import _lib
from .tensorf import AlphaGridMask, TensorVMSplit, _ptr, _require_cuda, _stream
from .utils import N_to_reso, mtx_to_sixD, sixD_to_mtx


def ids2pixel_view(W, H, ids):
    """local_tensorfs.py:14-21."""
    return ids % W, (ids // W) % H, ids // (W * H)


def ids2pixel(W, H, ids):
    """local_tensorfs.py:23-29."""
    return ids % W, (ids // W) % H


class _EvalPlan:
    """Everything an eval-mode `LocalTensorfs.forward` call derives from the module state, resolved
    once: active fields, per-launch LrfBatch structs, the device copies of cameras / intrinsics /
    exposure / blending rows.  A repeated call (the renderer's per-frame or per-batch loop) then only
    re-checks the guards, allocates outputs and enqueues the launches.

    `guards` = (container, index, tensor, version): the plan is valid while every guarded slot still
    holds the same tensor object at the same version (optimiser steps, load_state_dict, append_* all
    bump one of them).  Field parameters are guarded by TensorBase.field_and_prepared's own memo."""
    __slots__ = ("guards", "launches", "keep", "n", "dev", "out_struct", "key_refs")

    def valid(self):
        for cont, idx, t, ver in self.guards:
            cur = cont[idx]
            if ver == "eq":                          # plain value (e.g. nSamples)
                if cur != t:
                    return False
            elif cur is not t or (ver is not None and t._version != ver):
                return False
        return True


def _uva_pointer(t):
    """Device-usable address of a tensor the kernel may touch directly: CUDA memory, or PINNED host
    memory (zero-copy over PCIe: under unified addressing the device address equals the host one)."""
    if t.is_cuda:
        return t.data_ptr()
    if not t.is_pinned():
        raise RuntimeError("localrf_b200: host tensors handed to the render path must be pinned "
                           "(page-locked) memory; pageable memory is not device-accessible")
    return t.data_ptr()


class LocalTensorfs(torch.nn.Module):
    """Self-calibrating sequence of local radiance fields."""

    def __init__(self, fov, n_init_frames, n_overlap, WH, n_iters_per_frame, n_iters_reg,
                 lr_R_init, lr_t_init, lr_i_init, lr_exposure_init, rf_lr_init, rf_lr_basis,
                 lr_decay_target_ratio, N_voxel_list, update_AlphaMask_list, camera_prior, device,
                 lr_upsample_reset, **tensorf_args):
        super().__init__()
        # hyper-parameters, under the attribute names the reference's callers read
        # (local_tensorfs.py:61-82); the *_per_frame_* copies feed the per-field schedule rescaling
        self.W, self.H = WH
        for name, value in dict(
                fov=fov, n_init_frames=n_init_frames, n_overlap=n_overlap,
                n_iters_per_frame=n_iters_per_frame, n_iters_reg_per_frame=n_iters_reg,
                lr_R_init=lr_R_init, lr_t_init=lr_t_init, lr_i_init=lr_i_init,
                lr_exposure_init=lr_exposure_init, rf_lr_init=rf_lr_init, rf_lr_basis=rf_lr_basis,
                lr_decay_target_ratio=lr_decay_target_ratio, N_voxel_per_frame_list=N_voxel_list,
                update_AlphaMask_per_frame_list=update_AlphaMask_list, camera_prior=camera_prior,
                tensorf_args=tensorf_args, lr_upsample_reset=lr_upsample_reset,
                device=torch.device(device)).items():
            setattr(self, name, value)
        # schedule state of the field being optimised (rescaled in optimizer_step at iteration 1)
        self.is_refining, self.regularize, self.lr_factor = False, True, 1
        self.n_iters, self.n_iters_reg = n_iters_per_frame, n_iters_reg
        self.N_voxel_list, self.update_AlphaMask_list = N_voxel_list, update_AlphaMask_list

        # per-frame pose / exposure parameters, one optimiser each (local_tensorfs.py:85-92)
        self.r_c2w = torch.nn.ParameterList()
        self.t_c2w = torch.nn.ParameterList()
        self.exposure = torch.nn.ParameterList()
        self.r_optimizers, self.t_optimizers, self.exp_optimizers = [], [], []
        self.pose_linked_rf = []
        self.blending_weights = torch.nn.Parameter(
            torch.ones([1, 1], device=self.device), requires_grad=False)
        for _ in range(n_init_frames):
            self.append_frame()

        if self.camera_prior is not None:
            focal = self.camera_prior["transforms"]["fl_x"] * self.W / self.camera_prior["transforms"]["w"]
        else:
            focal = self.W / math.tan(fov * math.pi / 180 / 2) / 2
        self.init_focal = torch.nn.Parameter(torch.Tensor([focal]).to(self.device))
        self.focal_offset = torch.nn.Parameter(torch.ones(1, device=device))
        self.center_rel = torch.nn.Parameter(0.5 * torch.ones(2, device=device))
        if lr_i_init > 0:
            self.intrinsic_optimizer = torch.optim.Adam(
                [self.focal_offset, self.center_rel], betas=(0.9, 0.99), lr=self.lr_i_init)

        # radiance fields
        self.dp_group = None      # set to a process group (or True = WORLD) for data-parallel training
        self.tensorfs = torch.nn.ParameterList()
        self.rf_iter = []
        self.world2rf = torch.nn.ParameterList()
        self.append_rf()

    # ---------------------------------------------------------------------------------------------
    # progressive-schedule bookkeeping (host side; local_tensorfs.py:116-290)
    # ---------------------------------------------------------------------------------------------
    def append_rf(self, n_added_frames=1):
        """Opens a new field centred at the latest camera and cross-fades the blending rows of the
        last `n_overlap` frames from the previous field to the new one (:116-146)."""
        self.is_refining = False
        if len(self.tensorfs) > 0:
            n_ov = min(n_added_frames, self.n_overlap, self.blending_weights.shape[0] - 1)
            ramp = 1 / n_ov + torch.arange(0, 1, 1 / n_ov)
            bw = self.blending_weights.detach().clone()
            bw[-n_ov:, -1] = (1 - ramp).to(bw)
            new_col = torch.zeros_like(bw[:, :1])
            new_col[-n_ov:, 0] = ramp.to(bw)
            self.blending_weights = torch.nn.Parameter(torch.cat([bw, new_col], dim=1),
                                                       requires_grad=False)
            world2rf = -self.t_c2w[-1].clone().detach()
            # (the reference parks the previous field on the CPU here; it stays resident)
        else:
            world2rf = torch.zeros(3, device=self.device)
        self.tensorfs.append(TensorVMSplit(device=self.device, **self.tensorf_args))
        if self.__dict__.get("grid_storage", "fp32") != "fp32":
            self.tensorfs[-1].set_grid_storage(self.grid_storage)
        self.world2rf.append(world2rf.clone().detach())
        self.rf_iter.append(0)
        groups = self.tensorfs[-1].get_optparam_groups(self.rf_lr_init, self.rf_lr_basis)
        self.rf_optimizer = torch.optim.Adam(groups, betas=(0.9, 0.99))

    def set_grid_storage(self, kind="fp32"):
        """Inference storage type of every field's grids (TensorBase.set_grid_storage): "fp32" or "bf16"
        (eval renders gather bfloat16 copies: half the bytes per texel; training is unaffected).  Fields
        appended later inherit the setting."""
        for rf in self.tensorfs:
            rf.set_grid_storage(kind)
        self.__dict__["grid_storage"] = kind
        return self

    def append_frame(self):
        """Adds one frame's pose / exposure parameters, initialised from the previous frame
        (:148-177)."""
        if len(self.r_c2w) == 0:
            self.r_c2w.append(torch.eye(3, 2, device=self.device))
            self.t_c2w.append(torch.zeros(3, device=self.device))
            self.pose_linked_rf.append(0)
        else:
            prev_r = self.r_c2w[-1].clone().detach()[None]
            self.r_c2w.append(mtx_to_sixD(sixD_to_mtx(prev_r))[0])
            self.t_c2w.append(self.t_c2w[-1].clone().detach())
            self.blending_weights = torch.nn.Parameter(
                torch.cat([self.blending_weights, self.blending_weights[-1:, :]], dim=0),
                requires_grad=False)
            self.pose_linked_rf.append(int(torch.nonzero(self.blending_weights[-1, :])[0]))
        self.exposure.append(torch.eye(3, 3, device=self.device))

        if self.camera_prior is not None:
            idx = len(self.r_c2w) - 1
            rel_pose = self.camera_prior["rel_poses"][idx]
            last_r = sixD_to_mtx(self.r_c2w[-1].clone().detach()[None])[0]
            self.r_c2w[-1] = last_r @ rel_pose[:3, :3]
            self.t_c2w[-1].data += last_r @ rel_pose[:3, 3]

        adam = lambda p, lr: torch.optim.Adam([p], betas=(0.9, 0.99), lr=lr)
        self.r_optimizers.append(adam(self.r_c2w[-1], self.lr_R_init))
        self.t_optimizers.append(adam(self.t_c2w[-1], self.lr_t_init))
        self.exp_optimizers.append(adam(self.exposure[-1], self.lr_exposure_init))

    def _live_pose_ids(self):
        """Frames whose pose is still optimised: linked to the newest field, within its budget."""
        newest = len(self.rf_iter) - 1
        if self.rf_iter[-1] >= self.n_iters:
            return []
        return [i for i, rf in enumerate(self.pose_linked_rf) if rf == newest]

    def optimizer_step_poses_only(self, loss):
        live = self._live_pose_ids()
        for i in live:
            self.r_optimizers[i].zero_grad()
            self.t_optimizers[i].zero_grad()
        loss.backward()
        for i in live:
            self.r_optimizers[i].step()
            self.t_optimizers[i].step()

    def optimizer_step(self, loss, optimize_poses):
        """One optimisation step + the schedule events it triggers (:193-290)."""
        it = self.rf_iter[-1]
        if it == 0:
            self.lr_factor = 1
            self.n_iters = self.n_iters_per_frame
            self.n_iters_reg = self.n_iters_reg_per_frame
        elif it == 1:
            n_train = (self.blending_weights[:, -1] > 0).sum()
            self.n_iters = int(self.n_iters_per_frame * n_train)
            self.n_iters_reg = int(self.n_iters_reg_per_frame * n_train)
            self.lr_factor = self.lr_decay_target_ratio ** (1 / self.n_iters)
            self.N_voxel_list = {int(k * n_train): v for k, v in self.N_voxel_per_frame_list.items()}
            self.update_AlphaMask_list = [int(u * n_train) for u in self.update_AlphaMask_per_frame_list]
        self.regularize = self.rf_iter[-1] < self.n_iters_reg

        def decay(opt):
            for group in opt.param_groups:
                group["lr"] *= self.lr_factor

        live = self._live_pose_ids()
        for i in live:
            if optimize_poses:
                decay(self.r_optimizers[i]); decay(self.t_optimizers[i])
                self.r_optimizers[i].zero_grad(); self.t_optimizers[i].zero_grad()
            if self.lr_exposure_init > 0:
                decay(self.exp_optimizers[i])
                self.exp_optimizers[i].zero_grad()
        tune_intrinsics = (self.lr_i_init > 0 and self.blending_weights.shape[1] == 1
                           and self.is_refining)
        if tune_intrinsics:
            decay(self.intrinsic_optimizer)
            self.intrinsic_optimizer.zero_grad()
        self.rf_optimizer.zero_grad()

        loss.backward()
        if getattr(self, "dp_group", None) is not None:     # data-parallel training: average the step's
            from .dist import allreduce_gradients            # gradients over the ranks before any update
            allreduce_gradients(self, None if self.dp_group is True else self.dp_group)

        self.rf_optimizer.step()
        if self.is_refining:
            decay(self.rf_optimizer)

        if self.rf_iter[-1] in self.N_voxel_list:          # raise the grid resolution
            reso = N_to_reso(self.N_voxel_list[self.rf_iter[-1]], self.tensorfs[-1].aabb)
            self.tensorfs[-1].upsample_volume_grid(reso)
            if self.lr_upsample_reset:
                print("reset lr to initial")
                groups = self.tensorfs[-1].get_optparam_groups(self.rf_lr_init, self.rf_lr_basis)
                self.rf_optimizer = torch.optim.Adam(groups, betas=(0.9, 0.99))
        if self.rf_iter[-1] in self.update_AlphaMask_list:  # rebuild the occupancy mask
            self.tensorfs[-1].updateAlphaMask(tuple((self.tensorfs[-1].gridSize / 2).int()))

        for i in live:
            if optimize_poses:
                self.r_optimizers[i].step(); self.t_optimizers[i].step()
            if self.lr_exposure_init > 0:
                self.exp_optimizers[i].step()
        if tune_intrinsics:
            self.intrinsic_optimizer.step()
        if self.is_refining:
            self.rf_iter[-1] += 1
        return self.rf_iter[-1] >= self.n_iters - 1       # can_add_rf

    # ---------------------------------------------------------------------------------------------
    def get_cam2world(self, view_ids=None, starting_id=0):
        """[V,3,4] camera-to-world matrices from the 6-D rotations and translations (:292-299)."""
        if view_ids is not None:
            ids = view_ids.tolist() if torch.is_tensor(view_ids) else list(view_ids)
            r = torch.stack([self.r_c2w[i] for i in ids], dim=0)
            t = torch.stack([self.t_c2w[i] for i in ids], dim=0)
        else:
            r = torch.stack(list(self.r_c2w[starting_id:]), dim=0)
            t = torch.stack(list(self.t_c2w[starting_id:]), dim=0)
        return torch.cat([sixD_to_mtx(r), t[..., None]], dim=-1)

    _CKPT_FIELDS = ("fov", "n_init_frames", "n_overlap", "n_iters_per_frame", "lr_R_init", "lr_t_init",
                    "lr_i_init", "lr_exposure_init", "rf_lr_init", "rf_lr_basis",
                    "lr_decay_target_ratio", "lr_upsample_reset")

    def get_kwargs(self):
        """Constructor kwargs stored next to the state_dict in a checkpoint (:301-324)."""
        kwargs = {name: getattr(self, name) for name in self._CKPT_FIELDS}
        kwargs.update(camera_prior=None, WH=(self.W, self.H), n_iters_reg=self.n_iters_reg_per_frame,
                      N_voxel_list=self.N_voxel_per_frame_list,
                      update_AlphaMask_list=self.update_AlphaMask_per_frame_list)
        kwargs.update(self.tensorfs[0].get_kwargs())
        return kwargs

    def save(self, path):
        torch.save({"kwargs": self.get_kwargs(), "state_dict": self.state_dict()}, path)

    def load(self, state_dict):
        """Rebuilds the frame / field structure from the key names, then loads (:331-356)."""
        n_frames = 0
        for key in state_dict:
            if re.fullmatch(r"r_c2w.[0-9]*", key):
                n_frames += 1
            if re.fullmatch(r"tensorfs.[1-9][0-9]*.density_plane.0", key):
                plane0 = state_dict[key]
                line0 = state_dict[key[: -len("density_plane.0")] + "density_line.0"]
                # plane 0 is [1,C,G_y,G_x], line 0 is [1,C,G_z,1].  (The reference reads x and y
                # swapped here, which only works for G_x == G_y; the intended order is used.)
                self.tensorf_args["gridSize"] = [plane0.shape[3], plane0.shape[2], line0.shape[2]]
                self.append_rf()
        for i, rf in enumerate(self.tensorfs):
            if f"tensorfs.{i}.alphaMask.aabb" in state_dict:
                vol = state_dict[f"tensorfs.{i}.alphaMask.alpha_volume"].to(self.device)
                aabb = state_dict[f"tensorfs.{i}.alphaMask.aabb"].to(self.device)
                rf.alphaMask = AlphaGridMask(self.device, aabb, vol)
        for _ in range(n_frames - len(self.r_c2w)):
            self.append_frame()
        self.blending_weights = torch.nn.Parameter(
            torch.ones_like(state_dict["blending_weights"]), requires_grad=False)
        self.load_state_dict(state_dict)

    def get_dist_to_last_rf(self):
        return torch.norm(self.t_c2w[-1] + self.world2rf[-1])

    def get_reg_loss(self, tvreg, TV_weight_density, TV_weight_app, L1_weight_inital):
        tv_loss, l1_loss = 0, 0
        if self.rf_iter[-1] < self.n_iters:
            anneal = self.lr_factor ** self.rf_iter[-1]
            if TV_weight_density > 0:
                tv_loss += self.tensorfs[-1].TV_loss_density(tvreg).mean() * (TV_weight_density * anneal)
            if TV_weight_app > 0:
                tv_loss += self.tensorfs[-1].TV_loss_app(tvreg).mean() * (TV_weight_app * anneal)
            if L1_weight_inital > 0:
                l1_loss += self.tensorfs[-1].density_L1() * L1_weight_inital
        return tv_loss, l1_loss

    def focal(self, W):
        return self.init_focal * self.focal_offset * W / self.W

    def center(self, W, H):
        return torch.Tensor([W, H]).to(self.center_rel) * self.center_rel

    # ---------------------------------------------------------------------------------------------
    # the hot path (local_tensorfs.py:382-499)
    # ---------------------------------------------------------------------------------------------
    def _exposure_for(self, ids, test_id, dev):
        """Per-view 3x3 exposure [V,3,3]; held-out frames average their neighbours with the
        reference's edge rules (:481-493).  `ids` is the host list of view ids."""
        if not test_id:
            mats = [self.exposure[i] for i in ids]
        else:
            n = len(self.exposure)
            mats = []
            for v in ids:
                lo = max(v - 1, 0)
                if lo == v:
                    lo = 1
                hi = min(v + 1, n - 1)
                if lo == v:                     # the reference re-tests the rewritten lower index
                    hi = n - 2
                mats.append((self.exposure[lo].detach() + self.exposure[hi].detach()) / 2)
        out = mats[0].detach()[None] if len(mats) == 1 else torch.stack([m.detach() for m in mats])
        return out.to(dev, torch.float32).contiguous()

    def _forward_autograd(self, ray_ids, view_ids, ids, W, H, white_bg, is_train, cam2world,
                          world2rf, blend, active, chunk, test_id, floater_thresh):
        """Composed path (local_tensorfs.py:397-499 with torch ops around the CUDA lookups): rays are
        generated with torch ops so gradients reach poses and intrinsics, each active field renders
        through TensorBase.forward (composed when autograd records or the field uses positional
        encodings), blend / exposure / clamp are torch ops.  Chunked like the reference, because
        this path materialises per-sample tensors."""
        from .ray_utils import get_ray_directions_360, get_ray_directions_lean, get_rays_lean
        dev = ray_ids.device
        n, n_views = ray_ids.shape[0], len(ids)
        i, j = ids2pixel(W, H, ray_ids)
        ij = torch.stack([i, j], dim=-1)
        if self.fov == 360:
            directions = get_ray_directions_360(i, j, W, H)
        else:
            directions = get_ray_directions_lean(i, j, self.focal(W), self.center(W, H))
        if cam2world is None:
            cam2world = self.get_cam2world(ids)
        per_view = n // n_views
        rgbs = torch.zeros_like(directions)
        depth_maps = torch.zeros_like(directions[..., 0])
        budget = max(chunk // len(active), 1)
        step = n if n <= budget else max(per_view, (budget // per_view) * per_view)
        outs_rgb, outs_depth = [], []
        for lo in range(0, n, step):
            hi = min(lo + step, n)
            rgb_c = torch.zeros(hi - lo, 3, device=dev)
            depth_c = torch.zeros(hi - lo, device=dev)
            for k in active:
                cam2rf = cam2world.clone()
                cam2rf[:, :3, 3] = cam2rf[:, :3, 3] + world2rf[k]
                cam2rf = cam2rf.repeat_interleave(per_view, dim=0)[lo:hi]
                rays_o, rays_d = get_rays_lean(directions[lo:hi], cam2rf)
                rgb_k, depth_k = self.tensorfs[k](torch.cat([rays_o, rays_d], -1), is_train=is_train,
                                                  white_bg=white_bg, N_samples=-1,
                                                  refine=self.is_refining,
                                                  floater_thresh=floater_thresh)
                if blend is None:
                    w = 1.0
                    rgb_c, depth_c = rgb_c + rgb_k, depth_c + depth_k
                else:
                    w = blend.repeat_interleave(per_view, dim=0)[lo:hi, k]
                    rgb_c, depth_c = rgb_c + rgb_k * w[..., None], depth_c + depth_k * w
            outs_rgb.append(rgb_c); outs_depth.append(depth_c)
        rgbs = torch.cat(outs_rgb); depth_maps = torch.cat(outs_depth)
        if self.lr_exposure_init > 0:
            if test_id:
                exposure = self._exposure_for(ids, True, dev)
            else:
                exposure = torch.stack([self.exposure[v] for v in ids], dim=0)
            exposure = exposure.repeat_interleave(per_view, dim=0)
            rgbs = torch.bmm(exposure, rgbs[..., None])[..., 0]
        return rgbs.clamp(0, 1), depth_maps, directions, ij

    def _cached(self, slot, key, build, refs=()):
        """Small memo for tensors derived from parameters: rebuilt when the key changes.  Keys are made
        of ids + version counters of the source tensors; `refs` keeps those tensors alive while the
        entry exists, so a freed tensor's id can never be reused by a different one."""
        c = self.__dict__.setdefault("_memo", {})
        hit = c.get(slot)
        if hit is None or hit[0] != key:
            hit = (key, build(), tuple(refs))
            c[slot] = hit
        return hit[1]

    def _blend_host(self):
        bw = self.blending_weights
        return self._cached("bw_host", (id(bw), bw._version), lambda: bw.detach().cpu(), refs=[bw])

    def forward(self, ray_ids, view_ids, W, H, white_bg=True, is_train=True, cam2world=None,
                world2rf=None, blending_weights=None, chunk=16384, test_id=False,
                floater_thresh=0, stats=None, exchange=None, out=None):
        """-> (rgbs [N,3], depth_maps [N], directions [N,3], ij [N,2]) like the reference.

        Extensions (all optional, defaults = the reference's behaviour):
        `ray_ids` may be a PINNED host tensor: the kernel reads the ids over PCIe itself (no staging
        copy).  `out=(rgb [N,3], depth [N])` preallocated float32 outputs, on the device or in pinned
        host memory -- in the latter case finished pixels are stored straight into host memory by the
        kernel and only a stream synchronisation is left to the caller.
        `exchange=(PixelExchange, ray_lo)` (multi-GPU eval, dist.render_sharded): the kernel also
        stores every finished pixel into all peers' gathered buffers at row ray_lo + r and the step
        barrier is enqueued after the last launch; rgbs / depth_maps are then views of an interleaved
        [N,4] buffer."""
        if ray_ids.is_cuda:
            dev = ray_ids.device
        else:
            _uva_pointer(ray_ids)                       # raises unless pinned
            dev = self.blending_weights.device
            if dev.type != "cuda":
                raise RuntimeError("localrf_b200: the scene model is on the CPU; the render path runs "
                                   "only on a CUDA device (there is no CPU fallback)")
        n = ray_ids.shape[0]
        if n == 0:                                        # an empty batch / shard: nothing to launch
            if exchange is not None:
                exchange[0].skip_step(_stream(dev))       # ... but the peers still need this rank's step flag
            return (torch.empty(0, 3, device=dev), torch.empty(0, device=dev), torch.empty(0, 3, device=dev),
                    torch.empty(0, 2, dtype=torch.int64, device=dev))
        if not is_train and not torch.is_grad_enabled():
            # eval fast path: reuse the resolved plan of an identical earlier call
            vkey = (id(view_ids), view_ids._version) if torch.is_tensor(view_ids) else tuple(view_ids)
            key = (vkey, n, W, H, bool(white_bg), bool(test_id), float(floater_thresh),
                   None if cam2world is None else (id(cam2world), cam2world._version),
                   None if world2rf is None else id(world2rf),
                   None if blending_weights is None else (id(blending_weights), blending_weights._version),
                   id(self.blending_weights), len(self.tensorfs), exchange is not None, bool(self.is_refining), dev)
            plans = self.__dict__.setdefault("_plans", {})
            plan = plans.get(key)
            if plan is not None and plan.valid():
                return self._run_plan(plan, ray_ids, stats, exchange, out)
            res = self._forward_general(ray_ids, view_ids, W, H, white_bg, is_train, cam2world, world2rf,
                                        blending_weights, chunk, test_id, floater_thresh, stats, exchange,
                                        out, dev, want_plan=True)
            if isinstance(res, tuple) and len(res) == 2 and isinstance(res[0], _EvalPlan):
                plan, res = res
                # the key holds ids of caller objects: keep them alive so the ids cannot be recycled
                plan.key_refs = (view_ids, cam2world, world2rf, blending_weights)
                if len(plans) >= 256:
                    plans.clear()
                plans[key] = plan
            return res
        return self._forward_general(ray_ids, view_ids, W, H, white_bg, is_train, cam2world, world2rf,
                                     blending_weights, chunk, test_id, floater_thresh, stats, exchange,
                                     out, dev, want_plan=False)

    def _alloc_outputs(self, n, dev, exchange, out):
        """(pix or None, rgbs, depth, directions, ij) -- fresh tensors owned by the caller."""
        pix = None
        if exchange is not None:
            pix = torch.empty(n, 4, dtype=torch.float32, device=dev)
            rgbs, depth = pix[:, :3], pix[:, 3]
        elif out is not None:
            rgbs, depth = out
            for t, shape in ((rgbs, (n, 3)), (depth, (n,))):
                if t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous():
                    raise ValueError("out=(rgb [N,3], depth [N]) must be contiguous float32 tensors")
                _uva_pointer(t)
        else:
            rgbs = torch.empty(n, 3, dtype=torch.float32, device=dev)
            depth = torch.empty(n, dtype=torch.float32, device=dev)
        directions = torch.empty(n, 3, dtype=torch.float32, device=dev)
        ij = torch.empty(n, 2, dtype=torch.int64, device=dev)
        return pix, rgbs, depth, directions, ij

    def _run_plan(self, plan, ray_ids, stats, exchange, out):
        dev, n = plan.dev, plan.n
        rays_i = ray_ids
        if rays_i.dtype != torch.int64 or not rays_i.is_contiguous():
            rays_i = rays_i.to(torch.int64).contiguous()
        pix, rgbs, depth, directions, ij = self._alloc_outputs(n, dev, exchange, out)
        lib = _lib.lib()
        o = plan.out_struct
        o.pix = o.rgb = o.depth = None
        o.n_peers, o.signal_seq, o.wait_seq, o.mc_pix = 0, 0, 0, None   # (exchange fields are set for the last launch only)
        if pix is not None:
            o.pix = pix.data_ptr()
        else:
            o.rgb, o.depth = rgbs.data_ptr(), depth.data_ptr()
        o.directions, o.ij = directions.data_ptr(), ij.data_ptr()
        o.stats = stats.data_ptr() if stats is not None else None
        ids_ptr = rays_i.data_ptr()
        with torch.cuda.device(dev):
            stream = _stream(dev)
            last = len(plan.launches) - 1
            for pos, (rf, z, b) in enumerate(plan.launches):
                fs, prep = rf.field_and_prepared(z)
                b.ray_ids = ids_ptr
                if pos == last and exchange is not None:
                    exchange[0].fill_outputs(o, exchange[1])
                _lib.check(lib.lrf_render(C.byref(fs), _ptr(prep), C.byref(b), C.byref(o), stream))
            if exchange is not None:
                exchange[0].close_step(stream)
        return rgbs, depth, directions, ij

    def _forward_general(self, ray_ids, view_ids, W, H, white_bg, is_train, cam2world, world2rf,
                         blending_weights, chunk, test_id, floater_thresh, stats, exchange, out, dev,
                         want_plan):
        n = ray_ids.shape[0]
        # (not memoised: a new tensor can reuse a freed tensor's id and storage)
        ids = view_ids.tolist() if torch.is_tensor(view_ids) else [int(v) for v in view_ids]
        n_views = len(ids)
        if n_views == 0 or n % n_views != 0:
            raise ValueError("ray_ids must hold the same number of rays for every view")
        n_fields = len(self.tensorfs)
        guards = []

        # -- which fields render, with which per-view weights (:403-418) ----------------------------
        blend = None                                   # device [V, n_fields] or None (= weight 1)
        if is_train:                                   # one field trains at a time (:411-416)
            active = [n_fields - 1]
        elif blending_weights is None:
            rows = self._blend_host()[ids]
            active = torch.nonzero(rows.sum(dim=0))[:, 0].tolist()
            bw = self.blending_weights
            blend = bw.detach()[ids[0]:ids[0] + 1] if n_views == 1 else bw.detach()[view_ids]
            blend = blend.to(dev, torch.float32).contiguous()
        else:
            blend = blending_weights.detach().to(dev, torch.float32).contiguous()
            active = torch.nonzero(blend.sum(dim=0))[:, 0].tolist()
        if len(active) == 0:
            print("****** No valid RF")
            i, j = ids2pixel(W, H, ray_ids)
            ones = torch.ones_like(ray_ids).float()
            return torch.ones([n, 3], device=dev), ones, torch.zeros(n, 3, device=dev), torch.stack([i, j], -1)

        if world2rf is None:
            world2rf = self.world2rf
        if torch.is_grad_enabled():
            tracked = [self.init_focal, self.focal_offset, self.center_rel]
            tracked += [cam2world] if cam2world is not None else \
                [self.r_c2w[i] for i in ids] + [self.t_c2w[i] for i in ids]
            tracked += [self.exposure[i] for i in ids] if self.lr_exposure_init > 0 else []
            for k in active:
                tracked += list(self.tensorfs[k].parameters())
            needs_composed = any(t is not None and t.requires_grad for t in tracked)
        else:
            needs_composed = False
        if needs_composed or not all(self.tensorfs[k].fused_supported() for k in active):
            if exchange is not None or out is not None:
                raise NotImplementedError("exchange= / out= are served by the fused eval path only")
            if not ray_ids.is_cuda:
                ray_ids = ray_ids.to(dev)
            return self._forward_autograd(ray_ids, view_ids, ids, W, H, white_bg, is_train,
                                          cam2world, world2rf, blend, active, chunk, test_id,
                                          floater_thresh)
        if exchange is not None and is_train:
            raise ValueError("the fused pixel exchange serves eval batches (is_train=False)")

        # -- cameras, intrinsics, exposure (memoised on parameter versions) ---------------------------
        P = self._parameters
        if blending_weights is None and not is_train:
            guards.append((P, "blending_weights", self.blending_weights, self.blending_weights._version))
        if cam2world is None:
            src = [self.r_c2w[i] for i in ids] + [self.t_c2w[i] for i in ids]
            key = (tuple(ids), tuple((id(t), t._version) for t in src), str(dev))
            cam2world = self._cached("c2w", key, lambda: self.get_cam2world(ids).detach()
                                     .to(dev, torch.float32).contiguous(), refs=src)
            for i in set(ids):     # (guards index the lists' _parameters dicts: ~10x cheaper than __getitem__)
                guards.append((self.r_c2w._parameters, str(i), self.r_c2w[i], self.r_c2w[i]._version))
                guards.append((self.t_c2w._parameters, str(i), self.t_c2w[i], self.t_c2w[i]._version))
        else:
            cam2world = cam2world.detach().to(dev, torch.float32).contiguous()
        fov360 = self.fov == 360
        src = [self.init_focal, self.focal_offset, self.center_rel]
        key = (W, H, self.W, tuple((id(t), t._version) for t in src), str(dev))
        intr = self._cached("intr", key, lambda: torch.cat(
            [self.focal(W).detach().reshape(1), self.center(W, H).detach().reshape(2)])
            .to(dev, torch.float32).contiguous(), refs=src)
        for name in ("init_focal", "focal_offset", "center_rel"):
            guards.append((P, name, P[name], P[name]._version))
        exposure = None
        if self.lr_exposure_init > 0:
            n_e = len(self.exposure)
            used = sorted(set(ids + ([max(v - 1, 0) for v in ids] + [min(v + 1, n_e - 1) for v in ids]
                                     + [min(1, n_e - 1), max(n_e - 2, 0)] if test_id else [])))
            src = [self.exposure[i] for i in used]
            key = (tuple(ids), bool(test_id), n_e, tuple((id(t), t._version) for t in src), str(dev))
            exposure = self._cached("expo", key, lambda: self._exposure_for(ids, test_id, dev), refs=src)
            for i, t in zip(used, src):
                guards.append((self.exposure._parameters, str(i), t, t._version))
        rays_i = ray_ids.detach()
        if rays_i.dtype != torch.int64 or not rays_i.is_contiguous():
            rays_i = rays_i.to(torch.int64).contiguous()
        ids_ptr = _uva_pointer(rays_i)

        pix, rgbs, depth, directions, ij = self._alloc_outputs(n, dev, exchange, out)

        # train mode: the reference draws fresh jitter per chunk; eval has no randomness, so the
        # whole batch is one launch per field.  Chunks are kept on view boundaries (the kernel
        # derives a ray's view from its position in the launch).
        per_view = n // n_views
        budget = max(chunk // len(active), 1)
        if is_train and n > budget:
            chunk = max(per_view, (budget // per_view) * per_view)
        else:
            chunk = n
        lib = _lib.lib()
        launches, keep = [], [cam2world, intr, exposure, blend]
        with torch.cuda.device(dev):
            stream = _stream(dev)
            for lo in range(0, n, chunk):
                hi = min(lo + chunk, n)
                v_lo, v_hi = lo // per_view, (hi + per_view - 1) // per_view
                for pos, k in enumerate(active):
                    rf = self.tensorfs[k]
                    if rf.basis_mat.weight.device != dev:
                        rf.to(dev)
                    z = rf.sample_table(is_train, -1, dev)
                    fs, prep = rf.field_and_prepared(z)
                    w2 = world2rf[k]
                    w2rf = w2.detach() if (w2.device == dev and w2.dtype == torch.float32
                                            and w2.is_contiguous()) else \
                        w2.detach().to(dev, torch.float32).contiguous()
                    b = _lib.LrfBatch()
                    b.n_rays = hi - lo
                    b.ray_ids = ids_ptr + 8 * lo
                    b.W, b.H = int(W), int(H)
                    b.fov360 = int(fov360)
                    b.intrinsics = intr.data_ptr()
                    b.cam2world = cam2world.data_ptr() + 48 * v_lo
                    b.n_views = v_hi - v_lo
                    b.world2rf = w2rf.data_ptr()
                    if blend is not None:
                        b.blend = blend.data_ptr() + 4 * (v_lo * blend.shape[1] + k)
                        b.blend_stride = blend.shape[1]
                    b.accumulate = int(pos > 0)
                    b.finalize = int(pos == len(active) - 1)
                    if exposure is not None:
                        b.exposure = exposure.data_ptr() + 36 * v_lo
                    b.white_bg = int(bool(white_bg) or bool(is_train and torch.rand((1,)) < 0.5))
                    b.floater_thresh = float(floater_thresh)
                    b.refine = int(bool(self.is_refining))        # local_tensorfs.py:463
                    o = _lib.LrfOutputs()
                    if pix is not None:
                        o.pix = pix.data_ptr() + 16 * lo
                        if pos == len(active) - 1:
                            exchange[0].fill_outputs(o, exchange[1] + lo)
                    else:
                        o.rgb = _uva_pointer(rgbs) + 12 * lo
                        o.depth = _uva_pointer(depth) + 4 * lo
                    o.directions = directions.data_ptr() + 12 * lo
                    o.ij = ij.data_ptr() + 16 * lo
                    if stats is not None:
                        o.stats = stats.data_ptr()
                    _lib.check(lib.lrf_render(C.byref(fs), _ptr(prep), C.byref(b), C.byref(o), stream))
                    if want_plan:
                        launches.append((rf, z, b))
                        keep.append(w2rf)
                        guards.append((self.tensorfs._modules, str(k), rf, None))
                        if isinstance(world2rf, torch.nn.ParameterList):
                            guards.append((world2rf._parameters, str(k), w2, w2._version))
                        else:
                            guards.append((world2rf, k, w2, w2._version))
                        guards.append((rf.__dict__, "nSamples", rf.nSamples, "eq"))
            if exchange is not None:
                exchange[0].close_step(stream)
        result = (rgbs, depth, directions, ij)
        if want_plan and chunk == n:
            plan = _EvalPlan()
            plan.guards, plan.launches, plan.keep = guards, launches, keep
            plan.n, plan.dev, plan.out_struct, plan.key_refs = n, dev, _lib.LrfOutputs(), None
            return plan, result
        return result
