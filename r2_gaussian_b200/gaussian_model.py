"""GaussianModel: the parameter container that feeds render() / query().

Public surface of `r2_gaussian/gaussian/gaussian_model.py:36-556` (same attribute and method names, same
activation functions, optimizer groups, schedules, checkpoint tuple and pickle layout, same densify / clone /
split / prune rules) so the reference's `train.py` / `test.py` can drive it unchanged.  Differences are internal:

* `distCUDA2` comes from this repository's grid-hash kernel (`r2_gaussian_b200.simple_knn`);
* the optimizer is `FusedAdam` (one launch per step for all four groups, same state layout as torch's Adam);
* densification rebuilds every per-Gaussian tensor (4 parameters, 8 Adam moments, `max_radii2D`, statistics) ONCE
  and on the device: clone, split and all prune rules are evaluated into one row selection (`compact.select_rows`:
  mask -> scan -> stable index list, the count stays on the device) and ONE launch gathers all tensors through it
  (`compact.gather_rows`).  No boolean-mask indexing anywhere (each `t[mask]` of the reference is a nonzero() with a
  host synchronisation); a densification step costs two small host reads (how many clones / splits, how many
  survivors) instead of ~30.  Row order of the result equals the reference's ([survivors | clones | split
  children], then pruned), parameters / moments / RNG stream bit-equal.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
from torch import nn

from . import compact
from .gaussian_utils import (build_rotation, build_scaling_rotation, get_expon_lr_func, inverse_sigmoid,
                             inverse_softplus, strip_symmetric)
from .optim import FusedAdam
from .simple_knn import distCUDA2

EPS = 1e-5
# optimizer group name -> attribute
_GROUPS = (("xyz", "_xyz"), ("density", "_density"), ("scaling", "_scaling"), ("rotation", "_rotation"))


def _to_numpy(t):
    return t.detach().cpu().numpy()


class GaussianModel:
    # ------------------------------------------------------------------ activations
    def setup_functions(self):
        if self.scale_bound is not None:
            lo, hi = self.scale_bound
            assert lo < hi, "scale_min must be smaller than scale_max."
            self.scaling_activation = lambda x: torch.sigmoid(x) * (hi - lo) + lo
            self.scaling_inverse_activation = lambda x: inverse_sigmoid(torch.relu((x - lo) / (hi - lo)))
        else:
            self.scaling_activation = torch.exp
            self.scaling_inverse_activation = torch.log

        def covariance(scaling, scaling_modifier, rotation):
            L = build_scaling_rotation(scaling_modifier * scaling, rotation)
            return strip_symmetric(L @ L.transpose(1, 2))

        self.covariance_activation = covariance
        self.density_activation = torch.nn.Softplus()
        self.density_inverse_activation = inverse_softplus
        self.rotation_activation = torch.nn.functional.normalize

    def __init__(self, scale_bound=None):
        empty = torch.empty(0)
        self._xyz = self._scaling = self._rotation = self._density = empty
        self.max_radii2D = self.xyz_gradient_accum = self.denom = empty
        self.optimizer = None
        self.spatial_lr_scale = 0
        self.scale_bound = scale_bound
        self.setup_functions()

    # ------------------------------------------------------------------ accessors
    @property
    def get_scaling(self):
        return self.scaling_activation(self._scaling)

    @property
    def get_rotation(self):
        return self.rotation_activation(self._rotation)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_density(self):
        return self.density_activation(self._density)

    def raw_parameters(self):
        """Raw (pre-activation) parameters + what the kernels need to apply the activations themselves
        (`fused.py`: softplus, bounded sigmoid or exp, normalize folded into preprocess / per-Gaussian backward)."""
        return {"density": self._density, "scaling": self._scaling, "rotation": self._rotation,
                "scale_bound": None if self.scale_bound is None else (float(self.scale_bound[0]), float(self.scale_bound[1]))}

    def get_covariance(self, scaling_modifier=1):
        return self.covariance_activation(self.get_scaling, scaling_modifier, self._rotation)

    # ------------------------------------------------------------------ checkpoint tuple
    def capture(self):
        return (self._xyz, self._scaling, self._rotation, self._density, self.max_radii2D, self.xyz_gradient_accum,
                self.denom, self.optimizer.state_dict(), self.spatial_lr_scale, self.scale_bound)

    def restore(self, model_args, training_args):
        (self._xyz, self._scaling, self._rotation, self._density, self.max_radii2D, grad_accum, denom, opt_state,
         self.spatial_lr_scale, self.scale_bound) = model_args
        self.training_setup(training_args)
        self.xyz_gradient_accum, self.denom = grad_accum, denom
        self.optimizer.load_state_dict(opt_state)
        self.setup_functions()

    # ------------------------------------------------------------------ initialisation
    def create_from_pcd(self, xyz, density, spatial_lr_scale: float, dist2=None):
        """`dist2` (optional, not in the reference): precomputed mean squared 3-NN distances for these points --
        a Gaussian-sharded run computes them on the FULL cloud before splitting it, so that the initial scales do
        not depend on the number of ranks."""
        self.spatial_lr_scale = spatial_lr_scale
        points = torch.as_tensor(np.asarray(xyz)).float().cuda()
        print("Initialize gaussians from {} estimated points".format(points.shape[0]))
        raw_density = self.density_inverse_activation(torch.as_tensor(np.asarray(density))).float().cuda()
        nn2 = distCUDA2(points) if dist2 is None else torch.as_tensor(np.asarray(dist2)).float().cuda().reshape(-1)
        dist = torch.sqrt(torch.clamp_min(nn2, 0.001 ** 2))
        if self.scale_bound is not None:
            dist = torch.clamp(dist, self.scale_bound[0] + EPS, self.scale_bound[1] - EPS)   # keep the inverse finite
        raw_scale = self.scaling_inverse_activation(dist)[..., None].repeat(1, 3)
        quat = torch.zeros((points.shape[0], 4), device="cuda")
        quat[:, 0] = 1
        self._xyz = nn.Parameter(points.requires_grad_(True))
        self._scaling = nn.Parameter(raw_scale.requires_grad_(True))
        self._rotation = nn.Parameter(quat.requires_grad_(True))
        self._density = nn.Parameter(raw_density.requires_grad_(True))
        self.max_radii2D = torch.zeros((points.shape[0]), device="cuda")

    # ------------------------------------------------------------------ optimizer + schedules
    def training_setup(self, training_args):
        n = self.get_xyz.shape[0]
        self.xyz_gradient_accum = torch.zeros((n, 1), device="cuda")
        self.denom = torch.zeros((n, 1), device="cuda")
        scale = self.spatial_lr_scale
        prefix = {"xyz": "position", "density": "density", "scaling": "scaling", "rotation": "rotation"}
        groups, self._schedules = [], {}
        for name, attr in _GROUPS:
            init = getattr(training_args, prefix[name] + "_lr_init") * scale
            final = getattr(training_args, prefix[name] + "_lr_final") * scale
            steps = getattr(training_args, prefix[name] + "_lr_max_steps")
            groups.append({"params": [getattr(self, attr)], "lr": init, "name": name})
            self._schedules[name] = get_expon_lr_func(lr_init=init, lr_final=final, max_steps=steps)
        self.optimizer = FusedAdam(groups, lr=0.0, eps=1e-15)
        self.xyz_scheduler_args = self._schedules["xyz"]
        self.density_scheduler_args = self._schedules["density"]
        self.scaling_scheduler_args = self._schedules["scaling"]
        self.rotation_scheduler_args = self._schedules["rotation"]

    def update_learning_rate(self, iteration):
        for group in self.optimizer.param_groups:
            sched = self._schedules.get(group["name"])
            if sched is not None:
                group["lr"] = sched(iteration)

    # ------------------------------------------------------------------ export / import
    def construct_list_of_attributes(self):
        names = ["x", "y", "z", "nx", "ny", "nz", "density"]
        names += ["scale_{}".format(i) for i in range(self._scaling.shape[1])]
        names += ["rot_{}".format(i) for i in range(self._rotation.shape[1])]
        return names

    def save_ply(self, path):
        """Pickle (the reference keeps the `.ply` name in its API but writes a pickle, `:263-281`)."""
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        blob = {"xyz": _to_numpy(self._xyz), "density": _to_numpy(self._density), "scale": _to_numpy(self._scaling),
                "rotation": _to_numpy(self._rotation), "scale_bound": self.scale_bound}
        with open(path, "wb") as f:
            pickle.dump(blob, f, pickle.HIGHEST_PROTOCOL)

    def load_ply(self, path):
        with open(path, "rb") as f:
            blob = pickle.load(f)
        for key, attr in (("xyz", "_xyz"), ("density", "_density"), ("scale", "_scaling"), ("rotation", "_rotation")):
            t = torch.tensor(blob[key], dtype=torch.float, device="cuda")
            setattr(self, attr, nn.Parameter(t.requires_grad_(True)))
        self.scale_bound = blob["scale_bound"]
        self.setup_functions()

    # ------------------------------------------------------------------ optimizer surgery
    def _swap_param(self, group, new_tensor, moments):
        """Replace the parameter of an optimizer group, carrying (or resetting) its Adam state."""
        old = group["params"][0]
        state = self.optimizer.state.pop(old, None)
        new = nn.Parameter(new_tensor.requires_grad_(True))
        group["params"][0] = new
        if state is not None:
            state["exp_avg"], state["exp_avg_sq"] = moments(state["exp_avg"]), moments(state["exp_avg_sq"])
            self.optimizer.state[new] = state
        return new

    def replace_tensor_to_optimizer(self, tensor, name):
        out = {}
        for group in self.optimizer.param_groups:
            if group["name"] == name:
                out[name] = self._swap_param(group, tensor, lambda m: torch.zeros_like(tensor))
        return out

    def _gather_rows(self, select, extra=None, nsel=None, stats=()):
        """new = cat([old, extra[name]])[select] for every parameter and its Adam moments (zeros for the extra rows),
        plus the `stats` tensors ((old, extra-or-None) pairs), all in ONE launch of r2x_gather_rows.  `select`: int32
        index list on the device (first `nsel` entries valid), a bool mask (converted on the device), or None =
        everything.  Returns ({group name: new Parameter}, [gathered stats], nsel)."""
        n_old = int(self._xyz.shape[0])
        n_extra = 0 if extra is None else int(extra["xyz"].shape[0])
        if select is None:
            nsel = n_old + n_extra
        elif select.dtype in (torch.bool, torch.uint8):
            select, count = compact.select_rows(select)
            nsel = compact.read_counts(count)[0]
        specs, slots = [], []
        for group in self.optimizer.param_groups:
            name = group["name"]
            old = group["params"][0]
            add = None if extra is None else extra[name]
            st = self.optimizer.state.get(old, None)
            specs.append((old, add))
            if st is not None:
                specs.append((st["exp_avg"], None))
                specs.append((st["exp_avg_sq"], None))
            slots.append((group, name, old, st))
        specs += list(stats)
        got = compact.gather_rows(specs, select, nsel)
        out, k = {}, 0
        for group, name, old, st in slots:
            new = nn.Parameter(got[k].requires_grad_(True)); k += 1
            self.optimizer.state.pop(old, None)
            group["params"][0] = new
            if st is not None:
                st["exp_avg"], st["exp_avg_sq"] = got[k], got[k + 1]; k += 2
                self.optimizer.state[new] = st
            out[name] = new
        return out, got[k:], nsel

    def _adopt(self, params):
        for name, attr in _GROUPS:
            setattr(self, attr, params[name])

    def _prune_optimizer(self, mask):
        return self._gather_rows(mask)[0]

    def cat_tensors_to_optimizer(self, tensors_dict):
        return self._gather_rows(None, tensors_dict)[0]

    def prune_points(self, mask):
        params, (accum, denom, radii), _ = self._gather_rows(
            ~mask, stats=[(self.xyz_gradient_accum, None), (self.denom, None), (self.max_radii2D, None)])
        self._adopt(params)
        self.xyz_gradient_accum, self.denom, self.max_radii2D = accum, denom, radii

    def densification_postfix(self, new_xyz, new_densities, new_scaling, new_rotation, new_max_radii2D):
        extra = {"xyz": new_xyz, "density": new_densities, "scaling": new_scaling, "rotation": new_rotation}
        self._adopt(self._gather_rows(None, extra)[0])
        n = self.get_xyz.shape[0]
        self.xyz_gradient_accum = torch.zeros((n, 1), device="cuda")
        self.denom = torch.zeros((n, 1), device="cuda")
        self.max_radii2D = torch.cat([self.max_radii2D, new_max_radii2D], dim=-1)

    def reset_density(self, reset_density=1.0):
        capped = torch.min(self.get_density, torch.ones_like(self.get_density) * reset_density)
        self._density = self.replace_tensor_to_optimizer(self.density_inverse_activation(capped), "density")["density"]

    # ------------------------------------------------------------------ adaptive control
    def _rows(self, t, rows):
        """t[rows] for a bool mask (reference-style call) or an int index list (device-side selection)."""
        return t[rows] if rows.dtype == torch.bool else t.index_select(0, rows)

    def _split_children(self, rows, N):
        """N children per selected Gaussian: positions drawn from the parent (`:430-455`), scale / (0.8 N),
        density / N.  `rows`: bool mask or int index list of the parents."""
        scale = self._rows(self.get_scaling, rows).repeat(N, 1)
        offsets = torch.normal(mean=torch.zeros((scale.size(0), 3), device="cuda"), std=scale)
        frames = build_rotation(self._rows(self._rotation, rows)).repeat(N, 1, 1)
        xyz = torch.bmm(frames, offsets.unsqueeze(-1)).squeeze(-1) + self._rows(self.get_xyz, rows).repeat(N, 1)
        return {"xyz": xyz,
                "density": self.density_inverse_activation(self._rows(self.get_density, rows).repeat(N, 1) * (1 / N)),
                "scaling": self.scaling_inverse_activation(scale / (0.8 * N)),
                "rotation": self._rows(self._rotation, rows).repeat(N, 1)}, self._rows(self.max_radii2D, rows).repeat(N)

    def densify_and_split(self, grads, grad_threshold, densify_scale_threshold, N=2):
        n = self.get_xyz.shape[0]
        padded = torch.zeros((n), device="cuda")
        padded[: grads.shape[0]] = grads.squeeze()
        mask = (padded >= grad_threshold) & (torch.max(self.get_scaling, dim=1).values > densify_scale_threshold)
        children, radii = self._split_children(mask, N)
        self.densification_postfix(children["xyz"], children["density"], children["scaling"], children["rotation"], radii)
        self.prune_points(torch.cat((mask, torch.zeros(N * int(mask.sum()), device="cuda", dtype=bool))))

    def densify_and_clone(self, grads, grad_threshold, densify_scale_threshold):
        mask = (torch.norm(grads, dim=-1) >= grad_threshold) & (
            torch.max(self.get_scaling, dim=1).values <= densify_scale_threshold)
        halved = self.density_inverse_activation(self.get_density[mask] * 0.5)
        twins = (self._xyz[mask], halved, self._scaling[mask], self._rotation[mask], self.max_radii2D[mask])
        self._density[mask] = halved          # the original keeps half of the density too (`:493`)
        self.densification_postfix(*twins)

    def densify_and_prune(self, max_grad, min_density, max_screen_size, max_scale, max_num_gaussians,
                          densify_scale_threshold, bbox=None):
        """Clone + split + every prune rule (`:503-550`) with ONE rebuild of the per-Gaussian storage, selection and
        compaction on the device: two small host reads per call (clone / split counts, survivor count)."""
        grads = self.xyz_gradient_accum / self.denom
        grads = torch.where(grads.isnan(), torch.zeros_like(grads), grads)
        n0 = grads.shape[0]
        extra, extra_radii, drop_parent = None, None, None
        with torch.no_grad():
            density_old = self._density.detach()
            if densify_scale_threshold and (not max_num_gaussians or n0 < max_num_gaussians):
                max_s = torch.max(self.get_scaling, dim=1).values
                hot = torch.norm(grads, dim=-1) >= max_grad
                clone_m = hot & (max_s <= densify_scale_threshold)
                split_m = hot & (max_s > densify_scale_threshold)
                ci, cn = compact.select_rows(clone_m)
                si, sn = compact.select_rows(split_m)
                n_clone, n_split = compact.read_counts(cn, sn)                  # host read 1 of 2
                ci, si = ci[:n_clone].long(), si[:n_split].long()
                # clones: twin rows with half the density; the originals are halved too (`:493`)
                halved = self.density_inverse_activation(self.get_density * 0.5)
                twins = {"xyz": self._xyz.detach().index_select(0, ci), "density": halved.index_select(0, ci),
                         "scaling": self._scaling.detach().index_select(0, ci),
                         "rotation": self._rotation.detach().index_select(0, ci)}
                twin_radii = self.max_radii2D.index_select(0, ci)
                density_old = torch.where(clone_m.unsqueeze(-1), halved, density_old)
                # split children are drawn from the parents (which the clone step did not touch: disjoint masks)
                children, child_radii = self._split_children(si, 2)
                extra = {k: torch.cat((twins[k], children[k]), dim=0) for k in twins}
                extra_radii = torch.cat((twin_radii, child_radii), dim=-1)
                drop_parent = split_m
            # candidate rows after densification: [old | twins | children]
            cat = (lambda old, new: old if new is None else torch.cat((old, new), dim=0))
            xyz = cat(self._xyz.detach(), None if extra is None else extra["xyz"])
            dens = self.density_activation(cat(density_old, None if extra is None else extra["density"]))
            scal = self.scaling_activation(cat(self._scaling.detach(), None if extra is None else extra["scaling"]))
            radii = cat(self.max_radii2D, extra_radii)
            drop = (dens < min_density).squeeze(-1)
            if drop_parent is not None:
                drop = drop | cat(drop_parent, torch.zeros(drop.shape[0] - n0, dtype=torch.bool, device=drop.device))
            if bbox is not None:
                drop = drop | ((xyz < bbox[0].to(xyz.device)) | (xyz > bbox[1].to(xyz.device))).any(dim=1)
            if max_screen_size:
                drop = drop | (radii > max_screen_size)
            if max_scale:
                drop = drop | (scal.max(dim=1).values > max_scale)
            keep, kn = compact.select_rows(~drop)
            n = compact.read_counts(kn)[0]                                      # host read 2 of 2
            # the originals of the clones carry their halved density into the gather
            if density_old is not self._density:
                self._density.data = density_old if density_old.is_contiguous() else density_old.contiguous()
            stats = [(self.max_radii2D, extra_radii)]
            if extra is None:
                stats += [(self.xyz_gradient_accum, None), (self.denom, None)]
            params, got, _ = self._gather_rows(keep, extra, nsel=n, stats=stats)
            self._adopt(params)
            self.max_radii2D = got[0]
            if extra is not None:
                self.xyz_gradient_accum = torch.zeros((n, 1), device="cuda")
                self.denom = torch.zeros((n, 1), device="cuda")
            else:
                self.xyz_gradient_accum, self.denom = got[1], got[2]
        return grads

    def add_densification_stats(self, viewspace_point_tensor, update_filter):
        """Accumulate |dL/dmean2D| of the visible Gaussians (`:552-556`).  Written without boolean-mask indexing
        (which costs a host round trip per call): invisible rows add exactly 0."""
        seen = update_filter.unsqueeze(-1)
        norm = torch.norm(viewspace_point_tensor.grad[:, :2], dim=-1, keepdim=True)
        self.xyz_gradient_accum += torch.where(seen, norm, torch.zeros_like(norm))
        self.denom += seen.to(self.denom.dtype)

    def update_max_radii(self, radii, visibility_filter):
        """max_radii2D[vis] = max(max_radii2D[vis], radii[vis]) (train.py:150-152) without a host round trip."""
        r = radii.to(self.max_radii2D.dtype)
        self.max_radii2D = torch.where(visibility_filter, torch.maximum(self.max_radii2D, r), self.max_radii2D)
