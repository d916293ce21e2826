"""ObjectComposer (reference: model/object_composer.py:18-892) on the B200 fused render path.

Same constructor (``config``), sub-module names (``object_models_coarse.{m}....`` — checkpoints load unchanged),
``forward`` / ``set_step`` signatures and result dictionary.  Where the reference runs ~10^2 ATen kernels per object and
materialises (…, R, P, 192) tensors, ``forward`` is a handful of launches: per object a style prologue and one fused
field kernel (sampling + encoding + MLP [+ integration]), then one compositing kernel."""
from typing import Dict, List

import torch
import torch.nn as nn

from .. import _cabi, registry
from .utils.object_ids_helper import ObjectIDsHelper
from . import render


class ObjectComposer(nn.Module):

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.object_models_coarse = nn.ModuleList(self.create_object_models(fine=False))
        self.object_models_fine = nn.ModuleList(self.create_object_models(fine=True))
        self.apply_activation = self.config["model"]["apply_activation"]
        if self.object_models_coarse[0].model_config["nerf_model"]["output_features"] != 3 and self.apply_activation:
            raise Exception("The application of activations to the nerf output is requested, but the model seem not to output colors directly. Please make sure this is the behavior you desire")
        self.object_id_helper = ObjectIDsHelper(self.config)
        # arithmetic of the shipped-shape fields (others always run the exact fp32 CUDA-core kernel):
        #   "fp16x3" tensor cores, weights and activations split hi+lo — fp32-class parity (default)
        #   "fp16x2" tensor cores, weights split hi+lo          "fp16" tensor cores, single pass (fastest)
        #   "fp32"   CUDA cores only
        self.precision = self.config["model"].get("b200_precision", "mixed")
        # diagnostic switch: also return the per-sample raw alphas of every object under results["coarse"]["object_k"]["raw_alphas"]
        self.return_raw_alphas = False
        # Hutchinson divergence of the ray benders' displacement fields (reference :582-601) in training calls.  Off by default: its loss
        # weight is 0 in every shipped config (the trainers only log the value), it is random by construction, and it costs one more
        # pass through every ray bender; ``model.b200_divergence: True`` evaluates it (forward value only -- a non-zero
        # divergence_loss_lambda would need the second derivative of the bender, which this path does not provide, so it raises).
        self.compute_divergence = bool(self.config["model"].get("b200_divergence", False))
        lam = self.config.get("training", {}).get("loss_weights", {}).get("divergence_loss_lambda", 0.0) if isinstance(self.config, dict) else 0.0
        if lam:
            raise NotImplementedError("divergence_loss_lambda != 0 needs the second derivative of the ray bender (reference: create_graph=True, "
                                      ":598); the B200 render path evaluates the divergence forward only")

    def create_object_models(self, fine: bool) -> List[nn.Module]:
        object_models = []
        for current_object_config in self.config["model"]["object_models"]:
            if fine and "use_fine" in current_object_config and current_object_config["use_fine"] == False:  # noqa: E712
                current_model = None
            else:
                current_model = registry.build(current_object_config["architecture"], self.config, current_object_config)
            object_models.append(current_model)
        return object_models

    def set_step(self, current_step: int):
        for current_object_model in self.object_models_coarse:
            current_object_model.set_step(current_step)
        for current_object_model in self.object_models_fine:
            if current_object_model is not None:
                current_object_model.set_step(current_step)

    def _any_parameter_requires_grad(self) -> bool:
        """Attribute walk instead of ``self.parameters()``: on nn.DataParallel replicas the weights are plain tensor attributes."""
        models = list(self.object_models_coarse) + [m for m in self.object_models_fine if m is not None]
        return any(t.requires_grad for m in models for t in m.state_tensors())

    def _descs(self, canonical_pose: bool, fine: bool = False):
        """One PeObjectDesc per object instance.  ``fine``: the fine model of each instance on positions_count_coarse +
        positions_count_fine samples per ray (the coarse samples merged with the resampled ones, reference :563-566)."""
        helper = self.object_id_helper
        descs = []
        for object_idx in range(helper.objects_count):
            model_idx = helper.model_idx_by_object_idx(object_idx)
            coarse = self.object_models_coarse[model_idx]
            positions = coarse.model_config["positions_count_coarse"]
            m = coarse
            if fine:
                m = self.object_models_fine[model_idx]
                positions += m.model_config["positions_count_fine"]
            descs.append(m.object_desc(positions, helper.is_static(model_idx), canonical_pose))
        return descs

    def _uses_fine(self) -> bool:
        """The reference iterates over the result keys of object 0 (:847): a fine model on object 0 makes every object need one."""
        helper = self.object_id_helper
        fine = [self.object_models_fine[helper.model_idx_by_object_idx(k)] is not None for k in range(helper.objects_count)]
        if fine[0] and not all(fine):
            raise Exception("use_fine must be set for every object model or for none (the reference composes the 'fine' results of all objects)")
        return fine[0]

    @staticmethod
    def compute_raywise_object_z_bounds(ray_origins: torch.Tensor, ray_directions: torch.Tensor, bounding_box, object_validity: torch.Tensor):
        """Reference :104-151 (slab test per ray; rays that miss the box or belong to an absent object get z_near = z_far = 0).  The
        kernels evaluate this in registers; the torch form serves the fine pass, whose merged ray parameters are built on the host side."""
        eps = 1e-6
        corners = bounding_box.get_corner_points()[[0, 6]].to(ray_directions.device)
        z = (corners - ray_origins.unsqueeze(-2)).unsqueeze(-3) / (ray_directions.unsqueeze(-2) + eps)
        z_near = z.min(dim=-2)[0].max(dim=-1)[0]
        z_far = z.max(dim=-2)[0].min(dim=-1)[0]
        mask = torch.logical_or(z_far <= z_near, torch.logical_not(object_validity.unsqueeze(-1).expand_as(z_far)))
        return torch.where(mask, torch.zeros_like(z_near), z_near), torch.where(mask, torch.zeros_like(z_far), z_far)

    def _fine_ray_parameters(self, ray_origins, ray_directions, focal_normals, transformation_matrix_w2o, object_in_scene, perturb: bool,
                             rand, coarse_results) -> List[torch.Tensor]:
        """RayHelper.create_ray_positions_weighted per object (reference :563-566): the object's coarse ray parameters (a differentiable
        function of the rays, evaluated here with the torch helpers) merged with ``positions_count_fine`` inverse-CDF samples of its own
        coarse compositing weights (detached, like the reference), sorted.  (With ``perturb`` the reference draws a separate raw-alpha
        noise for these weights, :553; here they are the weights the coarse pass returns.)"""
        from ..utils.lib_3d.ray_helper import RayHelper
        helper = self.object_id_helper
        merged = []
        for k in range(helper.objects_count):
            model_idx = helper.model_idx_by_object_idx(k)
            coarse, fine = self.object_models_coarse[model_idx], self.object_models_fine[model_idx]
            cfg = coarse.model_config
            if cfg["positions_count_coarse"] < 3:
                raise Exception("fine sampling needs at least 3 coarse samples per ray (sample_pdf over the interior bins, ray_helper.py:1336)")
            o, d, _ = RayHelper.transform_rays(ray_origins, ray_directions, focal_normals, transformation_matrix_w2o[..., k])
            z_near, z_far = self.compute_raywise_object_z_bounds(o, d, coarse.bounding_box, object_in_scene[..., k])
            z_near = torch.clamp(z_near, min=cfg["z_near_min"], max=cfg["z_far_max"])
            z_far = torch.clamp(z_far, min=cfg["z_near_min"], max=cfg["z_far_max"])
            t_coarse = RayHelper.ray_parameters(z_near, z_far, cfg["positions_count_coarse"], rand[k] if perturb else None)
            weights = coarse_results[f"object_{k}"]["weights"].detach()
            mid = (t_coarse[..., 1:] + t_coarse[..., :-1]) / 2
            t_new = RayHelper.sample_pdf(mid.detach(), weights[..., 1:-1], fine.model_config["positions_count_fine"], perturb).detach()
            merged.append(torch.sort(torch.cat([t_coarse, t_new], dim=-1), dim=-1)[0])
        return merged

    def forward(self, ray_origins: torch.Tensor, ray_directions: torch.Tensor, focal_normals: torch.Tensor,
                transformation_matrix_w2o: torch.Tensor, style: torch.Tensor, deformation: torch.Tensor, object_in_scene: torch.Tensor,
                perturb: bool, video_indexes: torch.Tensor = None, canonical_pose: bool = False, rand=None, noise=None,
                peer_features=None, divergence_noise=None, handoff=None, global_only: bool = False) -> Dict:
        """Same contract as the reference (:786-812).  ``rand`` / ``noise`` optionally supply the perturbation tensors
        (otherwise drawn from torch's generator), so that a run can be reproduced sample for sample.  ``peer_features`` (inference):
        extra destinations of the composed scene's feature grid -- the fused all-gather of ``sharding.PeerGather``.
        ``divergence_noise`` ({"coarse": [e_k (..., R, P_k, 3) or None per object]}): the Hutchinson probe vectors (``compute_divergence``).
        ``handoff`` (inference): ``(strides, (H, W), channels)`` of the multiresolution decoder (reference: fold_strided_tensors +
        split_features_by_layer + CHW permute, environment_model_multiresolution_backpropagated_autoencoder.py:29-99) for a frame whose rays
        are the concatenated strided grids: ``results[...]["global"]["feature_grids"]`` holds one (..., channels_i, H/s_i, W/s_i) grid per
        stride.  In multi-object scenes the compositor writes them directly and reads only each ray's own channel range of the
        per-sample features (``integrated_features`` of the composed scene is then None); otherwise they are folded from it.
        ``global_only`` (inference, several objects): only ``results[...]["global"]`` is produced -- what the decoder path reads
        (environment_model_multiresolution_backpropagated_decoder.py:84); the ``object_k`` entries hold ``extra_outputs`` only."""
        objects_count = self.object_id_helper.objects_count
        if transformation_matrix_w2o.size(-1) != objects_count:
            raise Exception(f"Transformation matrix must specifies transformations for"
                            f"({transformation_matrix_w2o.size(-1)}) objects instead of ({objects_count})")
        if self.precision not in _cabi.PRECISIONS:
            raise Exception(f"unknown b200_precision '{self.precision}'")
        needs_grad = torch.is_grad_enabled() and (
            self._any_parameter_requires_grad()
            or any(torch.is_tensor(t) and t.requires_grad for t in (ray_origins, ray_directions, transformation_matrix_w2o, style, deformation)))
        helper = self.object_id_helper
        use_fine = self._uses_fine()
        record = needs_grad and not getattr(self, "allow_forward_without_grad", False)
        if (handoff is not None or global_only) and record:
            raise Exception("handoff / global_only are inference features: call under torch.no_grad()")
        if global_only and use_fine:
            raise Exception("global_only is not available with fine models (the fine pass resamples from every object's coarse weights)")
        if use_fine and perturb and rand is None:
            # the fine pass re-derives the coarse ray parameters from the same stratified jitter the coarse kernels used
            lead_r = list(ray_directions.shape[:-1])
            counts = [d.positions for d in self._descs(canonical_pose)]
            rand = [torch.rand(lead_r + [p], device=ray_directions.device) for p in counts]
            noise = {f"object_{k}": torch.randn(lead_r + [p], device=ray_directions.device) for k, p in enumerate(counts)}
            noise["global"] = torch.randn(lead_r + [sum(counts)], device=ray_directions.device)
        results = {}

        def run(model_type, model_list, descs, **extra):
            models = [model_list[helper.model_idx_by_object_idx(k)] for k in range(objects_count)] if record else None
            bn_running: List = []
            if self.compute_divergence and self.training and record:
                # reference :592-593: only in training and only where the displacements carry a graph (positional ray benders)
                shape = list(ray_directions.shape[:-1])
                given = divergence_noise.get(model_type) if divergence_noise else None
                extra["divergence_noise"] = [
                    (given[k] if given is not None else torch.randn(shape + [d.positions, 3], device=ray_directions.device))
                    if d.bender_kind == _cabi.BENDER_POSITIONAL else None for k, d in enumerate(descs)]
            res = render.render_scene(descs, helper.static_objects_count, ray_origins,
                                      ray_directions, transformation_matrix_w2o, style, deformation, object_in_scene, perturb,
                                      self.training, self.config["model"].get("fix_object_overlaps", True), self.apply_activation,
                                      _cabi.PRECISIONS[self.precision], bn_running=bn_running,
                                      return_raw_alphas=self.return_raw_alphas, models=models, global_only=global_only, **extra)
            if self.training:
                with torch.no_grad():
                    self._update_running_statistics(bn_running, model_list)
            results[model_type] = {}
            for k in range(objects_count):
                r = res[f"object_{k}"]
                r["extra_outputs"] = {}
                results[model_type][f"object_{k}"] = r
            results[model_type]["global"] = res["global"]
            return res

        coarse = run("coarse", self.object_models_coarse, self._descs(canonical_pose), rand=rand, noise=noise,
                     peer_features=None if use_fine else peer_features, handoff=None if use_fine else handoff)
        if use_fine:
            # hierarchical pass (reference :561-578): the fine models on the merged ray parameters, composed like the coarse results
            sample_t = self._fine_ray_parameters(ray_origins, ray_directions, focal_normals, transformation_matrix_w2o, object_in_scene,
                                                 perturb, rand, coarse)
            run("fine", self.object_models_fine, self._descs(canonical_pose, fine=True), sample_t=sample_t, peer_features=peer_features,
                handoff=handoff)
        if handoff is not None:
            # scenes whose grid does not come out of the compositor (one object: the fused kernel; training): fold from the (R, F) tensor
            from ..utils.lib_3d.ray_helper import RayHelper
            g = results["fine" if use_fine else "coarse"]["global"]
            if "feature_grids" not in g:
                g["feature_grids"] = RayHelper.fold_feature_grids(g["integrated_features"], handoff[0], handoff[1], handoff[2])
        # dummy tensor the reference adds for nn.DataParallel's hook handling (:889-890)
        results["pytorch_hook"] = torch.zeros((1, 1, 1, 1, 1, 1, 1, 1, 1), device=ray_directions.device)
        return results

    @staticmethod
    def compute_expected_positions(ray_positions: torch.Tensor, ray_displacements: torch.Tensor, weights: torch.Tensor, eps=1e-8):
        """Reference :603-622: weighted average of the bent sample positions along each ray (weights detached)."""
        weights = weights.detach().unsqueeze(-1)
        return ((ray_positions + ray_displacements) * weights).sum(dim=-2) / (weights.sum(dim=-2) + eps)

    def forward_expected_positions(self, ray_origins: torch.Tensor, ray_directions: torch.Tensor, focal_normals: torch.Tensor,
                                   transformation_matrix_w2o: torch.Tensor, style: torch.Tensor, deformation: torch.Tensor,
                                   object_in_scene: torch.Tensor, object_id: int, perturb: bool, video_indexes: torch.Tensor = None,
                                   canonical_pose: bool = False, rand=None, noise=None) -> Dict:
        """Reference :624-722 for ONE object instance: inputs are that object's (..., 4, 4) pose, (..., S) style, (..., D) deformation
        and (...) presence flag.  Returns {"coarse": (expected_positions (..., R, 3) in object space, opacity (..., R))}.

        The samples (ray parameter t, displacement, compositing weights) come from the same kernels as ``forward`` in a single-object
        scene; the weighted average itself is three small tensor ops.  With autograd on, the bent sample positions are a differentiable
        output of the render node: their upstream gradient goes back through the ray bender (parameters, deformation code) and the ray
        geometry (PeOutGrads.bent_positions); the weights are detached like the reference's (:614) and the opacity is differentiable."""
        needs_grad = torch.is_grad_enabled() and (self._any_parameter_requires_grad() or any(
            torch.is_tensor(t) and t.requires_grad for t in (ray_origins, ray_directions, transformation_matrix_w2o, style, deformation)))
        from ..utils.lib_3d.ray_helper import RayHelper
        helper = self.object_id_helper
        model_idx = helper.model_idx_by_object_idx(object_id)
        m = self.object_models_coarse[model_idx]
        if self.object_models_fine[model_idx] is not None:
            raise NotImplementedError("forward_expected_positions with a fine model (reference :698-720) is not provided")
        desc = m.object_desc(m.model_config["positions_count_coarse"], helper.is_static(model_idx), canonical_pose)
        if needs_grad:
            res = render.render_scene([desc], 1 if helper.is_static(model_idx) else 0, ray_origins, ray_directions,
                                      transformation_matrix_w2o.unsqueeze(-1), style.unsqueeze(-1), deformation.unsqueeze(-1),
                                      object_in_scene.unsqueeze(-1), perturb, self.training, False, self.apply_activation,
                                      _cabi.PRECISIONS[self.precision], rand=rand, noise=noise, models=[m], bent_gradients=True)["object_0"]
            w = res["weights"].detach().unsqueeze(-1)
            expected = (res["bent_positions"] * w).sum(dim=-2) / (w.sum(dim=-2) + 1e-8)
            return {"coarse": (expected, res["opacity"])}
        res = render.render_scene([desc], 1 if helper.is_static(model_idx) else 0, ray_origins, ray_directions,
                                  transformation_matrix_w2o.unsqueeze(-1), style.unsqueeze(-1), deformation.unsqueeze(-1),
                                  object_in_scene.unsqueeze(-1), perturb, self.training, False, self.apply_activation,
                                  _cabi.PRECISIONS[self.precision], rand=rand, noise=noise, return_samples=True)["object_0"]
        origins_o, directions_o, _ = RayHelper.transform_rays(ray_origins, ray_directions, focal_normals, transformation_matrix_w2o)
        positions = origins_o.unsqueeze(-2).unsqueeze(-2) + directions_o.unsqueeze(-2) * res["positions_t"].unsqueeze(-1)
        expected = self.compute_expected_positions(positions, res["displacements"], res["weights"])
        return {"coarse": (expected, res["opacity"])}

    def _update_running_statistics(self, bn_running, model_list=None):
        """BatchNorm running-stat update of the two AdaIn layers, in object order like the reference's sequential
        per-instance model calls (a model shared by two instances is updated twice)."""
        helper = self.object_id_helper
        model_list = self.object_models_coarse if model_list is None else model_list
        # grouped by momentum and applied with multi-tensor ops (a handful of launches instead of five per BatchNorm layer); a model
        # shared by several instances appears once per instance, so its updates stay sequential: one group per repetition
        seen = {}
        groups = {}
        for object_idx, (b1, b2) in enumerate(bn_running):
            model_idx = helper.model_idx_by_object_idx(object_idx)
            rep = seen.get(model_idx, 0)
            seen[model_idx] = rep + 1
            head = model_list[model_idx].nerf_model.features_head
            for layer, b in ((head[1], b1), (head[4], b2)):
                bn = layer.ada_in.normalization
                momentum = bn.momentum if bn.momentum is not None else 0.1
                g = groups.setdefault((rep, momentum), ([], [], []))
                g[0].extend([bn.running_mean, bn.running_var])
                g[1].extend([b[0], b[1]])
                g[2].append(bn.num_batches_tracked)
        for (rep, momentum), (running, batch, counters) in sorted(groups.items(), key=lambda kv: kv[0][0]):
            torch._foreach_mul_(running, 1.0 - momentum)
            torch._foreach_add_(running, batch, alpha=momentum)
            torch._foreach_add_(counters, 1)


def model(config):
    return ObjectComposer(config)
