"""RayBendingStyleNerfModel (reference: model/nerf_models/ray_bending_style_nerf_model.py:12-230).

One object model = ray bender + style-modulated NeRF field + bounding box.  Sub-models are chosen through the
config-string registry exactly as in the reference (:36-37).  Besides owning the parameters, the class keeps the PACKED
copy the kernels read (fp32 transposed + fp16 tensor-core slabs), refreshed whenever a parameter changes."""
import os
from typing import Dict, List, Optional, Tuple

import ctypes as C

import torch
import torch.nn as nn

from ...utils.lib_3d.bounding_box import BoundingBox
from ..annealable_positional_encoder import annealing_weights
from ... import _cabi, registry


class RayBendingStyleNerfModel(nn.Module):

    def __init__(self, config: Dict, model_config: Dict):
        super().__init__()
        self.config = config
        self.model_config = model_config
        self.empty_space_alpha = model_config["empty_space_alpha"]
        self.bounding_box = BoundingBox(model_config["bounding_box"])
        self.style_features = model_config["style_features"]
        self.deformation_features = model_config["deformation_features"]
        self.nerf_model_config = self.model_config["nerf_model"]
        self.ray_bender_model_config = self.model_config["ray_bender_model"]
        self.transfer_attributes_to_submodels_configs()
        self.nerf_model = registry.build(self.nerf_model_config["architecture"], config, self.nerf_model_config)
        self.ray_bender = registry.build(self.ray_bender_model_config["architecture"], config, self.ray_bender_model_config)
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._packed_aware = False

    def transfer_attributes_to_submodels_configs(self):
        for current_config in [self.nerf_model_config, self.ray_bender_model_config]:      # reference :39-50
            current_config["bounding_box"] = self.model_config["bounding_box"]
            current_config["empty_space_alpha"] = self.model_config["empty_space_alpha"]
            current_config["style_features"] = self.model_config["style_features"]
            current_config["deformation_features"] = self.model_config["deformation_features"]

    def set_step(self, current_step: int):
        self.ray_bender.set_step(current_step)

    def compute_bounding_box_filtering_mask(self, flat_ray_positions: torch.Tensor) -> torch.Tensor:
        return self.bounding_box.is_inside(flat_ray_positions)

    # ---- C-ABI description -----------------------------------------------------------------------
    def object_desc(self, positions: int, is_static: bool, canonical_pose: bool = False) -> _cabi.PeObjectDesc:
        desc = build_object_desc(self.nerf_model, self.ray_bender, self.bounding_box, self.model_config, positions, is_static,
                                 canonical_pose, self.packed_parameters())
        desc.aware_rounding = 1 if self._packed_aware else 0
        return desc

    def state_tensors(self):
        """Every tensor the packed blob is built from (learnable tensors + the BatchNorm running statistics), collected by
        ATTRIBUTE access: nn.DataParallel replicas (train.py:61) carry their weights as plain attributes and have empty
        ``_parameters`` / ``_buffers``, so ``.parameters()`` would yield nothing there."""
        tensors = [t for _, _, t in self.parameter_slots()]
        head = self.nerf_model.features_head
        for layer in (head[1], head[4]):
            bn = layer.ada_in.normalization
            tensors += [bn.running_mean, bn.running_var]
        return tensors

    def packed_parameters(self) -> torch.Tensor:
        tensors = self.state_tensors()
        # inference packs the weight stream against activation statistics (one calibration pass per parameter version); in training
        # the parameters change every step and the blob is repacked with the data-free zero-sum rounding
        aware = (not self.training) and _aware_shape(self.nerf_model) and os.environ.get("PE_TC_AWARE", "1") != "0"
        key = tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors) + (aware,)
        if self._packed is None or key != self._packed_key:
            self._packed = pack_object(self.nerf_model, self.ray_bender, self.bounding_box, self.model_config, aware=aware)
            self._packed_key = key
            self._packed_aware = aware
        return self._packed

    def parameter_struct(self):
        """PeObjectParams of this model (device pointers of its fp32 tensors) + the list keeping them alive."""
        return _param_struct(self.nerf_model, self.ray_bender)

    def parameter_slots(self):
        """[(PeObjectParamGrads field, index or None, parameter)] — every learnable tensor the backward produces a gradient for."""
        nerf, bender = self.nerf_model, self.ray_bender
        slots = []
        for i, layer in enumerate(nerf.backbone_layers):
            slots += [("backbone_w", i, layer.weight), ("backbone_b", i, layer.bias)]
        if getattr(nerf, "HAS_ALPHA_HEAD", True):
            slots += [("alpha_w", None, nerf.alpha_head.weight), ("alpha_b", None, nerf.alpha_head.bias)]
        head = nerf.features_head
        slots += [("head0_w", None, head[0].weight),
                  ("affine1_w", None, head[1].affine_transform.weight), ("affine1_b", None, head[1].affine_transform.bias),
                  ("head3_w", None, head[3].weight),
                  ("affine2_w", None, head[4].affine_transform.weight), ("affine2_b", None, head[4].affine_transform.bias),
                  ("head6_w", None, head[6].weight), ("head6_b", None, head[6].bias)]
        if bender is not None and bender.KIND == _cabi.BENDER_POSITIONAL:
            for i, layer in enumerate(bender.backbone_layers):
                slots += [("bender_w", i, layer.weight), ("bender_b", i, layer.bias)]
            slots.append(("bender_out_w", None, bender.output_head.weight))
        return slots

    def forward(self, ray_positions: torch.Tensor, ray_origins: torch.Tensor, ray_directions: torch.Tensor, style: torch.Tensor,
                deformation: torch.Tensor, video_indexes: torch.Tensor = None, canonical_pose: bool = False) -> Tuple[torch.Tensor]:
        """(..., P, 3) positions; (..., 3) origins/directions; (..., S) style; (..., D) deformation ->
        (..., P, F) features, (..., P) raw alphas, (..., P, 3) displacements, {} — reference :137-219."""
        P = ray_positions.size(-2)
        org = ray_origins.unsqueeze(-2).expand(list(ray_positions.shape))
        drs = ray_directions.unsqueeze(-2).expand(list(ray_positions.shape))
        sty = style.unsqueeze(-2)
        dfm = deformation.unsqueeze(-2)
        feats, alphas, disp = evaluate_field_on_positions(self.nerf_model, self.ray_bender, ray_positions, org, drs, sty, dfm,
                                                          self.training, canonical_pose, owner=self)
        del P
        return feats, alphas, disp, {}


def _param_struct(nerf, bender) -> Tuple[_cabi.PeObjectParams, list]:
    """Device pointers of the fp32 parameters (kept alive by the returned list)."""
    keep = []

    def p(t):
        if t is None:
            return None
        c = _cabi.f32(t)
        keep.append(c)
        return _cabi.ptr(c)

    ps = _cabi.PeObjectParams()
    for i, layer in enumerate(nerf.backbone_layers):
        ps.backbone_w[i] = p(layer.weight)
        ps.backbone_b[i] = p(layer.bias)
    if getattr(nerf, "HAS_ALPHA_HEAD", True):
        ps.alpha_w = p(nerf.alpha_head.weight)
        ps.alpha_b = p(nerf.alpha_head.bias)
    head = nerf.features_head
    ps.head0_w = p(head[0].weight)
    ps.affine1_w = p(head[1].affine_transform.weight)
    ps.affine1_b = p(head[1].affine_transform.bias)
    ps.bn1_mean = p(head[1].ada_in.normalization.running_mean)
    ps.bn1_var = p(head[1].ada_in.normalization.running_var)
    ps.head3_w = p(head[3].weight)
    ps.affine2_w = p(head[4].affine_transform.weight)
    ps.affine2_b = p(head[4].affine_transform.bias)
    ps.bn2_mean = p(head[4].ada_in.normalization.running_mean)
    ps.bn2_var = p(head[4].ada_in.normalization.running_var)
    ps.head6_w = p(head[6].weight)
    ps.head6_b = p(head[6].bias)
    if bender is not None and bender.KIND == _cabi.BENDER_POSITIONAL:
        for i, layer in enumerate(bender.backbone_layers):
            ps.bender_w[i] = p(layer.weight)
            ps.bender_b[i] = p(layer.bias)
        ps.bender_out_w = p(bender.output_head.weight)
    return ps, keep


def build_object_desc(nerf, bender, bounding_box: BoundingBox, model_config: Dict, positions: int, is_static: bool,
                      canonical_pose: bool, packed: Optional[torch.Tensor]) -> _cabi.PeObjectDesc:
    d = _cabi.PeObjectDesc()
    d.nerf_kind = nerf.KIND
    d.bender_kind = _cabi.BENDER_ZEROED if bender is None else bender.KIND
    d.width, d.layers, d.skip = nerf.layers_width, nerf.backbone_layers_count, nerf.skip_layer_idx
    d.octaves, d.features = nerf.position_encoder.octaves_count, nerf.output_features
    d.style_features = nerf.style_features
    d.deformation_features = model_config.get("deformation_features", 0)
    if d.bender_kind == _cabi.BENDER_POSITIONAL:
        d.b_width, d.b_layers, d.b_skip = bender.layers_width, bender.layers_count, bender.skip_layer_idx
        d.b_octaves = bender.positional_encoder.octaves_count
        w = annealing_weights(bender.positional_encoder.host_step(), d.b_octaves, bender.positional_encoder.num_steps)
        for i in range(d.b_octaves):
            d.b_anneal[i] = float(w[i])
    d.positions = positions
    d.is_static = 1 if is_static else 0
    d.canonical_pose = 1 if canonical_pose else 0
    for i, v in enumerate(bounding_box.as_floats()):
        d.bbox[i] = v
    d.z_near_min = float(model_config.get("z_near_min", 0.0))
    d.z_far_max = float(model_config.get("z_far_max", 0.0))
    d.empty_space_alpha = float(model_config["empty_space_alpha"])
    d.packed = None if packed is None else _cabi.ptr(packed)
    return d


def _aware_shape(nerf) -> bool:
    """The field shape with a tensor-core weight stream (pe_tc_field_ok, csrc/pe_common.cuh)."""
    return (nerf.KIND == _cabi.NERF_ADAIN and nerf.layers_width == 256 and nerf.backbone_layers_count == 8 and nerf.skip_layer_idx == 4
            and nerf.position_encoder.octaves_count == 10 and nerf.output_features == 192)


AWARE_CALIBRATION_POSITIONS = 8192


def activation_moments(nerf, bounding_box: BoundingBox, device) -> List[torch.Tensor]:
    """Second-moment matrices E[a a^T] of the fp16-rounded inputs of every trunk layer and of features_head.0, measured on
    AWARE_CALIBRATION_POSITIONS positions drawn uniformly in the bounding box (fixed seed; the trunk sees positions only, so the
    statistics depend on nothing but the weights and the box).  Packing-time calibration for the activation-aware rounding of the
    tensor-core weight stream (PeObjectParams.backbone_in_moments): a few small fp32 matmuls, not part of the render path."""
    with torch.no_grad():
        g = torch.Generator(device=device).manual_seed(0x5EED)
        dims = bounding_box.dimensions.to(device=device, dtype=torch.float32)
        size = dims[:, 1] - dims[:, 0]
        pos = dims[:, 0] + size * torch.rand((AWARE_CALIBRATION_POSITIONS, 3), generator=g, device=device)
        enc = nerf.position_encoder(pos / size)
        moments = []
        h = enc
        for i, layer in enumerate(nerf.backbone_layers):
            if i == nerf.skip_layer_idx:
                h = torch.cat([h, enc], dim=-1)
            q = h.half().float()
            moments.append((q.t() @ q / q.size(0)).contiguous())
            h = torch.relu(torch.nn.functional.linear(h, layer.weight.float(), layer.bias.float()))
        q = h.half().float()
        moments.append((q.t() @ q / q.size(0)).contiguous())
    return moments


def pack_object(nerf, bender, bounding_box: BoundingBox, model_config: Dict, aware: bool = False) -> torch.Tensor:
    """Packs the parameters into the kernel-side blob (pe_pack_object).  ``aware``: round the tensor-core weight stream against
    activation statistics (``activation_moments``)."""
    device = nerf.backbone_layers[0].weight.device        # attribute access: valid on nn.DataParallel replicas too
    if device.type != "cuda":
        raise _cabi.PeError("parameters must live on a CUDA device: the render path has no CPU implementation")
    L = _cabi.lib()
    desc = build_object_desc(nerf, bender, bounding_box, model_config, 1, True, False, None)
    nbytes = L.pe_packed_bytes(C.byref(desc))
    if nbytes == 0:
        raise _cabi.PeError(f"pe_packed_bytes: {L.pe_last_error().decode()}")
    packed = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        ps, keep = _param_struct(nerf, bender)
        if aware:
            allow = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            try:
                moments = activation_moments(nerf, bounding_box, device)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = allow
            keep.append(moments)
            for i in range(len(nerf.backbone_layers)):
                ps.backbone_in_moments[i] = _cabi.ptr(moments[i])
            ps.head0_in_moments = _cabi.ptr(moments[-1])
        _cabi.check(L.pe_pack_object(C.byref(desc), C.byref(ps), _cabi.ptr(packed), _cabi.current_stream(device)))
    del keep
    return packed


def _split_leading(lead, code: torch.Tensor):
    """Finds how many leading dims of ``lead`` index distinct codes: ``code`` (..., S) must be broadcast (size 1 or stride 0)
    over the remaining ones.  Returns the split index, or None when the code really varies per sample."""
    shape = list(code.shape[:-1])
    if len(shape) != len(lead):
        return None
    split = len(lead)
    while split > 0 and (shape[split - 1] == 1 or code.stride(split - 1) == 0):
        split -= 1
    if shape[:split] != list(lead[:split]):
        return None
    return split


def evaluate_field_on_positions(nerf, bender, positions, origins, directions, style, deformation, training: bool,
                                canonical_pose: bool, owner: Optional[RayBendingStyleNerfModel] = None):
    """Field evaluation on explicit object-space positions (module-level ``forward`` of the nerf models)."""
    from .. import render
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (positions, style, deformation)):
        raise NotImplementedError("backward through the stand-alone field operator is not implemented yet; "
                                  "use ObjectComposer.forward under torch.no_grad()")
    lead = list(positions.shape[:-1])
    split = _split_leading(lead, style)
    if split is not None and deformation is not None:
        s2 = _split_leading(lead, deformation)
        split = None if s2 is None else max(split, s2)
    if split is None:                      # genuinely per-sample codes: every sample is its own "image"
        split = len(lead)
    images = 1
    for v in lead[:split]:
        images *= v
    n = 1
    for v in lead[split:]:
        n *= v
    device = positions.device
    pos = _cabi.f32(positions).reshape(images, n, 3)
    S = style.size(-1)
    sty = _cabi.f32(style.expand(lead + [S]).reshape(images, n, S)[:, 0])
    dfm = None
    if deformation is not None:
        D = deformation.size(-1)
        dfm = _cabi.f32(deformation.expand(lead + [D]).reshape(images, n, D)[:, 0])
    org = _cabi.f32(origins.expand(lead + [3]).reshape(images, n, 3)[:, 0])
    drs = _cabi.f32(directions.expand(lead + [3]).reshape(images, n, 3))
    if owner is not None:
        packed, model_config, bbox = owner.packed_parameters(), owner.model_config, owner.bounding_box
    else:
        model_config, bbox = nerf.model_config, nerf.bounding_box
        packed = pack_object(nerf, bender, bbox, model_config)
    desc = build_object_desc(nerf, bender, bbox, model_config, 1, True, canonical_pose, packed)
    feats, alphas, disp = render.field_on_positions(desc, images, n, pos, org, drs, sty, dfm, training, device)
    F = nerf.output_features
    return feats.reshape(lead + [F]), alphas.reshape(lead), disp.reshape(lead + [3])


def model(config, model_config):
    return RayBendingStyleNerfModel(config, model_config)
