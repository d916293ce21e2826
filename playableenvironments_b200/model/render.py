"""Host-side launch of the fused render path through the C ABI (``pe_render_forward``).

Everything here is plumbing: flatten the leading (B, O, C) dims into ``images``, hand raw device pointers of torch
tensors to the library on the caller's current stream, wrap the outputs into the reference's nested result dict.
No arithmetic of the render path happens in Python."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch

from .. import _cabi

INTEGRATED_KEYS = ("integrated_features", "opacity", "weights", "depth", "disparity",
                   "integrated_displacements_magnitude", "integrated_divergence")

_launches = 0       # kernels launched by render calls in this process (bench.py reports it)


def take_launch_count() -> int:
    return int(_cabi.lib().pe_take_launch_count())


class _Workspace:
    """One cached workspace per device and stream: pe_render_forward never allocates."""
    _cache: Dict = {}

    @classmethod
    def get(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            cls._cache[key] = buf
        return buf


def _alloc_integrated(images_shape: Sequence[int], rays: int, P: int, F: int, device) -> Dict[str, torch.Tensor]:
    lead = list(images_shape) + [rays]
    e = lambda *s: torch.empty(lead + list(s), dtype=torch.float32, device=device)
    return {"integrated_features": e(F), "opacity": e(), "weights": e(P), "depth": e(), "disparity": e(),
            "integrated_displacements_magnitude": e(), "integrated_divergence": e()}


def _fill(struct: _cabi.PeIntegrated, tensors: Dict[str, torch.Tensor]):
    for k in INTEGRATED_KEYS:
        setattr(struct, k, _cabi.ptr(tensors[k]))


DIFF_KEYS = ("integrated_features", "opacity", "weights", "depth", "disparity", "integrated_displacements_magnitude")


def _flatten_inputs(K: int, ray_origins, ray_directions, w2o, style, deformation, object_in_scene):
    """(B, O, C) leading dims -> ``images``; plain torch ops, so autograd (when it is on) maps gradients back to the caller's tensors."""
    lead = list(ray_directions.shape[:-2])
    rays = ray_directions.size(-2)
    images = 1
    for v in lead:
        images *= v
    f32 = lambda t: t.to(torch.float32)
    origins = f32(ray_origins).expand(lead + [3]).reshape(images, 3).contiguous()
    dirs = f32(ray_directions).reshape(images, rays, 3).contiguous()
    # (..., 4, 4, objects) -> [images][objects][3][4]
    m = f32(w2o).expand(lead + [4, 4, K]).reshape(images, 4, 4, K)[:, :3, :, :].permute(0, 3, 1, 2).contiguous()
    S, D = style.size(-2), deformation.size(-2)
    sty = f32(style).expand(lead + [S, style.size(-1)]).reshape(images, S, -1)
    dfm = f32(deformation).expand(lead + [D, deformation.size(-1)]).reshape(images, D, -1)
    # objects-major copies, one kernel each; unbind (not K selects) so that autograd returns the K gradients with one stack
    styles = list(sty.permute(2, 0, 1).contiguous().unbind(0))[:K]
    deforms = list(dfm.permute(2, 0, 1).contiguous().unbind(0))[:K]
    ois = object_in_scene.expand(lead + [object_in_scene.size(-1)]).reshape(images, -1)[:, :K].to(torch.uint8).contiguous()
    return lead, images, rays, origins, dirs, m, styles, deforms, ois


def _scene_struct(meta, images: int, rays: int) -> _cabi.PeScene:
    scene = _cabi.PeScene()
    descs = meta["descs"]
    scene.images, scene.rays, scene.objects, scene.static_objects = images, rays, len(descs), meta["static_objects"]
    scene.perturb, scene.training = int(bool(meta["perturb"])), int(bool(meta["training"]))
    scene.fix_object_overlaps, scene.apply_activation = int(bool(meta["fix_object_overlaps"])), int(bool(meta["apply_activation"]))
    scene.precision, scene.explicit_positions = meta["precision"], 0
    scene.explicit_t = 1 if meta.get("sample_t") is not None else 0
    scene.divergence = 1 if meta.get("divergence_noise") is not None else 0
    scene.bent_gradients = 1 if meta.get("bent_gradients") else 0
    for k, d in enumerate(descs):
        scene.object[k] = d
    return scene


def _inputs_struct(meta, lead, rays, origins, dirs, w2o, styles, deforms, keep) -> _cabi.PeInputs:
    descs = meta["descs"]
    ins = _cabi.PeInputs()
    ins.ray_origins, ins.ray_directions, ins.w2o = _cabi.ptr(origins), _cabi.ptr(dirs), _cabi.ptr(w2o)
    for k in range(len(descs)):
        ins.style[k] = _cabi.ptr(styles[k])
        ins.deformation[k] = _cabi.ptr(deforms[k])
    ins.object_in_scene = _cabi.ptr(meta["ois"])
    sample_t = meta.get("sample_t")
    if sample_t is not None:          # fine pass: explicit ray parameters per object (object_composer.py:563-578)
        for k, d in enumerate(descs):
            t = _cabi.f32(sample_t[k]).reshape(-1)
            if t.numel() != dirs.size(0) * rays * d.positions:
                raise _cabi.PeError(f"sample_t[{k}]: expected (..., {rays}, {d.positions}) ray parameters")
            keep.append(t)
            ins.sample_t[k] = _cabi.ptr(t)
    div_noise = meta.get("divergence_noise")
    if div_noise is not None:         # Hutchinson divergence (object_composer.py:582-601): e per sample of every positional ray bender
        params = (_cabi.PeObjectParams * _cabi.PE_MAX_OBJECTS)()
        for k, m in enumerate(meta["models"]):
            ps, kp = m.parameter_struct()
            keep.append(kp)
            params[k] = ps
            if div_noise[k] is not None:
                e = _cabi.f32(div_noise[k]).reshape(-1)
                if e.numel() != dirs.size(0) * rays * descs[k].positions * 3:
                    raise _cabi.PeError(f"divergence_noise[{k}]: expected (..., {rays}, {descs[k].positions}, 3)")
                keep.append(e)
                ins.divergence_noise[k] = _cabi.ptr(e)
        keep.append(params)
        ins.divergence_params = C.cast(params, C.c_void_p)
    if meta["perturb"]:
        rand, noise = meta["rand"], meta["noise"]

        def dev(t, shape):
            c = _cabi.f32(t.expand(shape) if list(t.shape) != list(shape) else t).reshape(-1)
            keep.append(c)
            return c

        for k, d in enumerate(descs):
            if sample_t is None:
                ins.rand[k] = _cabi.ptr(dev(rand[k], lead + [rays, d.positions]))
            ins.noise[k] = _cabi.ptr(dev(noise[f"object_{k}"], lead + [rays, d.positions]))
        ins.noise_global = _cabi.ptr(dev(noise["global"], lead + [rays, sum(d.positions for d in descs)]))
    return ins


def _save_forward(meta) -> bool:
    """Training calls in a tensor-core mode keep the forward's workspace (per-sample tensors, masks, BatchNorm sums) for the
    backward instead of recomputing the forward there (PE_SAVE_FORWARD=0: recompute; the exact fp32 mode always recomputes,
    in the fp32-class tensor-core mode)."""
    import os
    if meta.get("return_samples"):      # per-sample outputs are redirected to the caller's tensors: the workspace copy would be stale
        return False
    return meta["precision"] != _cabi.PRECISION_FP32 and os.environ.get("PE_SAVE_FORWARD", "1") != "0"


def handoff_supported(hand, rays: int, features: int) -> bool:
    """(strides, (H, W), channels): the frame's rays are the concatenation of the strided grids, channel ranges are multiples of 32."""
    strides, (H, W), channels = hand
    return (0 < len(strides) <= _cabi.PE_MAX_HANDOFF and len(strides) == len(channels) and all(c % 32 == 0 and c > 0 for c in channels)
            and sum(channels) <= features and sum((H // st) * (W // st) for st in strides) == rays)


def _wants_tile_counts(descs, images: int, rays: int) -> bool:
    """Exact backward tile counts (pe_forward_tile_counts) cost the backward a host-side wait for the forward's last copy -- nothing when a
    loss sits between the two, the launch latency of the backward when it follows immediately.  They pay when the worst case (every
    slot inside its box) does not fit the activation stash: the backward would run in batches and repeat its recompute in each BatchNorm
    phase.  PE_BWD_TILE_COUNTS=1 / 0 forces them on / off."""
    env = os.environ.get("PE_BWD_TILE_COUNTS")
    if env is not None:
        return env != "0"
    cap = int(os.environ.get("PE_BWD_TC_MAX_TILES", "12288"))
    return any(images * ((rays * d.positions + 127) // 128) > cap for d in descs)


def _launch_forward(meta, lead, origins, dirs, w2o, styles, deforms, saved: Optional[List] = None, alias_single: bool = False) -> Dict:
    """``saved``: a list that receives the kept forward workspace (autograd path, see ``_save_forward``).
    ``alias_single`` (inference): in a scene with ONE object instance the composed scene is that object -- same samples, same order, same
    arithmetic (object_composer.py:399-447 with a single list) -- so ``results["global"]`` shares the object's tensors instead of a second
    copy written by the kernel (halves the HBM bytes of the headline frame: 768 + 512 B per ray once instead of twice)."""
    device = dirs.device
    L = _cabi.lib()
    descs = meta["descs"]
    images, rays = dirs.size(0), dirs.size(1)
    F = descs[0].features
    scene = _scene_struct(meta, images, rays)
    keep: List = []
    ins = _inputs_struct(meta, lead, rays, origins, dirs, w2o, styles, deforms, keep)
    outs = _cabi.PeOutputs()
    results: Dict = {}
    # inference callers that read the composed scene only (the decoder path: environment_model_multiresolution_backpropagated_decoder.py:84):
    # no per-object outputs are allocated and the compositor does not integrate the objects' own lists
    global_only = bool(meta.get("global_only")) and len(descs) > 1 and saved is None and not meta["training"]
    for k, d in enumerate(descs):
        r = {} if global_only else _alloc_integrated(lead, rays, d.positions, F, device)
        if not global_only:
            _fill(outs.object[k], r)
        results[f"object_{k}"] = r
        if meta.get("return_raw_alphas"):   # diagnostic: per-sample raw alphas (first return value family of the object models)
            r["raw_alphas"] = torch.empty(lead + [rays, d.positions], dtype=torch.float32, device=device)
            outs.raw_alphas[k] = _cabi.ptr(r["raw_alphas"])
        if meta.get("return_samples"):      # per-sample ray parameter t and displacement (forward_expected_positions)
            r["positions_t"] = torch.empty(lead + [rays, d.positions], dtype=torch.float32, device=device)
            r["displacements"] = torch.zeros(lead + [rays, d.positions, 3], dtype=torch.float32, device=device)
            outs.positions_t[k] = _cabi.ptr(r["positions_t"])
            outs.displacements[k] = _cabi.ptr(r["displacements"])
        if meta["training"] and meta.get("bn_running") is not None:
            b1 = torch.empty((2, d.width), dtype=torch.float32, device=device)
            b2 = torch.empty((2, d.width // 2), dtype=torch.float32, device=device)
            outs.bn1_running[k], outs.bn2_running[k] = _cabi.ptr(b1), _cabi.ptr(b2)
            meta["bn_running"].append((b1, b2))
    if alias_single and len(descs) == 1 and not meta["perturb"]:
        results["global"] = dict(results["object_0"])
        results["global"].pop("raw_alphas", None)
        results["global"].pop("positions_t", None)
        results["global"].pop("displacements", None)
    else:
        g = _alloc_integrated(lead, rays, sum(d.positions for d in descs), F, device)
        _fill(outs.global_, g)
        results["global"] = g
        hand = meta.get("handoff")
        if hand is not None and len(descs) > 1 and not meta["perturb"] and handoff_supported(hand, rays, F):
            # decoder hand-off written by the compositor (PeHandoff): per strided grid its own channel range, channels first; the
            # (R, F) feature tensor of the composed scene is neither accumulated in full nor written
            strides, (H, W), channels = hand
            grids, r0, c0 = [], 0, 0
            outs.handoff.segments = len(strides)
            for q, (st, ch) in enumerate(zip(strides, channels)):
                gh, gw = H // st, W // st
                grid = torch.empty(lead + [ch, gh, gw], dtype=torch.float32, device=device)
                outs.handoff.ray_begin[q], outs.handoff.ray_count[q] = r0, gh * gw
                outs.handoff.channel_begin[q], outs.handoff.channel_count[q] = c0, ch
                outs.handoff.grid[q] = _cabi.ptr(grid)
                grids.append(grid)
                r0, c0 = r0 + gh * gw, c0 + ch
            outs.global_.integrated_features = None
            g["integrated_features"] = None
            g["feature_grids"] = grids
    peers = meta.get("peer_features") or []
    if peers:
        # fused all-gather: extra destinations of the scene's feature grid (peer-mapped buffers of the other GPUs, sharding.PeerGather)
        if len(peers) > _cabi.PE_MAX_PEERS:
            raise _cabi.PeError(f"at most {_cabi.PE_MAX_PEERS} peer destinations")
        outs.peers = len(peers)
        for i, t in enumerate(peers):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != images * rays * F:
                raise _cabi.PeError("peer_features: contiguous fp32 tensors of images * rays * features elements")
            outs.peer_features[i] = t.data_ptr()
    with torch.cuda.device(device):
        if saved is not None:
            scene.keep_samples = 1
        nbytes = L.pe_workspace_bytes(C.byref(scene))
        if nbytes == 0:
            raise _cabi.PeError(f"pe_workspace_bytes: {L.pe_last_error().decode()}")
        if saved is not None:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)       # lives with the autograd node, not in the call cache
            saved.append(ws)
        else:
            ws = _Workspace.get(device, nbytes)
        _cabi.check(L.pe_render_forward(C.byref(scene), C.byref(ins), C.byref(outs), _cabi.ptr(ws), ws.numel(),
                                        _cabi.current_stream(device)))
        if saved is not None and _wants_tile_counts(descs, images, rays) and not torch.cuda.is_current_stream_capturing():
            # how many tiles the tensor-core backward will walk per object, counted on the kept masks and copied to pinned host memory
            # behind the forward: by the time backward runs (the loss sits in between) the copy has landed, and the backward sizes its
            # stash and batch count exactly instead of for the worst case (what autograd knows from the shapes of its saved tensors)
            counts = torch.empty(_cabi.PE_MAX_OBJECTS, dtype=torch.int64, device=device)
            _cabi.check(L.pe_forward_tile_counts(C.byref(scene), _cabi.ptr(ws), ws.numel(), _cabi.ptr(counts), _cabi.current_stream(device)))
            host = torch.empty(_cabi.PE_MAX_OBJECTS, dtype=torch.int64, pin_memory=True)
            host.copy_(counts, non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(torch.cuda.current_stream(device))
            saved.append((host, landed, counts))
    del keep
    return results


class RenderFunction(torch.autograd.Function):
    """ObjectComposer.forward as ONE autograd node.  ``backward`` hands the upstream gradients of every integrated output to
    ``pe_render_backward``, which recomputes the forward on the device and back-propagates through compositing, the style-modulated
    MLPs, the ray bender, the positional encodings and the ray geometry (what the reference gets by replaying its ~10^3-op graph)."""

    @staticmethod
    def forward(ctx, meta, origins, dirs, w2o, *flat):
        K = len(meta["descs"])
        styles, deforms = list(flat[:K]), list(flat[K:2 * K])
        n_t = K if meta.get("sample_t") is not None else 0       # fine pass: the explicit ray parameters are differentiable inputs
        meta = dict(meta)
        if n_t:
            meta["sample_t"] = [t.detach() for t in flat[2 * K:3 * K]]
        saved = [] if _save_forward(meta) else None
        res = _launch_forward(meta, meta["lead"], origins, dirs, w2o, styles, deforms, saved)
        ctx.saved_forward = saved[0] if saved else None
        ctx.tile_counts = saved[1] if saved and len(saved) > 1 else None
        ctx.meta = meta
        # the descs in ``meta`` hold RAW pointers into each model's packed blob: keep those tensors alive with the graph (a second
        # forward before this node's backward may repack -- train-mode BatchNorm bumps the running statistics every call)
        ctx.packed_keepalive = [m.packed_parameters() for m in meta["models"]] if meta.get("models") else []
        ctx.n_t = n_t
        ctx.save_for_backward(origins, dirs, w2o, *styles, *deforms, *(meta["sample_t"] if n_t else []))
        names = [f"object_{k}" for k in range(K)] + ["global"]
        outs = [res[n][key] for n in names for key in DIFF_KEYS]
        ctx.n_bent = 0
        if meta.get("bent_gradients"):
            # per-sample bent positions x + displacement(x), object space: differentiable outputs whose upstream gradient the backward
            # hands to the library (PeOutGrads.bent_positions) -- forward_expected_positions builds its weighted average on them
            for k in range(K):
                rot, tr = w2o[:, k, :, :3], w2o[:, k, :, 3]
                o_k = torch.einsum("iab,ib->ia", rot, origins) + tr
                d_k = torch.einsum("iab,irb->ira", rot, dirs)
                r = res[f"object_{k}"]
                outs.append(o_k[:, None, None, :] + d_k[:, :, None, :] * r["positions_t"].reshape(dirs.size(0), dirs.size(1), -1, 1)
                            + r["displacements"].reshape(dirs.size(0), dirs.size(1), -1, 3))
            ctx.n_bent = K
        extra = [res[n]["integrated_divergence"] for n in names]
        if meta.get("return_raw_alphas"):
            extra += [res[f"object_{k}"]["raw_alphas"] for k in range(K)]
        ctx.mark_non_differentiable(*extra)
        # outputs the loss never touches reach backward as None instead of zero tensors: no fills, and the compositing backward skips the
        # per-object lists of objects whose integrated outputs are unused (the usual training case: only "global" carries a gradient)
        ctx.set_materialize_grads(False)
        return tuple(outs + extra)

    @staticmethod
    def backward(ctx, *grads):
        meta = ctx.meta
        descs, models = meta["descs"], meta["models"]
        K = len(descs)
        saved = ctx.saved_tensors
        origins, dirs, w2o = saved[0], saved[1], saved[2]
        styles, deforms = list(saved[3:3 + K]), list(saved[3 + K:3 + 2 * K])
        n_t = ctx.n_t
        if n_t:
            meta = dict(meta)
            meta["sample_t"] = list(saved[3 + 2 * K:3 + 3 * K])
        device = dirs.device
        L = _cabi.lib()
        images, rays = dirs.size(0), dirs.size(1)
        scene = _scene_struct(meta, images, rays)
        keep: List = []
        ins = _inputs_struct(meta, meta["lead"], rays, origins, dirs, w2o, styles, deforms, keep)
        gout = _cabi.PeOutGrads()
        for i in range(K + 1):
            target = gout.object[i] if i < K else gout.global_
            for j, key in enumerate(DIFF_KEYS):
                g = grads[i * len(DIFF_KEYS) + j]
                if g is not None:
                    g = g.to(torch.float32).contiguous()
                    keep.append(g)
                    setattr(target, key, _cabi.ptr(g))
        for k in range(ctx.n_bent):
            g = grads[(K + 1) * len(DIFF_KEYS) + k]
            if g is not None:
                g = g.to(torch.float32).contiguous()
                keep.append(g)
                gout.bent_positions[k] = _cabi.ptr(g)
        need = ctx.needs_input_grad
        gin = _cabi.PeInGrads()
        # every input gradient is a 16-byte aligned piece of ONE zero-filled buffer (the kernels accumulate into them)
        wanted = [(origins, need[1]), (dirs, need[2]), (w2o, need[3])]
        wanted += [(styles[k], need[4 + k]) for k in range(K)] + [(deforms[k], need[4 + K + k]) for k in range(K)]
        wanted += [(meta["sample_t"][k], need[4 + 2 * K + k]) for k in range(n_t)]
        in_sizes = [((t.numel() + 3) // 4 * 4) if want else 0 for t, want in wanted]
        flat_in = torch.zeros(sum(in_sizes), dtype=torch.float32, device=device)
        pieces, off = [], 0
        for (t, want), sz in zip(wanted, in_sizes):
            pieces.append(flat_in[off:off + t.numel()].view(t.shape) if want else None)
            off += sz
        g_origins, g_dirs, g_w2o = pieces[0], pieces[1], pieces[2]
        g_styles, g_deforms, g_ts = pieces[3:3 + K], pieces[3 + K:3 + 2 * K], pieces[3 + 2 * K:]
        gin.ray_origins, gin.ray_directions, gin.w2o = _cabi.ptr(g_origins), _cabi.ptr(g_dirs), _cabi.ptr(g_w2o)
        for k in range(K):
            gin.style[k], gin.deformation[k] = _cabi.ptr(g_styles[k]), _cabi.ptr(g_deforms[k])
        for k in range(n_t):
            gin.sample_t[k] = _cabi.ptr(g_ts[k])
        params = (_cabi.PeObjectParams * _cabi.PE_MAX_OBJECTS)()
        g_params: List = []
        idx = 4 + 2 * K + n_t
        # one zero-filled buffer for every parameter gradient of the call (the kernels accumulate into it): one fill instead of one per
        # tensor (~60 per object model); 16-byte aligned pieces
        slots = [(k, field, i, tensor) for k, m in enumerate(models) for field, i, tensor in m.parameter_slots()]
        sizes = [((t.numel() + 3) // 4 * 4) if need[idx + j] else 0 for j, (_, _, _, t) in enumerate(slots)]
        flat_grads = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        offsets, off = [], 0
        for sz in sizes:
            offsets.append(off)
            off += sz
        slot_iter = iter(range(len(slots)))
        for k, m in enumerate(models):
            ps, kp = m.parameter_struct()
            keep.append(kp)
            params[k] = ps
            for field, i, tensor in m.parameter_slots():
                j = next(slot_iter)
                g = flat_grads[offsets[j]:offsets[j] + tensor.numel()].view(tensor.shape) if need[idx] else None
                idx += 1
                g_params.append(g)
                if g is not None:
                    if i is None:
                        setattr(gin.params[k], field, _cabi.ptr(g))
                    else:
                        getattr(gin.params[k], field)[i] = _cabi.ptr(g)
        with torch.cuda.device(device):
            fwd_ws = ctx.saved_forward
            scene.keep_samples = 1 if fwd_ws is not None else 0
            if fwd_ws is not None and ctx.tile_counts is not None:
                host, landed, _ = ctx.tile_counts
                landed.synchronize()
                for k in range(K):
                    scene.bwd_tiles[k] = int(host[k]) + 1
            nbytes = L.pe_backward_workspace_bytes(C.byref(scene))
            if nbytes == 0:
                raise _cabi.PeError(f"pe_backward_workspace_bytes: {L.pe_last_error().decode()}")
            ws = _Workspace.get(device, nbytes)
            if fwd_ws is not None:
                _cabi.check(L.pe_render_backward_saved(C.byref(scene), C.byref(ins), params, C.byref(gout), C.byref(gin), _cabi.ptr(fwd_ws),
                                                       fwd_ws.numel(), _cabi.ptr(ws), ws.numel(), _cabi.current_stream(device)))
            else:
                _cabi.check(L.pe_render_backward(C.byref(scene), C.byref(ins), params, C.byref(gout), C.byref(gin), _cabi.ptr(ws), ws.numel(),
                                                 _cabi.current_stream(device)))
        del keep
        return (None, g_origins, g_dirs, g_w2o, *g_styles, *g_deforms, *g_ts, *g_params)


def render_scene(descs: List[_cabi.PeObjectDesc], static_objects: int, ray_origins: torch.Tensor, ray_directions: torch.Tensor,
                 w2o: torch.Tensor, style: torch.Tensor, deformation: torch.Tensor, object_in_scene: torch.Tensor,
                 perturb: bool, training: bool, fix_object_overlaps: bool, apply_activation: bool, precision: int,
                 rand: Optional[List[torch.Tensor]] = None, noise: Optional[Dict[str, torch.Tensor]] = None,
                 bn_running: Optional[List] = None, return_raw_alphas: bool = False, models: Optional[List] = None,
                 return_samples: bool = False, peer_features: Optional[List[torch.Tensor]] = None,
                 sample_t: Optional[List[torch.Tensor]] = None, divergence_noise: Optional[List] = None,
                 bent_gradients: bool = False, handoff=None, global_only: bool = False) -> Dict:
    """One ObjectComposer.forward (reference: model/object_composer.py:786-892).  Returns {"object_k": {...}, "global": {...}}.
    With ``models`` (the object model of every instance) and autograd enabled the call is recorded as one RenderFunction node.
    ``sample_t`` (fine pass, :563-578): per object the explicit ray parameters (..., R, P_k) that replace the stratified samples."""
    device = ray_directions.device
    if device.type != "cuda":
        raise _cabi.PeError("ObjectComposer.forward needs CUDA tensors: the B200 render path has no CPU implementation")
    K = len(descs)
    if K > _cabi.PE_MAX_OBJECTS:
        raise _cabi.PeError(f"at most {_cabi.PE_MAX_OBJECTS} object instances per composer call")
    lead, images, rays, origins, dirs, m, styles, deforms, ois = _flatten_inputs(K, ray_origins, ray_directions, w2o, style, deformation,
                                                                                 object_in_scene)
    if perturb and (rand is None or noise is None):
        # torch's generator replaces the reference's torch.rand (ray_helper.py:1275) / torch.randn (object_composer.py:194)
        rand = [torch.rand(lead + [rays, d.positions], device=device) for d in descs] if sample_t is None else None
        noise = {f"object_{k}": torch.randn(lead + [rays, d.positions], device=device) for k, d in enumerate(descs)}
        noise["global"] = torch.randn(lead + [rays, sum(d.positions for d in descs)], device=device)
    meta = {"descs": descs, "static_objects": static_objects, "perturb": perturb, "training": training,
            "fix_object_overlaps": fix_object_overlaps, "apply_activation": apply_activation, "precision": precision,
            "rand": rand, "noise": noise, "ois": ois, "lead": lead, "bn_running": bn_running,
            "return_raw_alphas": return_raw_alphas, "models": models, "return_samples": return_samples, "peer_features": peer_features,
            "sample_t": sample_t, "divergence_noise": divergence_noise, "bent_gradients": bent_gradients, "handoff": handoff,
            "global_only": global_only}
    if divergence_noise is not None and models is None:
        raise _cabi.PeError("the Hutchinson divergence needs the object models (it is a training-time quantity)")
    if bent_gradients:
        meta["return_samples"] = True
    if peer_features and models is not None and torch.is_grad_enabled():
        raise _cabi.PeError("peer_features (fused all-gather of the feature grid) is an inference feature: call under torch.no_grad()")
    if handoff is not None and models is not None and torch.is_grad_enabled():
        raise _cabi.PeError("handoff (decoder grids written by the compositor) is an inference feature: call under torch.no_grad()")
    if models is not None and torch.is_grad_enabled():
        flat_params = [t for mdl in models for _, _, t in mdl.parameter_slots()]
        ts = [t.to(torch.float32).reshape(images, rays, -1) for t in sample_t] if sample_t is not None else []
        flat = RenderFunction.apply(meta, origins, dirs, m, *styles, *deforms, *ts, *flat_params)
        names = [f"object_{k}" for k in range(K)] + ["global"]
        results: Dict = {}
        it = iter(flat)
        for n in names:
            results[n] = {key: next(it) for key in DIFF_KEYS}
        if bent_gradients:
            for k in range(K):
                results[f"object_{k}"]["bent_positions"] = next(it).reshape(lead + [rays, descs[k].positions, 3])
        for n in names:
            results[n]["integrated_divergence"] = next(it)
        if return_raw_alphas:
            for k in range(K):
                results[f"object_{k}"]["raw_alphas"] = next(it)
        return results
    with torch.no_grad():
        return _launch_forward(meta, lead, origins, dirs, m, styles, deforms, alias_single=True)


def field_on_positions(desc: _cabi.PeObjectDesc, images: int, n: int, positions, origins, directions, style, deformation,
                       training: bool, device):
    """RayBendingStyleNerfModel.forward on explicit positions (reference: ray_bending_style_nerf_model.py:137-219)."""
    if device.type != "cuda":
        raise _cabi.PeError("the field operator needs CUDA tensors: there is no CPU implementation")
    L = _cabi.lib()
    scene = _cabi.PeScene()
    scene.images, scene.rays, scene.objects, scene.static_objects = images, n, 1, 1
    scene.training = int(bool(training))
    scene.precision, scene.explicit_positions = _cabi.PRECISION_FP32, 1
    scene.object[0] = desc
    ins = _cabi.PeInputs()
    ins.positions, ins.ray_origins, ins.ray_directions = _cabi.ptr(positions), _cabi.ptr(origins), _cabi.ptr(directions)
    ins.style[0] = _cabi.ptr(style)
    ins.deformation[0] = _cabi.ptr(deformation)
    feats = torch.empty((images, n, desc.features), dtype=torch.float32, device=device)
    alphas = torch.empty((images, n), dtype=torch.float32, device=device)
    disp = torch.empty((images, n, 3), dtype=torch.float32, device=device)
    outs = _cabi.PeOutputs()
    outs.raw_features[0], outs.raw_alphas[0], outs.displacements[0] = _cabi.ptr(feats), _cabi.ptr(alphas), _cabi.ptr(disp)
    with torch.cuda.device(device):
        nbytes = L.pe_workspace_bytes(C.byref(scene))
        ws = _Workspace.get(device, nbytes)
        _cabi.check(L.pe_render_forward(C.byref(scene), C.byref(ins), C.byref(outs), _cabi.ptr(ws), ws.numel(),
                                        _cabi.current_stream(device)))
    return feats, alphas, disp
