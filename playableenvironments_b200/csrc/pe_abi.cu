// extern "C" entry points of libpe_b200.so (see include/pe_b200.h) and the host-side orchestration of one
// ObjectComposer.forward call (model/object_composer.py:786-892).
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>
#include <mutex>

#include "pe_kernels.cuh"


static thread_local char g_error[512] = "";
static thread_local int64_t g_launches = 0;

void pe_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void pe_count_launch(int n) { g_launches += n; }

int pe_device_sm_count(int* out) {
    int dev = 0;
    PE_CUDA_CHECK(cudaGetDevice(&dev));
    PE_CUDA_CHECK(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return PE_OK;
}

extern "C" int pe_abi_version(void) { return PE_ABI_VERSION; }
extern "C" const char* pe_last_error(void) { return g_error; }
extern "C" int64_t pe_take_launch_count(void) { const int64_t n = g_launches; g_launches = 0; return n; }

static int validate_object(const PeObjectDesc& d) {
    if (d.nerf_kind != PE_NERF_ADAIN && d.nerf_kind != PE_NERF_SKYBOX_V3) { pe_set_error("unknown nerf model kind %d", d.nerf_kind); return PE_ERR_INVALID; }
    if (d.bender_kind != PE_BENDER_ZEROED && d.bender_kind != PE_BENDER_POSITIONAL) { pe_set_error("unknown ray bender kind %d", d.bender_kind); return PE_ERR_INVALID; }
    if (d.layers < 1 || d.layers > PE_MAX_LAYERS || d.skip < 0) { pe_set_error("backbone_layers_count must be in [1,%d]", PE_MAX_LAYERS); return PE_ERR_INVALID; }
    if (d.skip >= d.layers) { pe_set_error("Skip layer must refer to a valid backbone layer idx"); return PE_ERR_INVALID; }   // adain_style_nerf_model.py:32-33
    if (d.octaves < 0 || d.octaves > PE_MAX_OCTAVES || d.b_octaves > PE_MAX_OCTAVES) { pe_set_error("at most %d octaves", PE_MAX_OCTAVES); return PE_ERR_INVALID; }
    if (d.width < 8 || d.width % 8 || d.features < 1 || d.positions < 1 || d.positions > 0xffff) { pe_set_error("bad field shape"); return PE_ERR_INVALID; }
    if (d.bender_kind == PE_BENDER_POSITIONAL && (d.b_layers < 1 || d.b_layers > PE_MAX_LAYERS || d.b_width < 8 || d.b_width % 8)) { pe_set_error("bad ray bender shape"); return PE_ERR_INVALID; }
    return PE_OK;
}

extern "C" size_t pe_packed_bytes(const PeObjectDesc* desc) {
    if (!desc || validate_object(*desc) != PE_OK) return 0;
    return (size_t)pe_layout(*desc).total;
}

extern "C" int pe_pack_object(const PeObjectDesc* desc, const PeObjectParams* params, void* packed, pe_stream_t stream) {
    if (!desc || !params || !packed) { pe_set_error("null argument"); return PE_ERR_INVALID; }
    int rc = validate_object(*desc);
    if (rc != PE_OK) return rc;
    return pe_launch_pack(*desc, pe_layout(*desc), *params, packed, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
struct ObjWorkspace {
    float *t, *raw, *dispmag, *feat, *aff1, *aff2, *run1, *run2, *fold_v, *fold_s, *bent, *h7, *div, *div_gpos;
    uint8_t* flags;
    int32_t *tile_list, *tile_count;
    uint8_t* inbox;
    double* stats;
};

struct Workspace {
    ObjWorkspace obj[PE_MAX_OBJECTS];
    float* div_stash;          // activation stash of the divergence pass (pe_field_bwd_kernel, ray-bender-only mode), per block
    int64_t div_stash_floats;
    size_t bytes;
};

static bool object_has_divergence(const PeScene& s, int k) {
    return s.divergence && !s.explicit_positions && s.object[k].bender_kind == PE_BENDER_POSITIONAL && !s.object[k].canonical_pose;
}
static int backward_grid();

// Train mode (batch-statistics BatchNorm: three launches with the statistics phases of the tcgen05 kernel) runs on the tensor cores
// too; PE_TC_TRAIN=0 keeps it on the fp32 field kernel.
static bool tc_allowed(const PeScene& s) {
    if (s.precision == PE_PRECISION_FP32 || s.explicit_positions) return false;
    if (s.training) { const char* env = getenv("PE_TC_TRAIN"); if (env && atoi(env) == 0) return false; }
    return true;
}

// Set while pe_render_backward recomputes the forward: every per-sample tensor must land in the workspace (the compositing backward
// reads t / raw alpha / features / masks), so the self-contained and folded-head shortcuts are off.
static thread_local bool g_keep_samples = false;
// ... and whether that recompute may run the ray bender on the tensor cores (the forward it mirrors did: a performance mode)
static thread_local bool g_recompute_tc_bender = false;
struct KeepSamples {
    bool prev;
    explicit KeepSamples(bool on = true) : prev(g_keep_samples) { g_keep_samples = on; }
    ~KeepSamples() { g_keep_samples = prev; }
};

// Arithmetic of one object.  The mixed mode keeps objects with fewer than 64 samples per ray in the fp32-class mode: alpha = 1 - exp(-relu(raw) * delta)
// amplifies an absolute raw-alpha error by the sample spacing delta (court P = 4: delta ~ 20, Minecraft ground in front of the skybox: ~85;
// players P = 32: the fp16 rounding of the ACTIVATIONS alone leaves 2-3e-3 on the compositing weights, measured, profiles/r2_mixed_mode.md),
// and their share of a frame's FLOPs is small.
static int object_precision(const PeScene& s, int k) {
    if (s.precision == PE_PRECISION_MIXED) {
        const char* env = getenv("PE_TC_X3_BELOW");           // diagnostic: the samples-per-ray threshold below which objects run fp16x3
        if (s.object[k].positions < (env ? atoi(env) : 64)) return PE_PRECISION_FP16X3;
    }
    return s.precision;
}

// Two-pass layers of a `mixed` object (bit l = layer l, pe_field_tc.cu), -1: the kernel's default (PE_TC_MIXED_MASK).  With an
// activation-aware weight stream (PeObjectDesc.aware_rounding) and >= 96 samples per ray far fewer layers need the second pass
// (PE_TC_AWARE_MASK, measured in profiles/r2_aware_rounding.md; emulation: 7e-4 worst output at 128 samples per ray, 1.0e-3 at 64).
static int object_pass2_mask(const PeScene& s, int k) {
    const char* amin = getenv("PE_TC_AWARE_MIN_POSITIONS");  // diagnostic
    if (s.precision != PE_PRECISION_MIXED || !s.object[k].aware_rounding || s.object[k].positions < (amin ? atoi(amin) : 96) || s.training) return -1;
    const char* env = getenv("PE_TC_AWARE_MASK");
    return env ? (int)strtol(env, nullptr, 0) : PE_TC_AWARE_MASK;
}

static bool object_uses_tc(const PeScene& s, int k) { return tc_allowed(s) && pe_tc_shape_ok(s.object[k]); }
static bool object_uses_prepass(const PeScene& s, int k);
static bool prepass_object(const PeScene& s, int k) { return object_uses_prepass(s, k); }

// Objects with a positional ray bender: sampling + bender run as an exact fp32 pre-pass, the field runs on the tensor cores over
// the non-empty tiles only, the compositor integrates the object.  PE_TC_PREPASS=0 sends them to the fp32 field kernel instead.
static bool object_uses_prepass(const PeScene& s, int k) {
    const char* env = getenv("PE_TC_PREPASS");
    if (env && atoi(env) == 0) return false;
    return tc_allowed(s) && pe_tc_prepass_ok(s.object[k]);
}

// Static (zeroed-bender) objects of a multi-object scene take the same hand-off without a bender: sampling as its own pass
// (pe_sample_kernel), the field over the tiles that hold a sample inside the box only, the compositor integrates the object.  An object
// that covers part of the view (the shipped Tennis scene's upright slab behind the court, a Minecraft block) no longer pays for the
// rays that miss it -- the reference gathers the in-box samples before its MLP, ray_bending_style_nerf_model.py:170-176.  The backward
// is unchanged: it walks compacted in-box samples already.  PE_TC_SKIP_EMPTY=0: every tile, sampling inside the kernel, as before.
static bool object_lists_tiles(const PeScene& s, int k) {
    const char* env = getenv("PE_TC_SKIP_EMPTY");
    if (env && atoi(env) == 0) return false;
    return object_uses_tc(s, k) && s.objects > 1 && s.object[k].positions <= 128;
}
// objects whose integrated outputs come out of the fused field kernel itself (the compositor integrates the others)
static bool object_integrates_itself(const PeScene& s, int k) { return object_uses_tc(s, k) && !object_lists_tiles(s, k); }

// per-sample features are only materialised where something downstream reads them
static bool needs_feature_buffer(const PeScene& s, int k) {
    if (s.explicit_positions) return false;                 // caller supplies raw_features
    if (g_keep_samples) return true;                        // forward recomputed for the backward
    if (!object_uses_tc(s, k)) return true;                 // the fp32 path integrates in the compositor
    // tc path integrates its own object; the composition needs the samples when there are several objects, or when
    // perturb is on (the composed scene draws its own raw-alpha noise, object_composer.py:886 -> :194)
    return s.objects > 1 || s.perturb;
}

// Folded-head mode of the tcgen05 kernel (pe_tc_common.cuh): head layer 6 applied once per ray.  Possible when no
// per-sample feature leaves the kernel and a warp's 32 rows belong to one ray.  PE_TC_FOLD=0 disables it.
static bool object_folds_head(const PeScene& s, int k) {
    const char* env = getenv("PE_TC_FOLD");
    if (env && atoi(env) == 0) return false;
    return object_uses_tc(s, k) && !needs_feature_buffer(s, k) && !s.apply_activation && s.object[k].positions % 32 == 0;
}

static bool backward_on_tc(const PeScene& s, int k);

static Workspace carve(const PeScene& s, void* base) {
    Workspace w = {};
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off = (off + bytes + 255) / 256 * 256; return p; };
    for (int k = 0; k < s.objects; ++k) {
        const PeObjectDesc& d = s.object[k];
        const size_t P = s.explicit_positions ? 1 : d.positions;
        const size_t n = (size_t)s.images * s.rays * P;
        ObjWorkspace& o = w.obj[k];
        o.t = (float*)take(n * 4);
        o.raw = (float*)take(n * 4);
        o.dispmag = (float*)take(n * 4);
        o.inbox = (uint8_t*)take(n);
        o.feat = needs_feature_buffer(s, k) ? (float*)take(n * d.features * 4) : nullptr;
        const bool fold = object_folds_head(s, k);
        o.fold_v = fold ? (float*)take((size_t)s.images * s.rays * 128 * 4) : nullptr;
        o.fold_s = fold ? (float*)take((size_t)s.images * s.rays * 4) : nullptr;
        const bool prepass = object_uses_prepass(s, k) || object_lists_tiles(s, k);
        const size_t rpt = prepass ? 128 / P : 1, tiles = ((size_t)s.rays + rpt - 1) / rpt * (size_t)s.images;   // pre-pass objects have P <= 128
        o.bent = prepass ? (float*)take(n * 12) : nullptr;
        o.flags = prepass ? (uint8_t*)take(n) : nullptr;
        o.tile_list = prepass ? (int32_t*)take(tiles * 4) : nullptr;
        o.tile_count = prepass ? (int32_t*)take(4) : nullptr;
        // train mode on the tensor cores: the trunk output of every sample (fp32, 1 KB), written by the first BatchNorm phase and read by
        // the other two (and by the BatchNorm passes of the fp32 field backward) instead of re-evaluating encoding + trunk
        // (PE_TC_TRUNK_CACHE=0 disables)
        const char* cenv = getenv("PE_TC_TRUNK_CACHE");
        const bool cache = s.training && (object_uses_tc(s, k) || prepass) && !(cenv && atoi(cenv) == 0);
        o.h7 = cache ? (float*)take(n * d.width * 4) : nullptr;
        o.aff1 = (float*)take((size_t)s.images * 2 * d.width * 4);
        o.aff2 = (float*)take((size_t)s.images * d.width * 4);
        o.stats = (double*)take((size_t)(3 * d.width + 4) * 8);
        o.run1 = (float*)take((size_t)2 * d.width * 4);
        o.run2 = (float*)take((size_t)d.width * 4);
        const bool dv = object_has_divergence(s, k);
        o.div = dv ? (float*)take(n * 4) : nullptr;
        o.div_gpos = dv ? (float*)take(n * 12) : nullptr;
        if (dv) {
            const int64_t f = pe_field_bwd_stash_floats(d, pe_layout(d));
            w.div_stash_floats = f > w.div_stash_floats ? f : w.div_stash_floats;
        }
    }
    w.div_stash = w.div_stash_floats ? (float*)take((size_t)w.div_stash_floats * backward_grid() * 4) : nullptr;
    w.bytes = off;
    return w;
}

static int validate_scene(const PeScene& s) {
    if (s.objects < 1 || s.objects > PE_MAX_OBJECTS) { pe_set_error("objects must be in [1,%d]", PE_MAX_OBJECTS); return PE_ERR_INVALID; }
    if (s.images < 0 || s.rays < 0) { pe_set_error("negative sizes"); return PE_ERR_INVALID; }
    for (int k = 0; k < s.objects; ++k) {
        int rc = validate_object(s.object[k]);
        if (rc != PE_OK) return rc;
        if (!s.object[k].packed) { pe_set_error("object %d has no packed parameters", k); return PE_ERR_INVALID; }
        if (s.object[k].features != s.object[0].features) { pe_set_error("all objects must share output_features"); return PE_ERR_INVALID; }
    }
    return PE_OK;
}

// Streams the objects of a scene run on (pe_render_forward), one pool per device, created on first use and kept for the process.
struct ObjectStreams {
    std::mutex mutex;
    cudaStream_t stream[PE_MAX_OBJECTS];
    cudaEvent_t fork, join[PE_MAX_OBJECTS];
};
static ObjectStreams* object_streams() {
    static std::mutex table_mutex;
    static ObjectStreams* table[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(table_mutex);
    if (!table[dev]) {
        ObjectStreams* p = new ObjectStreams();
        bool ok = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming) == cudaSuccess;
        for (int k = 0; k < PE_MAX_OBJECTS && ok; ++k)
            ok = cudaStreamCreateWithFlags(&p->stream[k], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p->join[k], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); delete p; return nullptr; }      // (no pool: the caller's stream takes everything)
        table[dev] = p;
    }
    return table[dev];
}

extern "C" size_t pe_workspace_bytes(const PeScene* scene) {
    if (!scene || validate_scene(*scene) != PE_OK) return 0;
    if (scene->keep_samples) { KeepSamples keep; return carve(*scene, nullptr).bytes + 256; }
    return carve(*scene, nullptr).bytes + 256;
}

extern "C" int pe_render_forward(const PeScene* scene, const PeInputs* in, const PeOutputs* out, void* workspace,
                                 size_t workspace_bytes, pe_stream_t stream_) {
    if (!scene || !in || !out) { pe_set_error("null argument"); return PE_ERR_INVALID; }
    const PeScene& s = *scene;
    int rc = validate_scene(s);
    if (rc != PE_OK) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    if ((size_t)workspace % 256) { pe_set_error("workspace must be 256-byte aligned"); return PE_ERR_WORKSPACE; }
    // keep_samples: every per-sample tensor lands in the (caller-kept) workspace, like the backward's own forward recompute
    KeepSamples keep_guard(s.keep_samples != 0 || g_keep_samples);
    const Workspace ws = carve(s, workspace);
    if (ws.bytes > workspace_bytes) { pe_set_error("workspace too small: %zu < %zu", workspace_bytes, ws.bytes); return PE_ERR_WORKSPACE; }
    if (s.perturb && !s.explicit_positions) {
        for (int k = 0; k < s.objects; ++k)
            if (!in->rand[k] && !in->sample_t[k]) { pe_set_error("perturb needs rand[%d]", k); return PE_ERR_INVALID; }
    }
    for (int k = 0; k < s.objects; ++k)
        if ((s.explicit_t != 0) != (in->sample_t[k] != nullptr)) { pe_set_error("scene.explicit_t and sample_t[%d] disagree", k); return PE_ERR_INVALID; }
    if (s.explicit_t && s.explicit_positions) { pe_set_error("explicit_t and explicit_positions are exclusive"); return PE_ERR_INVALID; }
    if (s.images == 0 || s.rays == 0) return PE_OK;
    int sm_count = 148;
    rc = pe_device_sm_count(&sm_count);
    if (rc != PE_OK) return rc;

    if (out->peers < 0 || out->peers > PE_MAX_PEERS) { pe_set_error("peers must be in [0,%d]", PE_MAX_PEERS); return PE_ERR_INVALID; }
    bool peers_fused = false;
    // one object instance: style prologues, sampling / ray bender / tile lists, field kernel(s) -- everything before the compositor
    auto render_object = [&](int k, cudaStream_t stream) -> int {
        int rc = PE_OK;
        const PeObjectDesc& d = s.object[k];
        const PeLayout L = pe_layout(d);
        const ObjWorkspace& o = ws.obj[k];
        const unsigned char* blob = (const unsigned char*)d.packed;
        auto P32 = [&](int64_t off) { return (const float*)(blob + off); };
        const bool tc = object_uses_tc(s, k);

        PeFieldArgs fa = {};
        fa.ob = d; fa.L = L;
        fa.images = s.images; fa.rays = s.rays; fa.objects = s.objects; fa.k = k;
        fa.perturb = s.perturb; fa.explicit_positions = s.explicit_positions; fa.training = s.training;
        fa.apply_activation = s.apply_activation; fa.precision = object_precision(s, k); fa.pass2_mask = object_pass2_mask(s, k) < 0 ? 0 : (object_pass2_mask(s, k) | 0x10000);
        fa.origins = in->ray_origins; fa.dirs = in->ray_directions; fa.w2o = in->w2o;
        fa.deformation = in->deformation[k]; fa.rand = in->rand[k]; fa.t_in = in->sample_t[k]; fa.positions = in->positions; fa.ois = in->object_in_scene;
        fa.aff1 = o.aff1; fa.aff2 = o.aff2;
        fa.t_out = out->positions_t[k] ? out->positions_t[k] : o.t;
        fa.raw_out = out->raw_alphas[k] ? out->raw_alphas[k] : o.raw;
        const bool self_contained = tc && s.objects == 1 && !s.perturb && !g_keep_samples;     // the fused kernel integrates the whole scene itself
        fa.feat_out = out->raw_features[k] ? out->raw_features[k] : o.feat;
        fa.disp_out = out->displacements[k];
        fa.dispmag_out = o.dispmag;
        fa.inbox_out = o.inbox;
        if (self_contained) {       // nothing downstream reads per-sample tensors: do not write them
            fa.t_out = out->positions_t[k]; fa.raw_out = out->raw_alphas[k]; fa.dispmag_out = nullptr; fa.inbox_out = nullptr;
        }
        fa.stats = o.stats;
        fa.integ = out->object[k];
        fa.noise = s.perturb ? in->noise[k] : nullptr;
        if (tc && !fa.feat_out && o.fold_v) { fa.fold_v = o.fold_v; fa.fold_s = o.fold_s; }
        if (fa.fold_v && s.objects == 1 && !s.training && !prepass_object(s, k)) {      // the fused kernel produces the scene's grid itself
            fa.peers = out->peers;
            for (int q = 0; q < out->peers; ++q) fa.peer_features[q] = out->peer_features[q];
            peers_fused = true;
        }
        fa.h7_out = o.h7;
        if (!tc && !fa.feat_out) { pe_set_error("internal: no feature buffer for object %d", k); return PE_ERR_INVALID; }
        if (d.bender_kind == PE_BENDER_POSITIONAL && !fa.deformation) { pe_set_error("object %d needs a deformation code", k); return PE_ERR_INVALID; }
        if (!in->style[k]) { pe_set_error("object %d needs a style code", k); return PE_ERR_INVALID; }

        PeStyleArgs s1 = {};
        s1.images = s.images; s1.style_features = d.style_features; s1.channels = d.width; s1.training = 0;
        s1.style = in->style[k]; s1.aff_w = P32(L.aff1_w); s1.aff_b = P32(L.aff1_b);
        s1.run_mean = P32(L.bn1_mean); s1.run_var = P32(L.bn1_var); s1.stats = o.stats; s1.out = o.aff1; s1.running_out = o.run1;
        PeStyleArgs s2 = s1;
        s2.channels = d.width / 2; s2.aff_w = P32(L.aff2_w); s2.aff_b = P32(L.aff2_b);
        s2.run_mean = P32(L.bn2_mean); s2.run_var = P32(L.bn2_var); s2.stats = o.stats + 2 * d.width + 2; s2.out = o.aff2; s2.running_out = o.run2;

        const bool prepass = object_uses_prepass(s, k);
        const bool lists = object_lists_tiles(s, k);
        auto launch_field = [&](int phase) {
            fa.phase = phase;
            const PeIntegrated none = {};
            if (lists) {
                PeFieldArgs pre = fa;
                pre.bent = o.bent; pre.flags = o.flags; pre.integ = none;
                pre.tile_list = o.tile_list; pre.tile_count = o.tile_count;
                if (!(s.training && phase != 1)) {       // (train mode: the samples and the tile list of phase 1 stay valid for phases 2 and 0)
                    int rc2 = pe_launch_sample(pre, sm_count, stream); if (rc2) return rc2;
                    rc2 = pe_launch_tile_list(pre, 2, o.tile_list, o.tile_count, stream); if (rc2) return rc2;
                }
                return pe_launch_field_tc(pre, none, sm_count, stream);
            }
            if (prepass) {
                PeFieldArgs pre = fa;
                pre.bent = o.bent; pre.flags = o.flags; pre.integ = none;
                pre.tile_list = o.tile_list; pre.tile_count = o.tile_count;
                int rc2;
                if (s.training && phase != 1) {
                    // statistics phase 2 / final pass of a train-mode call: positions, masks and the tile list of phase 1 are still valid
                    PeFieldArgs tcargs = pre;
                    tcargs.phase = phase;
                    return pe_launch_field_tc(tcargs, none, sm_count, stream);
                }
                // The bender's output feeds 2^9-octave Fourier features (x2pi/size: an absolute error e of the normalised
                // displacement becomes a phase error of 3217 e), so the tensor-core bender (hi/lo split, all four partial products,
                // but the tensor core's own fp32 accumulation) lands at ~3e-4 of the reference on the rendered outputs: used in the
                // performance modes (mixed, fp16, fp16x2); the parity-first mode (fp16x3) keeps the exact fp32 bender.  PE_TC_BENDER=0/1 forces.
                const char* benv = getenv("PE_TC_BENDER");
                const bool tc_bender = benv ? atoi(benv) != 0 : ((g_keep_samples && !s.keep_samples) ? g_recompute_tc_bender : s.precision != PE_PRECISION_FP16X3);
                if (pe_tc_bender_ok(d) && tc_bender) {
                    // 1a. exact fp32 sampling: t, positions, outer mask; empty-space values everywhere
                    rc2 = pe_launch_sample(pre, sm_count, stream); if (rc2) return rc2;
                    // 1b. the ray bender on the tensor cores (fp16x3: fp32-class) over the tiles with samples inside the box
                    rc2 = pe_launch_tile_list(pre, 1, o.tile_list, o.tile_count, stream); if (rc2) return rc2;
                    rc2 = pe_launch_bender_tc(pre, sm_count, stream); if (rc2) return rc2;
                } else {
                    // 1. exact fp32 sampling + ray bender: t, bent positions, masks, displacements; empty-space values everywhere
                    pre.phase = PE_PHASE_PREPASS;
                    rc2 = pe_launch_field_fp32(pre, sm_count, stream); if (rc2) return rc2;
                }
                // 2. which tiles hold a sample to evaluate
                rc2 = pe_launch_tile_list(pre, 2, o.tile_list, o.tile_count, stream); if (rc2) return rc2;
                // 3. the field on the tensor cores over those tiles (the compositor integrates the object)
                PeFieldArgs tcargs = pre;
                tcargs.phase = phase;
                return pe_launch_field_tc(tcargs, none, sm_count, stream);
            }
            if (!tc) return pe_launch_field_fp32(fa, sm_count, stream);
            const PeIntegrated& gout = (s.objects == 1 && !s.perturb) ? out->global : none;
            return pe_launch_field_tc(fa, gout, sm_count, stream);
        };
        if (s.training) {
            // train-mode BatchNorm (adain.py:47): statistics over all in-box samples of this object in this call.
            // Three passes, two global reductions.
            PE_CUDA_CHECK(cudaMemsetAsync(o.stats, 0, (size_t)(3 * d.width + 4) * 8, stream));
            rc = pe_launch_style(s1, stream); if (rc) return rc;      // any valid affine for the unused epilogue
            rc = pe_launch_style(s2, stream); if (rc) return rc;
            rc = launch_field(1); if (rc) return rc;
            s1.training = 1;
            rc = pe_launch_style(s1, stream); if (rc) return rc;
            rc = launch_field(2); if (rc) return rc;
            s2.training = 1;
            rc = pe_launch_style(s2, stream); if (rc) return rc;
            rc = launch_field(0); if (rc) return rc;
            if (out->bn1_running[k]) PE_CUDA_CHECK(cudaMemcpyAsync(out->bn1_running[k], o.run1, (size_t)2 * d.width * 4, cudaMemcpyDeviceToDevice, stream));
            if (out->bn2_running[k]) PE_CUDA_CHECK(cudaMemcpyAsync(out->bn2_running[k], o.run2, (size_t)d.width * 4, cudaMemcpyDeviceToDevice, stream));
        } else {
            rc = pe_launch_style(s1, stream); if (rc) return rc;
            rc = pe_launch_style(s2, stream); if (rc) return rc;
            rc = launch_field(0); if (rc) return rc;
        }
        if (object_has_divergence(s, k) && in->divergence_noise[k]) {
            // Hutchinson divergence of the displacement field (object_composer.py:582-601): e . (J e) per sample by ONE vector-Jacobian
            // product through the ray bender -- the ray-bender-only mode of the fp32 field backward with dL/d displacement = e and no
            // parameter gradients (what torch.autograd.grad(displacements, positions, e) is to the reference)
            if (!in->divergence_params) { pe_set_error("divergence_noise needs divergence_params"); return PE_ERR_INVALID; }
            PeFieldBwdArgs dvb = {};
            dvb.f = fa;
            dvb.f.integ = PeIntegrated{};
            dvb.w = in->divergence_params[k];
            dvb.g_bent_in = in->divergence_noise[k]; dvb.g_bent_flag = 1;
            dvb.g_pos = o.div_gpos; dvb.div_out = o.div;
            dvb.stash = ws.div_stash; dvb.stash_floats = ws.div_stash_floats;
            dvb.bwd_phase = 0;
            rc = pe_launch_field_bwd(dvb, sm_count, stream); if (rc) return rc;
        }
        return PE_OK;
    };
    // The objects of a scene are independent until the compositor, and in the frames the callers render (play.py: 11 520 rays, one
    // train.py replica: 20 480) an object's launches are a few tiles per SM or less: each object runs on its own stream of a per-device
    // pool (fork from the caller's stream, join before the compositor), so one object's kernels fill the SMs another's leave idle.
    // Stream capture sees an ordinary fork / join.  PE_CONCURRENT_OBJECTS=0: everything on the caller's stream.
    ObjectStreams* pool = nullptr;
    {
        const char* cenv = getenv("PE_CONCURRENT_OBJECTS");
        const bool concurrent = s.objects > 1 && !s.explicit_positions && !s.divergence && !(cenv && atoi(cenv) == 0);
        if (concurrent) pool = object_streams();
    }
    if (pool) {
        std::lock_guard<std::mutex> lock(pool->mutex);      // the pool's events are re-recorded by every call: one call enqueues at a time
        PE_CUDA_CHECK(cudaEventRecord(pool->fork, stream));
        int first_rc = PE_OK;
        int forked = 0;
        for (int k = 0; k < s.objects && first_rc == PE_OK; ++k, ++forked) {
            cudaError_t e = cudaStreamWaitEvent(pool->stream[k], pool->fork, 0);
            if (e == cudaSuccess) { first_rc = render_object(k, pool->stream[k]); e = cudaEventRecord(pool->join[k], pool->stream[k]); }
            if (e != cudaSuccess && first_rc == PE_OK) { pe_set_error("CUDA error: %s", cudaGetErrorString(e)); first_rc = PE_ERR_CUDA; }
        }
        for (int k = 0; k < forked; ++k) cudaStreamWaitEvent(stream, pool->join[k], 0);      // always join what was forked (stream capture)
        if (first_rc != PE_OK) return first_rc;
    } else {
        for (int k = 0; k < s.objects; ++k) {
            rc = render_object(k, stream);
            if (rc != PE_OK) return rc;
        }
    }
    if (s.explicit_positions) return PE_OK;

    PeCompositeArgs ca = {};
    ca.images = s.images; ca.rays = s.rays; ca.objects = s.objects; ca.static_objects = s.static_objects;
    ca.features = s.object[0].features; ca.fix_overlaps = s.fix_object_overlaps; ca.perturb = s.perturb;
    ca.dirs = in->ray_directions;
    ca.noise_global = s.perturb ? in->noise_global : nullptr;
    bool all_tc = true;
    for (int k = 0; k < s.objects; ++k) {
        const ObjWorkspace& o = ws.obj[k];
        ca.positions[k] = s.object[k].positions;
        ca.total_positions += s.object[k].positions;
        ca.t[k] = out->positions_t[k] ? out->positions_t[k] : o.t;
        ca.raw[k] = out->raw_alphas[k] ? out->raw_alphas[k] : o.raw;
        ca.feat[k] = out->raw_features[k] ? out->raw_features[k] : o.feat;
        ca.dispmag[k] = o.dispmag;
        ca.div[k] = (object_has_divergence(s, k) && in->divergence_noise[k]) ? o.div : nullptr;
        ca.inbox[k] = o.inbox;
        ca.noise[k] = s.perturb ? in->noise[k] : nullptr;
        ca.object[k] = out->object[k];
        all_tc = all_tc && object_integrates_itself(s, k);
    }
    ca.global = out->global;
    if (out->handoff.segments) {
        if (s.objects < 2 || s.perturb) { pe_set_error("the decoder hand-off is written by the compositor: multi-object scenes without perturbation"); return PE_ERR_UNSUPPORTED; }
        ca.handoff = out->handoff;
    }
    // objects evaluated by the tcgen05 kernel integrate themselves in its epilogue; a single such object IS the scene
    ca.do_objects = all_tc ? 0 : 1;
    ca.do_global = (all_tc && s.objects == 1 && !s.perturb) ? 0 : 1;
    if (all_tc && s.objects > 1) ca.do_objects = 0;
    if (!all_tc) {
        // mixed scenes: the compositor integrates the fp32 objects; tc objects already wrote theirs
        for (int k = 0; k < s.objects; ++k)
            if (object_integrates_itself(s, k)) memset(&ca.object[k], 0, sizeof(PeIntegrated));
    }
    auto wants = [](const PeIntegrated& o) {
        return o.integrated_features || o.opacity || o.weights || o.depth || o.disparity || o.integrated_displacements_magnitude || o.integrated_divergence;
    };
    if (!wants(ca.global) && !ca.handoff.segments) ca.do_global = 0;      // (inference, single object: the caller aliases the scene's outputs to the object's)
    bool any_out = wants(ca.global) || ca.handoff.segments != 0;
    for (int k = 0; k < s.objects; ++k) any_out = any_out || wants(ca.object[k]);
    if ((ca.do_objects || ca.do_global) && any_out) {
        rc = pe_launch_composite(ca, stream);
        if (rc) return rc;
    }
    if (out->peers > 0 && !peers_fused) {
        // paths whose grid comes out of the compositor (or of the general field kernel): one copy per destination after the last kernel
        const float* src = out->global.integrated_features ? out->global.integrated_features : (s.objects == 1 ? out->object[0].integrated_features : nullptr);
        if (!src) { pe_set_error("peer_features needs the composed scene's integrated_features as an output"); return PE_ERR_INVALID; }
        const size_t bytes = (size_t)s.images * s.rays * s.object[0].features * sizeof(float);
        for (int q = 0; q < out->peers; ++q) PE_CUDA_CHECK(cudaMemcpyAsync(out->peer_features[q], src, bytes, cudaMemcpyDefault, stream));
    }
    return PE_OK;
}


// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct ObjBwdWorkspace {
    float *cw_obj, *cw_glob, *g_raw, *g_t, *g_dm, *g_pos, *g_od, *adain_sums, *bn_fix;
    float* g_pos2;                                     // scene.bent_gradients: dL/d position of the extra pass
    float *g_bent, *scale;                             // tensor-core field backward: dL/d bent position, gradient scale [S, 1/S] + scratch
    double* bn_sums;
    int32_t *slot_list, *slot_count, *tile_begin;      // compacted tiles of the field backward (NULL: dense tiles)
};

// The field backward walks tiles of 32 samples; with the in-box masks of the (tensor-core) forward recompute at hand it walks only the
// samples inside the object's box, compacted per image (PE_BWD_COMPACT=0: tiles of 32 consecutive slots as in the all-fp32 path).
static bool backward_compacts(const PeScene& s, int k) {
    const char* env = getenv("PE_BWD_COMPACT");
    if (env && atoi(env) == 0) return false;
    return s.object[k].nerf_kind == PE_NERF_ADAIN && (object_uses_tc(s, k) || object_uses_prepass(s, k));
}

// The field backward of the shipped field shape runs on the tensor cores (pe_bwd_tc.cu) over tiles of 128 compacted samples; PE_BWD_TC=0
// keeps it on the exact fp32 kernel.  Its activation / gradient stash holds `capacity` tiles (1.45 MB each); more tiles than that are
// processed in batches (the number of in-box samples is only known on the device, so the batch count is the worst case and surplus
// launches find no tile).
static bool backward_on_tc(const PeScene& s, int k) {
    if (s.explicit_t) return false;          // fine pass: exact fp32 backward (see PeScene.explicit_t)
    return backward_compacts(s, k) && !s.apply_activation && pe_bwd_tc_object_ok(s.object[k]) && pe_layout(s.object[k]).tcT_base != 0;
}
static int64_t bwd_tc_tiles_upper_bound(const PeScene& s, int k) {
    const int64_t worst = (int64_t)s.images * (((int64_t)s.rays * s.object[k].positions + PE_BWD_TILE - 1) / PE_BWD_TILE);
    // the caller knows the exact count from the kept forward (pe_forward_tile_counts): no surplus batches, a stash of the real size
    if (s.bwd_tiles[k] > 0) return pe_min64(worst, s.bwd_tiles[k] - 1 > 1 ? s.bwd_tiles[k] - 1 : 1);
    return worst;
}
static int64_t bwd_tc_capacity(const PeScene& s) {
    int64_t ub = 0;
    for (int k = 0; k < s.objects; ++k)
        if (backward_on_tc(s, k)) ub = bwd_tc_tiles_upper_bound(s, k) > ub ? bwd_tc_tiles_upper_bound(s, k) : ub;
    const char* env = getenv("PE_BWD_TC_MAX_TILES");
    const int64_t cap = env ? atoll(env) : 12288;
    return pe_min64(ub, cap > 0 ? cap : 1);
}

struct BwdWorkspace {
    unsigned char* tc_stash;
    int64_t tc_capacity;
    void* fwd;
    size_t fwd_bytes;
    ObjBwdWorkspace obj[PE_MAX_OBJECTS];
    void* zero_begin;          // region cleared at the start of every call (sums)
    size_t zero_bytes;
    float* stash;
    int64_t stash_floats;      // per block
    size_t bytes;
};

static PeScene backward_scene(const PeScene& s) {
    PeScene b = s;
    // The backward needs the forward's per-sample tensors, AdaIn constants and BatchNorm sums again: recomputed in the fp32-class
    // tensor-core mode where the shapes allow (exact fp32 kernels otherwise, or with PE_TC_BACKWARD_RECOMPUTE=0); the field backward
    // itself (pe_field_bwd_kernel) recomputes and differentiates each tile in exact fp32.
    const char* env = getenv("PE_TC_BACKWARD_RECOMPUTE");
    b.precision = (env && atoi(env) == 0) ? PE_PRECISION_FP32 : PE_PRECISION_FP16X3;
    return b;
}

static BwdWorkspace carve_backward(const PeScene& s, void* base, int grid) {
    BwdWorkspace w = {};
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off = (off + bytes + 255) / 256 * 256; return p; };
    KeepSamples keep;
    w.fwd_bytes = s.keep_samples ? 0 : carve(s, nullptr).bytes;       // (a saved forward lives in the caller's buffer)
    w.fwd = take(w.fwd_bytes);
    for (int k = 0; k < s.objects; ++k) {
        const PeObjectDesc& d = s.object[k];
        const size_t n = (size_t)s.images * s.rays * d.positions;
        ObjBwdWorkspace& o = w.obj[k];
        o.cw_obj = (float*)take(n * 4); o.cw_glob = (float*)take(n * 4);
        o.g_raw = (float*)take(n * 4); o.g_t = (float*)take(n * 4); o.g_dm = (float*)take(n * 4);
        o.g_pos = (float*)take(n * 12);
        o.g_od = d.nerf_kind == PE_NERF_SKYBOX_V3 ? (float*)take(n * 24) : nullptr;
        o.g_pos2 = (s.bent_gradients && d.bender_kind == PE_BENDER_POSITIONAL) ? (float*)take(n * 12) : nullptr;
        const bool compact = backward_compacts(s, k);
        o.slot_list = compact ? (int32_t*)take(n * 4) : nullptr;
        o.slot_count = compact ? (int32_t*)take((size_t)s.images * 4) : nullptr;
        o.tile_begin = compact ? (int32_t*)take(((size_t)s.images + 1) * 4) : nullptr;
        const bool tcb = backward_on_tc(s, k);
        o.g_bent = (tcb && object_uses_prepass(s, k)) ? (float*)take(n * 12) : nullptr;
        o.scale = tcb ? (float*)take(64) : nullptr;
    }
    w.tc_capacity = bwd_tc_capacity(s);
    w.tc_stash = w.tc_capacity ? (unsigned char*)take(pe_bwd_tc_stash_bytes(w.tc_capacity)) : nullptr;
    const size_t z0 = off;
    w.zero_begin = base ? (char*)base + off : nullptr;
    for (int k = 0; k < s.objects; ++k) {
        const PeObjectDesc& d = s.object[k];
        ObjBwdWorkspace& o = w.obj[k];
        o.adain_sums = (float*)take((size_t)s.images * 3 * d.width * 4);
        o.bn_sums = (double*)take((size_t)3 * d.width * 8);
        o.bn_fix = (float*)take((size_t)3 * d.width * 4);
    }
    w.zero_bytes = off - z0;
    int64_t stash = 0;
    for (int k = 0; k < s.objects; ++k) {
        const int64_t f = pe_field_bwd_stash_floats(s.object[k], pe_layout(s.object[k]));
        stash = f > stash ? f : stash;
    }
    w.stash_floats = stash;
    w.stash = (float*)take((size_t)stash * 4 * grid);
    w.bytes = off;
    return w;
}

static int backward_grid() {
    int sm = 0;
    if (pe_device_sm_count(&sm) != PE_OK || sm <= 0) sm = 160;     // sizing without a device: upper bound
    return pe_field_bwd_grid(sm);
}

// With a saved forward (scene.keep_samples) the backward takes its path decisions from the scene itself -- the saved workspace was carved
// for it -- instead of the recompute's fp32-class scene.
static PeScene backward_scene_for(const PeScene& scene) { return scene.keep_samples ? scene : backward_scene(scene); }

extern "C" size_t pe_backward_workspace_bytes(const PeScene* scene) {
    if (!scene || validate_scene(*scene) != PE_OK) return 0;
    if (scene->explicit_positions) { pe_set_error("backward on explicit positions is not supported"); return 0; }
    return carve_backward(backward_scene_for(*scene), nullptr, backward_grid()).bytes + 256;
}

extern "C" int pe_forward_tile_counts(const PeScene* scene, const void* saved_forward, size_t saved_forward_bytes, int64_t* counts,
                                      pe_stream_t stream_) {
    if (!scene || !saved_forward || !counts) { pe_set_error("null argument"); return PE_ERR_INVALID; }
    int rc = validate_scene(*scene);
    if (rc != PE_OK) return rc;
    if (!scene->keep_samples) { pe_set_error("pe_forward_tile_counts reads the workspace of a forward with scene.keep_samples = 1"); return PE_ERR_INVALID; }
    const PeScene& s = *scene;
    if ((size_t)saved_forward % 256 || carve(s, nullptr).bytes > saved_forward_bytes) { pe_set_error("saved forward workspace: misaligned or too small"); return PE_ERR_WORKSPACE; }
    cudaStream_t stream = (cudaStream_t)stream_;
    PE_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * PE_MAX_OBJECTS, stream));
    if (s.images == 0 || s.rays == 0) return PE_OK;
    const Workspace ws = carve(s, const_cast<void*>(saved_forward));
    for (int k = 0; k < s.objects; ++k) {
        if (!backward_on_tc(s, k)) continue;
        // the very masks pe_render_backward_saved compacts (outer mask of a ray-bender object, in-box flags otherwise)
        const uint8_t* mask_src = object_uses_prepass(s, k) ? ws.obj[k].flags : ws.obj[k].inbox;
        rc = pe_launch_count_tiles(mask_src, 1, s.images, (int64_t)s.rays * s.object[k].positions, PE_BWD_TILE, counts + k, stream);
        if (rc) return rc;
    }
    return PE_OK;
}

extern "C" int pe_render_backward(const PeScene* scene, const PeInputs* in, const PeObjectParams* params, const PeOutGrads* grad_out,
                                  const PeInGrads* grad_in, void* workspace, size_t workspace_bytes, pe_stream_t stream_) {
    if (scene && scene->keep_samples) { pe_set_error("scene.keep_samples is set: call pe_render_backward_saved with the kept forward workspace"); return PE_ERR_INVALID; }
    return pe_render_backward_saved(scene, in, params, grad_out, grad_in, nullptr, 0, workspace, workspace_bytes, stream_);
}

extern "C" int pe_render_backward_saved(const PeScene* scene, const PeInputs* in, const PeObjectParams* params, const PeOutGrads* grad_out,
                                        const PeInGrads* grad_in, const void* saved_forward, size_t saved_forward_bytes, void* workspace,
                                        size_t workspace_bytes, pe_stream_t stream_) {
    if (!scene || !in || !params || !grad_out || !grad_in) { pe_set_error("null argument"); return PE_ERR_INVALID; }
    int rc = validate_scene(*scene);
    if (rc != PE_OK) return rc;
    if (scene->explicit_positions) { pe_set_error("backward on explicit positions is not supported"); return PE_ERR_UNSUPPORTED; }
    if ((saved_forward != nullptr) != (scene->keep_samples != 0)) { pe_set_error("a saved forward workspace goes with scene.keep_samples = 1, and only with it"); return PE_ERR_INVALID; }
    if (saved_forward && scene->precision == PE_PRECISION_FP32) { pe_set_error("the saved-forward backward needs a tensor-core precision mode (compacted sample lists come from its masks)"); return PE_ERR_UNSUPPORTED; }
    const PeScene s = backward_scene_for(*scene);
    cudaStream_t stream = (cudaStream_t)stream_;
    if ((size_t)workspace % 256) { pe_set_error("workspace must be 256-byte aligned"); return PE_ERR_WORKSPACE; }
    int sm_count = 148;
    rc = pe_device_sm_count(&sm_count);
    if (rc != PE_OK) return rc;
    const int grid = pe_field_bwd_grid(sm_count);
    const BwdWorkspace bw = carve_backward(s, workspace, grid);
    if (bw.bytes > workspace_bytes) { pe_set_error("backward workspace too small: %zu < %zu", workspace_bytes, bw.bytes); return PE_ERR_WORKSPACE; }
    if (s.images == 0 || s.rays == 0) return PE_OK;

    // 1. the forward's per-sample t / raw alpha / features / |displacement| / in-box flags, AdaIn scale-shift, BatchNorm sums: kept by the
    //    caller's forward, or recomputed here
    const PeOutputs none = {};
    KeepSamples keep;
    if (saved_forward) {
        if ((size_t)saved_forward % 256 || carve(s, nullptr).bytes > saved_forward_bytes) { pe_set_error("saved forward workspace: misaligned or too small"); return PE_ERR_WORKSPACE; }
    } else {
        g_recompute_tc_bender = scene->precision == PE_PRECISION_MIXED || scene->precision == PE_PRECISION_FP16 || scene->precision == PE_PRECISION_FP16X2;
        rc = pe_render_forward(&s, in, &none, bw.fwd, bw.fwd_bytes, stream_);
        if (rc != PE_OK) return rc;
    }
    const Workspace ws = carve(s, saved_forward ? const_cast<void*>(saved_forward) : bw.fwd);
    PE_CUDA_CHECK(cudaMemsetAsync(bw.zero_begin, 0, bw.zero_bytes, stream));

    // 2. compositing backward
    PeCompositeBwdArgs cb = {};
    PeCompositeArgs& ca = cb.f;
    ca.images = s.images; ca.rays = s.rays; ca.objects = s.objects; ca.static_objects = s.static_objects;
    ca.features = s.object[0].features; ca.fix_overlaps = s.fix_object_overlaps; ca.perturb = s.perturb;
    ca.dirs = in->ray_directions;
    ca.noise_global = s.perturb ? in->noise_global : nullptr;
    for (int k = 0; k < s.objects; ++k) {
        const ObjWorkspace& o = ws.obj[k];
        ca.positions[k] = s.object[k].positions;
        ca.total_positions += s.object[k].positions;
        ca.t[k] = o.t; ca.raw[k] = o.raw; ca.feat[k] = o.feat; ca.dispmag[k] = o.dispmag; ca.inbox[k] = o.inbox;
        ca.noise[k] = s.perturb ? in->noise[k] : nullptr;
        cb.g_object[k] = grad_out->object[k];
        cb.cw_obj[k] = bw.obj[k].cw_obj; cb.cw_glob[k] = bw.obj[k].cw_glob;
        cb.g_raw[k] = bw.obj[k].g_raw; cb.g_t[k] = bw.obj[k].g_t; cb.g_dm[k] = bw.obj[k].g_dm;
    }
    cb.g_global = grad_out->global;
    cb.g_dirs = grad_in->ray_directions;
    rc = pe_launch_composite_bwd(cb, stream);
    if (rc != PE_OK) return rc;

    // 3. per object: field backward (three passes in train mode: the BatchNorm backward needs two global reductions), style, geometry
    for (int k = 0; k < s.objects; ++k) {
        const PeObjectDesc& d = s.object[k];
        const PeLayout L = pe_layout(d);
        const ObjWorkspace& o = ws.obj[k];
        const ObjBwdWorkspace& b = bw.obj[k];
        const int W = d.width;
        PeFieldBwdArgs fb = {};
        PeFieldArgs& fa = fb.f;
        fa.ob = d; fa.L = L;
        fa.images = s.images; fa.rays = s.rays; fa.objects = s.objects; fa.k = k;
        fa.perturb = s.perturb; fa.training = s.training; fa.apply_activation = s.apply_activation; fa.precision = s.precision;
        fa.origins = in->ray_origins; fa.dirs = in->ray_directions; fa.w2o = in->w2o;
        fa.deformation = in->deformation[k]; fa.rand = in->rand[k]; fa.t_in = in->sample_t[k]; fa.ois = in->object_in_scene;
        fa.aff1 = o.aff1; fa.aff2 = o.aff2;
        fb.w = params[k];
        fb.gw = grad_in->params[k];
        fb.cw_obj = b.cw_obj; fb.cw_glob = b.cw_glob;
        fb.g_feat_obj = grad_out->object[k].integrated_features;
        fb.g_feat_glob = grad_out->global.integrated_features;
        fb.g_raw = b.g_raw; fb.g_dm = b.g_dm;
        fb.g_pos = b.g_pos; fb.g_od = b.g_od;
        fb.adain_sums = b.adain_sums; fb.bn_sums = b.bn_sums; fb.bn_fix = b.bn_fix;
        fb.g_deformation = grad_in->deformation[k];
        fb.stash = bw.stash; fb.stash_floats = bw.stash_floats;
        fb.h7_cache = o.h7; fb.inbox_in = o.h7 ? o.inbox : nullptr;
        const bool on_tc = bw.tc_capacity > 0 && backward_on_tc(s, k);
        const bool prepass = object_uses_prepass(s, k);
        if (b.slot_list) {
            // samples inside the box (outer mask: the ray bender's backward covers them all), from the forward recompute's masks
            const uint8_t* mask_src = prepass ? o.flags : o.inbox;
            rc = pe_launch_compact_slots(mask_src, 1, s.images, (int64_t)s.rays * d.positions, b.slot_list, b.slot_count, b.tile_begin, stream,
                                         on_tc ? PE_BWD_TILE : 32);
            if (rc) return rc;
            PE_CUDA_CHECK(cudaMemsetAsync(b.g_pos, 0, (size_t)s.images * s.rays * d.positions * 12, stream));     // slots outside the list
            fb.slot_list = b.slot_list; fb.slot_count = b.slot_count; fb.tile_begin = b.tile_begin;
        }
        if (!fb.w.head0_w || !fb.w.head3_w || !fb.w.head6_w) { pe_set_error("backward needs the fp32 parameters of object %d", k); return PE_ERR_INVALID; }
        if (on_tc) {
            // ---- field backward on the tensor cores (pe_bwd_tc.cu): recompute with stash -> dX chain -> dW, per batch of stash tiles ----
            PeBwdTcArgs ta = {};
            ta.f = fa;
            if (prepass) { ta.f.bent = o.bent; ta.f.flags = o.flags; }
            ta.slot_list = b.slot_list; ta.slot_count = b.slot_count; ta.tile_begin = b.tile_begin;
            ta.tile_capacity = (int32_t)bw.tc_capacity; ta.stash = bw.tc_stash; ta.bstash = bw.tc_stash;
            ta.g_deformation = grad_in->deformation[k];
            const bool bender_tc = prepass && pe_bwd_tc_bender_ok(d, L);
            if (prepass) PE_CUDA_CHECK(cudaMemsetAsync(b.g_bent, 0, (size_t)s.images * s.rays * d.positions * 12, stream));
            ta.cw_obj = b.cw_obj; ta.cw_glob = b.cw_glob;
            ta.g_feat_obj = grad_out->object[k].integrated_features; ta.g_feat_glob = grad_out->global.integrated_features;
            ta.g_raw = b.g_raw; ta.g_dm = b.g_dm;
            ta.scale = b.scale; ta.bn_fix = b.bn_fix;
            ta.g_pos = b.g_pos; ta.g_bent = prepass ? b.g_bent : nullptr;
            ta.gw = grad_in->params[k];
            ta.adain_sums = b.adain_sums; ta.bn_sums = b.bn_sums;
            rc = pe_launch_bwd_scale(ta, b.scale, (unsigned int*)(b.scale + 8), stream); if (rc) return rc;
            const int64_t ub = bwd_tc_tiles_upper_bound(s, k);
            const int64_t batches = (ub + bw.tc_capacity - 1) / bw.tc_capacity;
            if (s.training) {
                // train-mode BatchNorm backward: two global reductions before the full pass (second, then first AdaIn layer of the head)
                for (int64_t bt = 0; bt < batches; ++bt) {
                    rc = pe_launch_bwd_fwd(ta, bt * bw.tc_capacity, sm_count, stream); if (rc) return rc;
                    rc = pe_launch_bwd_chain(ta, bt * bw.tc_capacity, 1, sm_count, stream); if (rc) return rc;
                }
                rc = pe_launch_bn_fix(o.stats + 2 * W + 2, b.bn_sums + 2 * W, W / 2, b.bn_fix + 2 * W, stream); if (rc) return rc;
                for (int64_t bt = 0; bt < batches; ++bt) {
                    if (batches > 1) { rc = pe_launch_bwd_fwd(ta, bt * bw.tc_capacity, sm_count, stream); if (rc) return rc; }
                    rc = pe_launch_bwd_chain(ta, bt * bw.tc_capacity, 2, sm_count, stream); if (rc) return rc;
                }
                rc = pe_launch_bn_fix(o.stats, b.bn_sums, W, b.bn_fix, stream); if (rc) return rc;
            }
            for (int64_t bt = 0; bt < batches; ++bt) {
                if (batches > 1 || !s.training) { rc = pe_launch_bwd_fwd(ta, bt * bw.tc_capacity, sm_count, stream); if (rc) return rc; }
                rc = pe_launch_bwd_chain(ta, bt * bw.tc_capacity, 0, sm_count, stream); if (rc) return rc;
                rc = pe_launch_bwd_dw(ta, bt * bw.tc_capacity, sm_count, stream); if (rc) return rc;
                if (bender_tc) {
                    // the ray bender of this batch's tiles, on the same stash memory: dL/d bent position -> bender -> g_pos, parameter gradients
                    rc = pe_launch_bwd_scale_bender(ta, b.scale, (unsigned int*)(b.scale + 8), stream); if (rc) return rc;
                    rc = pe_launch_bwd_bender(ta, bt * bw.tc_capacity, sm_count, stream); if (rc) return rc;
                }
            }
            if (prepass && !bender_tc) {
                // the ray bender's backward stays on the exact fp32 kernel (tiles of 32 listed slots): dL/d bent position -> bender -> g_pos
                rc = pe_launch_compact_slots(o.flags, 1, s.images, (int64_t)s.rays * d.positions, b.slot_list, b.slot_count, b.tile_begin, stream, 32);
                if (rc) return rc;
                fb.g_bent_in = b.g_bent;
                fb.bwd_phase = 0;
                rc = pe_launch_field_bwd(fb, sm_count, stream); if (rc) return rc;
            }
        } else {
        if (s.training) {
            fb.bwd_phase = 1;
            rc = pe_launch_field_bwd(fb, sm_count, stream); if (rc) return rc;
            rc = pe_launch_bn_fix(o.stats + 2 * W + 2, b.bn_sums + 2 * W, W / 2, b.bn_fix + 2 * W, stream); if (rc) return rc;
            fb.bwd_phase = 2;
            rc = pe_launch_field_bwd(fb, sm_count, stream); if (rc) return rc;
            rc = pe_launch_bn_fix(o.stats, b.bn_sums, W, b.bn_fix, stream); if (rc) return rc;
        }
        fb.bwd_phase = 0;
        rc = pe_launch_field_bwd(fb, sm_count, stream); if (rc) return rc;
        }

        const unsigned char* blob = (const unsigned char*)d.packed;
        auto P32 = [&](int64_t off) { return (const float*)(blob + off); };
        PeStyleBwdArgs sb = {};
        sb.images = s.images; sb.style_features = d.style_features; sb.channels = W; sb.training = s.training;
        sb.style = in->style[k]; sb.aff_w = P32(L.aff1_w); sb.run_mean = P32(L.bn1_mean); sb.run_var = P32(L.bn1_var);
        sb.stats = o.stats; sb.adain_sums = b.adain_sums; sb.adain_stride = 3 * W;
        sb.g_aff_w = grad_in->params[k].affine1_w; sb.g_aff_b = grad_in->params[k].affine1_b; sb.g_style = grad_in->style[k];
        rc = pe_launch_style_bwd(sb, stream); if (rc) return rc;
        sb.channels = W / 2; sb.aff_w = P32(L.aff2_w); sb.run_mean = P32(L.bn2_mean); sb.run_var = P32(L.bn2_var);
        sb.stats = o.stats + 2 * W + 2; sb.adain_sums = b.adain_sums + 2 * W;
        sb.g_aff_w = grad_in->params[k].affine2_w; sb.g_aff_b = grad_in->params[k].affine2_b;
        rc = pe_launch_style_bwd(sb, stream); if (rc) return rc;

        if (grad_in->ray_origins || grad_in->ray_directions || grad_in->w2o || (in->sample_t[k] && grad_in->sample_t[k])) {
            PeGeometryBwdArgs gb = {};
            gb.ob = d; gb.images = s.images; gb.rays = s.rays; gb.objects = s.objects; gb.k = k; gb.perturb = s.perturb;
            gb.origins = in->ray_origins; gb.dirs = in->ray_directions; gb.w2o = in->w2o; gb.ois = in->object_in_scene; gb.rand = in->rand[k];
            gb.t_in = in->sample_t[k]; gb.g_t_in = in->sample_t[k] ? grad_in->sample_t[k] : nullptr;
            gb.g_pos = b.g_pos; gb.g_t = b.g_t; gb.g_od = b.g_od;
            gb.g_origins = grad_in->ray_origins; gb.g_dirs = grad_in->ray_directions; gb.g_w2o = grad_in->w2o;
            rc = pe_launch_geometry_bwd(gb, stream); if (rc) return rc;
        }

        if (s.bent_gradients && grad_out->bent_positions[k]) {
            // backward of forward_expected_positions (object_composer.py:603-722): an upstream gradient on the bent sample positions
            // x + displacement(x).  The backward is linear in its upstream gradients, so this is one more pass: the ray bender
            // (ray-bender-only mode of the fp32 field backward: parameter / deformation gradients accumulate, dL/dx to g_pos2), then the
            // ray geometry once more.
            const float* g_x = grad_out->bent_positions[k];
            if (d.bender_kind == PE_BENDER_POSITIONAL) {
                PeFieldBwdArgs xb = {};
                xb.f = fb.f; xb.w = fb.w; xb.gw = fb.gw;
                xb.g_deformation = grad_in->deformation[k];
                xb.g_bent_in = g_x; xb.g_bent_flag = 1;
                xb.g_pos = b.g_pos2;
                xb.stash = bw.stash; xb.stash_floats = bw.stash_floats;
                xb.bwd_phase = 0;
                rc = pe_launch_field_bwd(xb, sm_count, stream); if (rc) return rc;
                g_x = b.g_pos2;
            }
            if (grad_in->ray_origins || grad_in->ray_directions || grad_in->w2o || (in->sample_t[k] && grad_in->sample_t[k])) {
                PeGeometryBwdArgs gb = {};
                gb.ob = d; gb.images = s.images; gb.rays = s.rays; gb.objects = s.objects; gb.k = k; gb.perturb = s.perturb;
                gb.origins = in->ray_origins; gb.dirs = in->ray_directions; gb.w2o = in->w2o; gb.ois = in->object_in_scene; gb.rand = in->rand[k];
                gb.t_in = in->sample_t[k]; gb.g_t_in = nullptr;
                gb.g_pos = g_x; gb.g_t = nullptr; gb.g_od = nullptr;
                gb.g_origins = grad_in->ray_origins; gb.g_dirs = grad_in->ray_directions; gb.g_w2o = grad_in->w2o;
                rc = pe_launch_geometry_bwd(gb, stream); if (rc) return rc;
            }
        }
    }
    return PE_OK;
}
