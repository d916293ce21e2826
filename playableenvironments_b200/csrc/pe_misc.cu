// Small kernels around the field evaluation: style prologue (AdaIn affine with folded BatchNorm),
// parameter packing, ray generation, stand-alone positional encoding and the decoder hand-off fold.
#include "pe_kernels.cuh"
#include <stdlib.h>

namespace {

constexpr float BN_EPS = 1e-5f;       // torch.nn.BatchNorm1d default (model/layers/adain.py:47)

// [scale|bias] = affine_transform(style) (adain.py:30-32); BatchNorm1d(affine=False) folded: y = x*sc + sh.
__global__ void __launch_bounds__(256) pe_style_kernel(const PeStyleArgs A) {
    // one warp per channel (the two rows of the affine transform read coalesced, lanes stride the style features): this kernel sits
    // on the critical path in front of every field launch, twice per object
    const int img = blockIdx.x;
    const int C = A.channels, S = A.style_features;
    const float* style = A.style + (int64_t)img * S;
    const int lane = threadIdx.x & 31;
    for (int c = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); c < C; c += gridDim.y * (blockDim.x >> 5)) {
        float scale = 0.f, bias = 0.f;
        for (int s = lane; s < S; s += 32) {
            const float v = style[s];
            scale = fmaf(A.aff_w[(int64_t)c * S + s], v, scale);
            bias = fmaf(A.aff_w[(int64_t)(C + c) * S + s], v, bias);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            scale += __shfl_xor_sync(0xffffffffu, scale, o);
            bias += __shfl_xor_sync(0xffffffffu, bias, o);
        }
        if (lane != 0) continue;
        scale += A.aff_b[c];
        bias += A.aff_b[C + c];
        float mean = A.run_mean[c], var = A.run_var[c];
        if (A.training) {
            const double n = A.stats[2 * C];          // in-box sample count, written after the sums
            if (n > 0.0) {
                const double m = A.stats[c] / n;
                const double v = fmax(A.stats[C + c] / n - m * m, 0.0);     // biased variance normalises
                mean = (float)m; var = (float)v;
                if (img == 0 && A.running_out) {          // batch mean and UNBIASED variance: the host applies the momentum update
                    A.running_out[c] = (float)m;
                    A.running_out[C + c] = (float)(n > 1.0 ? v * n / (n - 1.0) : v);
                }
            } else if (img == 0 && A.running_out) {       // no in-box sample: the reference leaves the statistics untouched
                A.running_out[c] = A.run_mean[c];
                A.running_out[C + c] = A.run_var[c];
            }
        }
        const float sc = scale / sqrtf(var + BN_EPS);
        A.out[((int64_t)img * 2 + 0) * C + c] = sc;
        A.out[((int64_t)img * 2 + 1) * C + c] = bias - mean * sc;
    }
}

__global__ void pe_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int K) {
    // src [N][K] (nn.Linear) -> dst [K][N]
    const int64_t total = (int64_t)N * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / N), n = (int)(i - (int64_t)k * N);
        dst[i] = src[(int64_t)n * K + k];
    }
}

__global__ void pe_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// The fp32 section of the packed blob in ONE launch: a table of copies (K == 0: N floats) and transposes ([N][K] -> [K][N]), one
// item per blockIdx.y.  (Packing runs for every model whenever a parameter changes -- every training step -- and was ~50 launches.)
constexpr int PACK_TABLE = 64;
struct PackTable {
    const float* src[PACK_TABLE];
    float* dst[PACK_TABLE];
    int32_t N[PACK_TABLE], K[PACK_TABLE];
};
__global__ void pe_pack_table_kernel(const __grid_constant__ PackTable T) {
    const int it = blockIdx.y;
    const float* __restrict__ src = T.src[it];
    float* __restrict__ dst = T.dst[it];
    const int N = T.N[it], K = T.K[it];
    const int64_t total = K ? (int64_t)N * K : N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (K) { const int k = (int)(i / N), n = (int)(i - (int64_t)k * N); dst[i] = src[(int64_t)n * K + k]; }
        else dst[i] = src[i];
    }
}

// PositionalEncoder.forward / AnnealablePositionalEncoder.forward (model/positional_encoder.py:41-65,
// model/annealable_positional_encoder.py:46-76)
__global__ void pe_posenc_kernel(const float* __restrict__ x, int64_t n, int dims, int octaves, int append,
                                 const float* __restrict__ weights, float* __restrict__ out) {
    const int E = dims * (append + 2 * octaves);
    const int64_t total = n * E;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / E;
        int e = (int)(i - row * E);
        float v;
        if (append && e < dims) {
            v = x[row * dims + e];
        } else {
            if (append) e -= dims;
            const int oct = e / (2 * dims);
            const int rem = e - oct * 2 * dims;
            const int fn = rem / dims, dim = rem - fn * dims;
            const float a = __fmul_rn(exp2f((float)oct), x[row * dims + dim]);
            v = fn ? cosf(a) : sinf(a);
            if (weights) v = __fmul_rn(v, weights[oct]);
        }
        out[i] = v;
    }
}

// RayHelper.create_camera_rays (utils/lib_3d/ray_helper.py:15-52) evaluated only at the pixels picked by
// sample_all_rays_strided_grid (:433-482, centre pixel idx*s + s//2), then transform_rays with c2w (:1203-1227).
__global__ void pe_rays_kernel(const float* __restrict__ focal, const float* __restrict__ c2w, int images, int H, int W,
                               const int* __restrict__ strides, int n_strides, int R, float* __restrict__ dirs,
                               float* __restrict__ origins, float* __restrict__ positions) {
    const int64_t total = (int64_t)images * R;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int img = (int)(i / R);
        int r = (int)(i - (int64_t)img * R);
        int s = 1, gw = W;
        for (int q = 0; q < n_strides; ++q) {
            s = strides[q];
            gw = W / s;
            const int cnt = (H / s) * gw;
            if (r < cnt) break;
            r -= cnt;
        }
        const int row = (r / gw) * s + s / 2, col = (r % gw) * s + s / 2;
        const float f = focal[img];
        const float d[3] = {__fdiv_rn((float)col - (float)W / 2.f, f), -__fdiv_rn((float)row - (float)H / 2.f, f), -1.f};
        const float* m = c2w + (int64_t)img * 12;
        float o[3];
        pe_transform(m, d, false, o);
        dirs[i * 3 + 0] = o[0]; dirs[i * 3 + 1] = o[1]; dirs[i * 3 + 2] = o[2];
        if (positions) { positions[i * 2 + 0] = (float)row / (float)H; positions[i * 2 + 1] = (float)col / (float)W; }
        if (r == 0 && origins && (i - (int64_t)img * R) == 0) {
            origins[img * 3 + 0] = m[3]; origins[img * 3 + 1] = m[7]; origins[img * 3 + 2] = m[11];
        }
    }
}

// fold_strided_tensors + per-stride channel split + HWC->CHW
// (environment_model_backpropagated_autoencoder.py:129-168, ..._multiresolution_backpropagated_autoencoder.py:59-99)
__global__ void pe_fold_kernel(const float* __restrict__ feats, int images, int R, int F, int ray0, int gh, int gw,
                               int c0, int channels, float* __restrict__ grid) {
    const int64_t total = (int64_t)images * channels * gh * gw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % gw);
        const int y = (int)((i / gw) % gh);
        const int c = (int)((i / ((int64_t)gw * gh)) % channels);
        const int img = (int)(i / ((int64_t)gw * gh * channels));
        grid[i] = feats[((int64_t)img * R + ray0 + y * gw + x) * F + c0 + c];
    }
}

inline int grid_for(int64_t n, int block = 256) { return (int)pe_min64((n + block - 1) / block, 148 * 8); }

}  // namespace

int pe_launch_style(const PeStyleArgs& args, cudaStream_t stream) {
    if (args.images == 0) return PE_OK;
    pe_style_kernel<<<dim3(args.images, (args.channels + 7) / 8), 256, 0, stream>>>(args);
    PE_LAUNCH_CHECK("pe_style_kernel");
    return PE_OK;
}

int pe_tc_pack(const PeObjectDesc& desc, const PeLayout& L, const PeObjectParams& params, void* packed, cudaStream_t stream);

int pe_launch_pack(const PeObjectDesc& d, const PeLayout& L, const PeObjectParams& p, void* packed, cudaStream_t stream) {
    int rc;
    const int W = d.width, F = d.features, S = d.style_features;
    PackTable table;
    int items = 0;
    // transposes and copies are collected and launched together below
    auto transpose_to = [&](const float* src, void* blob, int64_t off, int N, int K, cudaStream_t) -> int {
        if (!src) { pe_set_error("missing parameter tensor"); return PE_ERR_INVALID; }
        if (items == PACK_TABLE) { pe_set_error("internal: pack table overflow"); return PE_ERR_INVALID; }
        table.src[items] = src; table.dst[items] = reinterpret_cast<float*>((char*)blob + off); table.N[items] = N; table.K[items] = K;
        ++items;
        return PE_OK;
    };
    auto copy_to = [&](const float* src, void* blob, int64_t off, int64_t n, cudaStream_t) -> int {
        if (!src) { pe_set_error("missing parameter tensor"); return PE_ERR_INVALID; }
        if (items == PACK_TABLE) { pe_set_error("internal: pack table overflow"); return PE_ERR_INVALID; }
        table.src[items] = src; table.dst[items] = reinterpret_cast<float*>((char*)blob + off); table.N[items] = (int)n; table.K[items] = 0;
        ++items;
        return PE_OK;
    };
#define PE_TRY(x) do { rc = (x); if (rc != PE_OK) return rc; } while (0)
    for (int l = 0; l < d.layers; ++l) {
        PE_TRY(transpose_to(p.backbone_w[l], packed, L.bb_w[l], W, L.k_in[l], stream));
        PE_TRY(copy_to(p.backbone_b[l], packed, L.bb_b[l], W, stream));
    }
    if (d.nerf_kind == PE_NERF_ADAIN) {
        PE_TRY(copy_to(p.alpha_w, packed, L.alpha_w, W, stream));
        PE_TRY(copy_to(p.alpha_b, packed, L.alpha_b, 1, stream));
    }
    PE_TRY(transpose_to(p.head0_w, packed, L.head0_w, W, W, stream));
    PE_TRY(transpose_to(p.head3_w, packed, L.head3_w, W / 2, W, stream));
    PE_TRY(transpose_to(p.head6_w, packed, L.head6_w, F, W / 2, stream));
    PE_TRY(copy_to(p.head6_b, packed, L.head6_b, F, stream));
    PE_TRY(copy_to(p.affine1_w, packed, L.aff1_w, (int64_t)2 * W * S, stream));
    PE_TRY(copy_to(p.affine1_b, packed, L.aff1_b, 2 * W, stream));
    PE_TRY(copy_to(p.bn1_mean, packed, L.bn1_mean, W, stream));
    PE_TRY(copy_to(p.bn1_var, packed, L.bn1_var, W, stream));
    PE_TRY(copy_to(p.affine2_w, packed, L.aff2_w, (int64_t)W * S, stream));
    PE_TRY(copy_to(p.affine2_b, packed, L.aff2_b, W, stream));
    PE_TRY(copy_to(p.bn2_mean, packed, L.bn2_mean, W / 2, stream));
    PE_TRY(copy_to(p.bn2_var, packed, L.bn2_var, W / 2, stream));
    if (d.bender_kind == PE_BENDER_POSITIONAL) {
        for (int l = 0; l < d.b_layers; ++l) {
            PE_TRY(transpose_to(p.bender_w[l], packed, L.bd_w[l], d.b_width, L.b_k_in[l], stream));
            PE_TRY(copy_to(p.bender_b[l], packed, L.bd_b[l], d.b_width, stream));
        }
        PE_TRY(transpose_to(p.bender_out_w, packed, L.bd_out_w, 3, d.b_width, stream));
    }
    if (items) {
        pe_pack_table_kernel<<<dim3(32, items), 256, 0, stream>>>(table);
        PE_LAUNCH_CHECK("pe_pack_table_kernel");
    }
    if (L.tc_supported) {
        PE_TRY(pe_tc_pack(d, L, p, packed, stream));
        PE_TRY(pe_tcT_pack(d, L, p, packed, stream));
    }
#undef PE_TRY
    return PE_OK;
}

extern "C" int pe_positional_encoding(const float* x, int64_t n, int32_t dims, int32_t octaves, int32_t append_original,
                                      const float* weights, float* out, pe_stream_t stream) {
    if (dims <= 0 || octaves < 0 || octaves > PE_MAX_OCTAVES) { pe_set_error("bad positional encoding shape"); return PE_ERR_INVALID; }
    if (n == 0) return PE_OK;
    const int E = dims * ((append_original ? 1 : 0) + 2 * octaves);
    pe_posenc_kernel<<<grid_for(n * E), 256, 0, (cudaStream_t)stream>>>(x, n, dims, octaves, append_original ? 1 : 0, weights, out);
    PE_LAUNCH_CHECK("pe_posenc_kernel");
    return PE_OK;
}

extern "C" int pe_generate_rays(const float* focal, const float* c2w, int32_t images, int32_t height, int32_t width,
                                const int32_t* strides, int32_t n_strides, float* directions, float* origins,
                                float* positions, pe_stream_t stream) {
    if (n_strides <= 0 || n_strides > 8) { pe_set_error("1..8 strides supported"); return PE_ERR_INVALID; }
    int R = 0;
    for (int q = 0; q < n_strides; ++q) {
        if (strides[q] <= 0 || height % strides[q] || width % strides[q]) {
            pe_set_error("The image size is not divisible by the stride");   // ray_helper.py:548-551
            return PE_ERR_INVALID;
        }
        R += (height / strides[q]) * (width / strides[q]);
    }
    // strides are passed by value through a tiny device copy owned by the caller's stream
    int* d_strides = nullptr;
    PE_CUDA_CHECK(cudaMallocAsync((void**)&d_strides, sizeof(int) * n_strides, (cudaStream_t)stream));
    PE_CUDA_CHECK(cudaMemcpyAsync(d_strides, strides, sizeof(int) * n_strides, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    pe_rays_kernel<<<grid_for((int64_t)images * R), 256, 0, (cudaStream_t)stream>>>(focal, c2w, images, height, width, d_strides, n_strides, R,
                                                                                     directions, origins, positions);
    PE_LAUNCH_CHECK("pe_rays_kernel");
    PE_CUDA_CHECK(cudaFreeAsync(d_strides, (cudaStream_t)stream));
    return PE_OK;
}

extern "C" int pe_fold_feature_grids(const float* features, int32_t images, int32_t height, int32_t width, int32_t n_features,
                                     const int32_t* strides, const int32_t* channels, int32_t n_strides,
                                     float* const* grids, pe_stream_t stream) {
    int R = 0;
    for (int q = 0; q < n_strides; ++q) R += (height / strides[q]) * (width / strides[q]);
    int ray0 = 0, c0 = 0;
    for (int q = 0; q < n_strides; ++q) {
        const int gh = height / strides[q], gw = width / strides[q];
        if (c0 + channels[q] > n_features) { pe_set_error("channel split exceeds the feature count"); return PE_ERR_INVALID; }
        const int64_t total = (int64_t)images * channels[q] * gh * gw;
        if (total) {
            pe_fold_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(features, images, R, n_features, ray0, gh, gw, c0, channels[q], grids[q]);
            PE_LAUNCH_CHECK("pe_fold_kernel");
        }
        ray0 += gh * gw;
        c0 += channels[q];
    }
    return PE_OK;
}
