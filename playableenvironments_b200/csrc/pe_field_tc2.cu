// Fused ray-march kernel, CTA-pair variant: cluster of 2 CTAs driving tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA).
//
// Why a pair: each CTA stores only its N/2 rows of every weight slab (the pair's tensor cores exchange B halves), so a 66 KB ring
// per CTA holds a WHOLE 256x256 layer.  That makes a ping-pong schedule possible with every slab fetched once per 512 sample rows:
//
//     MMA   X(l) | Y(l) | X(l+1) | Y(l+1) | ...          (leader CTA's elected thread issues for both CTAs)
//     epi        | X(l) | Y(l)   | X(l+1) | ...          (each CTA's epilogue warps, on their own 128 rows of X and Y)
//
// i.e. the layer epilogues, the sampling/encoding prologue and the volume-rendering finale all overlap the other tile's MMAs; L2->SM
// weight traffic is a quarter of a one-tile-per-CTA design.  The per-tile work of the epilogue warps is shared with the
// single-CTA kernel (pe_tc_common.cuh); only the handshakes differ (cluster-scope mbarriers, remote arrives, multicast commits).
#include "pe_tc_common.cuh"
#include <stdlib.h>

namespace {
using namespace pe;
using namespace pe_tc;

constexpr int THREADS = 384;
constexpr int NUM_STAGES2 = 8;
constexpr int STAGE2_BYTES = 8192;                          // half slab: 128 rows x 32 k x 2 B
constexpr int BIAS2_OFF = NUM_STAGES2 * STAGE2_BYTES;       // the bias chunk (<= 2 KB) sits behind the last stage
constexpr int RING2_BYTES = BIAS2_OFF + 2048;
constexpr int SMEM2_BAR = 2 * A_BYTES + RING2_BYTES;
constexpr int SMEM2_ONES = SMEM2_BAR + 256;
constexpr int SMEM2_TOTAL = SMEM2_ONES + 256;

struct Sync2 {        // epilogue <-> leader-CTA MMA handshakes across the pair
    uint64_t* acc_full;        // local barrier, signalled by the leader's multicast tcgen05.commit
    uint32_t ready_addr;       // shared::cluster address of the LEADER's a_ready barrier of this group
    uint32_t phase;
    long long waited;
    long long* ts;             // optional timeline (debug)
    int n;
    __device__ __forceinline__ void wait_acc() {
        const long long t0 = clock64(); mbar_wait(acc_full, phase); const long long t1 = clock64();
        waited += t1 - t0; phase ^= 1; tc_fence_after();
        if (ts) { ts[n++] = t0; ts[n++] = t1; }
    }
    __device__ __forceinline__ void arrive_ready() { fence_proxy_async(); tc_fence_before(); mbar_arrive_cluster(ready_addr); }
    __device__ __forceinline__ void arrive_fold() {}      // folded-head mode is not used by this kernel
};

// one "sub-layer": up to 8 ring entries (slab x pass) consumed by tile X, then by tile Y
struct SubLayer { int slab0, nslabs, entries; bool first, last; };

__device__ __forceinline__ int sublayer_count(int slabs, int num_passes) { const int per = NUM_STAGES2 / num_passes; return (slabs + per - 1) / per; }
__device__ __forceinline__ SubLayer sublayer(int slabs, int num_passes, int i) {
    const int per = NUM_STAGES2 / num_passes;
    SubLayer s;
    s.slab0 = i * per;
    s.nslabs = min(per, slabs - s.slab0);
    s.entries = s.nslabs * num_passes;
    s.first = i == 0;
    s.last = s.slab0 + s.nslabs == slabs;
    return s;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
pe_field_tc2_kernel(const PeFieldArgs A, const PeIntegrated G2, const int num_passes, const int dbg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem + 2 * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM2_BAR);     // [8] own half slab landed
    uint64_t* peer_full = full_bar + NUM_STAGES2;                            // [8] (leader only) peer's half slab landed
    uint64_t* empty_bar = peer_full + NUM_STAGES2;                           // [8] stage consumed by both tiles
    uint64_t* acc_full = empty_bar + NUM_STAGES2;                            // [2]
    uint64_t* a_ready = acc_full + 2;                                        // [2] (leader only) 256 arrivals: both CTAs' rows
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);
    unsigned char* ones = smem + SMEM2_ONES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);

    const int P = ob.positions;
    const int rpt = TILE_M / P;
    const int tiles_per_image = (A.rays + rpt - 1) / rpt;
    const int64_t total_tiles = (int64_t)tiles_per_image * A.images;
    const int64_t total_iters = (total_tiles + 3) / 4;               // 4 tiles of 128 rows per cluster iteration
    const int64_t cluster_id = blockIdx.x >> 1, cluster_stride = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_STAGES2; ++s) { mbar_init(full_bar + s, 1); mbar_init(peer_full + s, 1); mbar_init(empty_bar + s, 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(acc_full + g, 1); mbar_init(a_ready + g, 2 * TILE_M); }
        mbar_fence_init();
    }
    if (threadIdx.x < 128) {
        const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
        reinterpret_cast<__half*>(ones)[threadIdx.x] = __float2half_rn((r < 8 && c < 2) ? 1.f : 0.f);
    }
    fence_proxy_async_all();
    if (warp == 2) tmem_alloc_2cta(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // barriers of both CTAs are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ weight producer (each CTA: its N/2 rows of every slab) ================================
        if (elect_one()) {
            uint32_t empty_bits = 0;                              // per-stage phase parity
            for (int64_t it = cluster_id; it < total_iters; it += cluster_stride) {
                const unsigned char* layer_src = blob + L.tc2_base;
                for (int l = 0; l < NUM_LAYERS; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
                    const uint32_t half_bytes = (uint32_t)n * 32;                 // (n/2) rows x 32 k x 2 B
                    const int nsub = sublayer_count(slabs, num_passes);
                    for (int si = 0; si < nsub; ++si) {
                        const SubLayer sl = sublayer(slabs, num_passes, si);
                        for (int e = 0; e < sl.entries; ++e) {
                            const int stage = NUM_STAGES2 - sl.entries + e;
                            const int s = sl.slab0 + e / num_passes, pass = e % num_passes;
                            const bool with_bias = has_bias && sl.last && e == sl.entries - 1;
                            mbar_wait(empty_bar + stage, ((empty_bits >> stage) & 1) ^ 1);
                            empty_bits ^= 1u << stage;
                            mbar_arrive_expect_tx(full_bar + stage, half_bytes + (with_bias ? (uint32_t)n * 8 : 0u));
                            bulk_copy_g2s(ring + stage * STAGE2_BYTES,
                                          layer_src + (int64_t)pass * L.tc_bytes_per_pass + (int64_t)s * n * 64 + (int64_t)rank * half_bytes,
                                          half_bytes, full_bar + stage);
                            if (with_bias)
                                bulk_copy_g2s(ring + BIAS2_OFF, layer_src + (int64_t)slabs * n * 64 + (int64_t)rank * n * 8, (uint32_t)n * 8, full_bar + stage);
                        }
                    }
                    layer_src += (int64_t)slabs * n * 64 + (has_bias ? n * 16 : 0);
                }
            }
        }
    } else if (warp == 1 && rank == 1) {
        // ================================ relay: tell the leader that this CTA's half slab has landed ================================
        if (elect_one()) {
            uint32_t full_bits = 0;
            const uint32_t peer_full_leader = mapa_u32(smem_u32(peer_full), 0);
            for (int64_t it = cluster_id; it < total_iters; it += cluster_stride) {
                for (int l = 0; l < NUM_LAYERS; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
                    const int nsub = sublayer_count(slabs, num_passes);
                    for (int si = 0; si < nsub; ++si) {
                        const SubLayer sl = sublayer(slabs, num_passes, si);
                        for (int e = 0; e < sl.entries; ++e) {
                            const int stage = NUM_STAGES2 - sl.entries + e;
                            mbar_wait(full_bar + stage, (full_bits >> stage) & 1);
                            full_bits ^= 1u << stage;
                            mbar_arrive_cluster(peer_full_leader + stage * 8);
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ================================ MMA issuer (leader CTA, one thread, drives both CTAs' tensor cores) ================================
        if (elect_one()) {
            uint32_t full_bits = 0, ready_phase[2] = {0u, 0u};
            long long t_ready = 0, t_full = 0, t_peer = 0, t_begin = clock64();
            const uint32_t a_addr[2] = {smem_u32(smem), smem_u32(smem + A_BYTES)};
            const uint32_t ring_addr = smem_u32(ring);
            const uint64_t ones_desc = umma_smem_desc(smem_u32(ones), 128, 0);
            int iter = 0;
            long long* mts = reinterpret_cast<long long*>(A.stats);
            for (int64_t it = cluster_id; it < total_iters; it += cluster_stride, ++iter) {
                const bool rec = (dbg & 2) && blockIdx.x == 2 && iter == 5;
                if (rec) mts[60] = clock64();
                for (int l = 0; l < NUM_LAYERS; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
                    const uint32_t idesc = umma_idesc_f16(2 * TILE_M, n);
                    const uint32_t lbo_b = (uint32_t)n * 8;               // (n/2 rows / 8) core matrices x 128 B between K chunks
                    const uint64_t bias_desc = umma_smem_desc(ring_addr + BIAS2_OFF, 0, 128);     // one K chunk, aliased for the (zero) second half
                    const int nsub = sublayer_count(slabs, num_passes);
                    for (int si = 0; si < nsub; ++si) {
                        const SubLayer sl = sublayer(slabs, num_passes, si);
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            if (sl.first) {
                                const long long t0 = clock64();
                                mbar_wait(a_ready + g, ready_phase[g]);
                                const long long t1 = clock64();
                                t_ready += t1 - t0;
                                if (rec) { mts[4 * l + 2 * g] = t0; mts[4 * l + 2 * g + 1] = t1; }
                                ready_phase[g] ^= 1;
                                tc_fence_after();
                            }
                            for (int e = 0; e < sl.entries; ++e) {
                                const int stage = NUM_STAGES2 - sl.entries + e;
                                const int s = sl.slab0 + e / num_passes;
                                if (g == 0) {                     // first use of the stage in this sub-layer: both halves must have landed
                                    const long long t0 = clock64();
                                    mbar_wait(full_bar + stage, (full_bits >> stage) & 1);
                                    const long long t1 = clock64();
                                    mbar_wait(peer_full + stage, (full_bits >> stage) & 1);
                                    t_full += t1 - t0; t_peer += clock64() - t1;
                                    tc_fence_after();
                                }
                                const uint32_t b_addr = ring_addr + stage * STAGE2_BYTES;
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const uint32_t a_chunk = chunk0 + 4 * s + 2 * j;
                                    const uint64_t da = umma_smem_desc(a_addr[g] + a_chunk * CHUNK_BYTES, CHUNK_BYTES, 128);
                                    const uint64_t db = umma_smem_desc(b_addr + 2 * j * lbo_b, lbo_b, 128);
                                    umma_f16_ss_2cta(tmem_base + g * 256, da, db, idesc, (si | e | j) != 0 ? 1u : 0u);
                                }
                                const bool with_bias = has_bias && sl.last && e == sl.entries - 1;
                                if (with_bias) umma_f16_ss_2cta(tmem_base + g * 256, ones_desc, bias_desc, idesc, 1u);
                                if (g == 1) umma_commit_2cta(empty_bar + stage, 3);        // both tiles are done with the stage, in both CTAs
                            }
                            if (sl.last) umma_commit_2cta(acc_full + g, 3);
                        }
                        for (int e = 0; e < sl.entries; ++e) full_bits ^= 1u << (NUM_STAGES2 - sl.entries + e);
                    }
                }
                if (rec) {
                    printf("PE_TC2 mma start=%lld |", mts[60]);
                    for (int l = 0; l < NUM_LAYERS; ++l)
                        printf(" L%d X wait@%lld +%lld Y wait@%lld +%lld", l, mts[4 * l] - mts[60], mts[4 * l + 1] - mts[4 * l], mts[4 * l + 2] - mts[60], mts[4 * l + 3] - mts[4 * l + 2]);
                    printf(" | end=%lld\n", clock64() - mts[60]);
                }
            }
            if (dbg && blockIdx.x == 2)
                printf("PE_TC2 mma thread: total=%lld cycles, wait a_ready=%lld, wait full(own)=%lld, wait full(peer)=%lld\n",
                       clock64() - t_begin, t_ready, t_full, t_peer);
        }
    } else if (warp >= 4) {
        // ================================ epilogue groups (own 128 rows of tile X / Y) ================================
        const int g = (warp - 4) >> 2;
        TileCtx X;
        X.A = &A; X.G2 = &G2;
        X.abuf = smem + g * A_BYTES;
        X.abuf_lo = nullptr;
        X.m = ((warp & 3) << 5) | lane; X.lane = lane; X.wq = warp & 3; X.half = 0; X.gw = warp & 3;
        X.taddr = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + g * 256;
        X.bar_id = 1 + g;
        X.P = P; X.rpt = rpt; X.rows_used = rpt * P; X.tiles_per_image = tiles_per_image; X.total_tiles = total_tiles;
        X.size[0] = ob.bbox[1] - ob.bbox[0]; X.size[1] = ob.bbox[3] - ob.bbox[2]; X.size[2] = ob.bbox[5] - ob.bbox[4];
        X.alpha_bias = __ldg(reinterpret_cast<const float*>(blob + L.alpha_b));
        X.alpha_w = reinterpret_cast<const float*>(blob + L.alpha_w);
        X.dbg = dbg;
        X.fold = false;
        X.stat_phase = 0;
        X.single = G2.integrated_features != nullptr || G2.opacity != nullptr || G2.weights != nullptr;
        Sync2 sync{acc_full + g, mapa_u32(smem_u32(a_ready + g), 0), 0u, 0LL, nullptr, 0};
        const long long t_begin = clock64();
        int iter = 0;
        for (int64_t it = cluster_id; it < total_iters; it += cluster_stride, ++iter) {
            const bool rec = (dbg & 2) && blockIdx.x == 2 && X.m == 0 && iter == 5;
            sync.ts = rec ? reinterpret_cast<long long*>(A.stats) + 64 + 32 * g : nullptr;
            sync.n = 0;
            if (rec) sync.ts[30] = clock64();
            TileAhead ahead;
            epilogue_tile<1, false>(X, it * 4 + 2 * g + rank, sync, ahead, 0);
            if (rec) {
                const long long* q = sync.ts;
                printf("PE_TC2 epi g%d start=%lld |", g, q[30]);
                for (int i = 0; i < 11; ++i) printf(" L%d wait@%lld got+%lld", i, q[2 * i] - q[30], q[2 * i + 1] - q[2 * i]);
                printf(" | end=%lld\n", clock64() - q[30]);
            }
        }
        if (dbg && blockIdx.x == 2 && X.m == 0)
            printf("PE_TC2 epilogue group %d: total=%lld cycles, waiting for accumulators=%lld\n", g, clock64() - t_begin, sync.waited);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // nobody leaves (or frees TMEM) while the peer can still signal or read
    if (warp == 2) tmem_dealloc_2cta(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------
// weight stream of the CTA-pair kernel: every slab stored as [rank 0: rows 0..N/2) | rank 1: rows N/2..N)], each half in the
// UMMA K-major no-swizzle layout of an (N/2)-row operand; bias chunk = (N/2) rows x 8 K columns [hi, lo, 0...] per rank.
// ------------------------------------------------------------------------------------------------------
__global__ void pe_tc2_pack_layer_kernel(const float* __restrict__ w, int N, int K_src, int K_pad, unsigned char* __restrict__ hi,
                                         unsigned char* __restrict__ lo) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int H = N / 2, r = n / H, nr = n - r * H;
    float run = 0.f;
    for (int k = 0; k < K_pad; ++k) {
        const float v = k < K_src ? w[(int64_t)n * K_src + k] : 0.f;
        const __half near = __float2half_rn(v);
        const float fn = __half2float(near);
        __half other = near;
        if (fn != v) other = fn < v ? __float2half_ru(v) : __float2half_rd(v);
        const float e_near = fn - v, e_other = __half2float(other) - v;
        const bool pick_other = fabsf(run + e_other) < fabsf(run + e_near);       // zero-sum rounding, see pe_field_tc.cu
        const __half h = pick_other ? other : near;
        run += pick_other ? e_other : e_near;
        const int slab = k / PE_TC_SLAB_K, kk = k - slab * PE_TC_SLAB_K;
        const int64_t off = (int64_t)slab * N * 64 + (int64_t)r * N * 32 + (int64_t)(kk >> 3) * (H * 16) + (nr >> 3) * 128 + (nr & 7) * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = __float2half_rn(v - __half2float(h));
    }
}

__global__ void pe_tc2_pack_bias_kernel(const float* __restrict__ bias, int N, unsigned char* __restrict__ dst) {
    const int total = N * 8;
    const int H = N / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / 8, kk = i - n * 8;
        const int r = n / H, nr = n - r * H;
        const float v = bias[n];
        const __half h = __float2half_rn(v);
        __half out = __float2half_rn(0.f);
        if (kk == 0) out = h;
        if (kk == 1) out = __float2half_rn(v - __half2float(h));
        const int64_t off = (int64_t)r * N * 8 + (nr >> 3) * 128 + (nr & 7) * 16 + kk * 2;
        *reinterpret_cast<__half*>(dst + off) = out;
    }
}

}  // namespace

int pe_tc2_pack(const PeObjectDesc& d, const PeLayout& L, const PeObjectParams& p, void* packed, cudaStream_t stream) {
    unsigned char* hi = (unsigned char*)packed + L.tc2_base;
    unsigned char* lo = hi + L.tc_bytes_per_pass;
    struct Item { const float* w; const float* b; int N, K_src, K_pad; };
    const Item items[NUM_LAYERS] = {
        {p.backbone_w[0], p.backbone_b[0], 256, 63, 64},   {p.backbone_w[1], p.backbone_b[1], 256, 256, 256},
        {p.backbone_w[2], p.backbone_b[2], 256, 256, 256}, {p.backbone_w[3], p.backbone_b[3], 256, 256, 256},
        {p.backbone_w[4], p.backbone_b[4], 256, 319, 320}, {p.backbone_w[5], p.backbone_b[5], 256, 256, 256},
        {p.backbone_w[6], p.backbone_b[6], 256, 256, 256}, {p.backbone_w[7], p.backbone_b[7], 256, 256, 256},
        {p.head0_w, nullptr, 256, 256, 256},               {p.head3_w, nullptr, 128, 256, 256},
        {p.head6_w, p.head6_b, 192, 128, 128}};
    int64_t off = 0;
    for (int l = 0; l < NUM_LAYERS; ++l) {
        const Item& it = items[l];
        if (!it.w || (l != 8 && l != 9 && !it.b)) { pe_set_error("missing parameter tensor for tensor-core layer %d", l); return PE_ERR_INVALID; }
        pe_tc2_pack_layer_kernel<<<(it.N + 63) / 64, 64, 0, stream>>>(it.w, it.N, it.K_src, it.K_pad, hi + off, lo + off);
        PE_LAUNCH_CHECK("pe_tc2_pack_layer_kernel");
        off += (int64_t)it.N * it.K_pad * 2;
        if (it.b) {
            pe_tc2_pack_bias_kernel<<<(it.N * 8 + 255) / 256, 256, 0, stream>>>(it.b, it.N, hi + off);
            PE_LAUNCH_CHECK("pe_tc2_pack_bias_kernel");
            off += (int64_t)it.N * 16;
        }
    }
    if (off > L.tc_bytes_per_pass) { pe_set_error("internal: CTA-pair weight stream size mismatch"); return PE_ERR_INVALID; }
    return PE_OK;
}

int pe_launch_field_tc2(const PeFieldArgs& args, const PeIntegrated& global_out, int sm_count, cudaStream_t stream) {
    if (!pe_tc_shape_ok(args.ob) || args.training || args.explicit_positions || args.phase != 0) {
        pe_set_error("tensor-core field kernel: unsupported configuration");
        return PE_ERR_UNSUPPORTED;
    }
    const int num_passes = args.precision == PE_PRECISION_FP16X2 ? 2 : 1;
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_field_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_TOTAL));
    const int rpt = TILE_M / args.ob.positions;
    const int64_t tiles = (int64_t)((args.rays + rpt - 1) / rpt) * args.images;
    const int64_t iters = (tiles + 3) / 4;
    if (iters == 0) return PE_OK;
    const int clusters = (int)pe_min64(iters, sm_count / 2);
    const char* dbg_env = getenv("PE_TC_DEBUG");
    pe_field_tc2_kernel<<<2 * clusters, THREADS, SMEM2_TOTAL, stream>>>(args, global_out, num_passes, dbg_env ? atoi(dbg_env) : 0);
    PE_LAUNCH_CHECK("pe_field_tc2_kernel");
    return PE_OK;
}
