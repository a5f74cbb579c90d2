// Token-path GEMM of the mmFormer transformer blocks on tcgen05 (sm_100a): the nn.Linear layers of SelfAttention (qkv, proj) and
// FeedForward (dim -> mlp_dim -> dim), reference models/mmformer.py:192-280, forward, data gradient and weight gradient through
// ONE kernel:
//     D[m][n] = sum_k A(m, k) * B(n, k)  (+ bias[n]),      bf16 operands, fp32 accumulation in TMEM, D bf16 or fp32
// where each operand is either K-major (stored [rows][K], K contiguous) or MN-major (stored [K][rows], rows contiguous):
//     forward          Y  = X  W^T + b :  A = X  [M][K]  K-major,   B = W  [N][K]  K-major
//     data gradient    dX = dY W       :  A = dY [M][N'] K-major,   B = W  [N'][K'] read as [K = N'][rows = K']  MN-major
//     weight gradient  dW = dY^T X     :  A = dY [M][N'] read as [K = M][rows = N'] MN-major,  B = X [M][K'] read as [K = M][rows = K'] MN-major
// so no operand is ever transposed in memory.  128 x 64 output tile per CTA (the token GEMMs are 250..1000 x 512..4096: 16..512
// tiles), K in blocks of 64: both operands arrive by TMA (SWIZZLE_128B; one box [64 K][rows] for a K-major operand, boxes of
// [64 rows][64 K] for an MN-major one; rows / K beyond the matrix are the tensor map's zero fill), eight-slot ring, one
// MMA-issuing thread (M = 128, N = 64, K = 16), epilogue TMEM -> registers -> (+ bias) -> global.  Outputs with few tiles and a long
// K (FFN down: 250 x 512 x 4096) run split-K into a zero-filled fp32 workspace (float4 atomics) + a small finalize launch.
// Warp roles: warps 0-3 epilogue, warp 4 MMA issue + TMEM allocation, warp 5 TMA producer.
// Batched form (pb_gemm_tc_batched; the attention products Q K^T, P V and their four gradients, reference mmformer.py:203-213, one
// problem per (sample, head)): the tensor maps carry the two batch indices as their own dimensions, so operand rows beyond M / N / K
// are zero fill and never the next head's data; blockIdx.z is the (sample, head) pair, no split-K.
#include <cstdlib>
#include <cstring>
#include "tc_common.cuh"

namespace {

constexpr int kGThreads = 192;
constexpr int kGBM = 128, kGBN = 64, kGBK = 64;
constexpr int kGSlots = 3;                       // 3 x (16 + 8) KB = 72 KB: three CTAs per SM (the larger token GEMMs are 2-4 waves of tiles
                                                 // otherwise, each paying the full load latency)
constexpr int kGTileA = kGBM * kGBK * 2;         // 16 KB
constexpr int kGTileB = kGBN * kGBK * 2;         // 8 KB

struct GemmP {
    int M, N, K, ldd;
    int a_kmajor, b_kmajor, d_fp32;
    int splits;                       // split-K: blockIdx.z takes a range of K blocks and adds its partial tile into the fp32 workspace
    int nb1;                          // batched: blockIdx.z = b0 * nb1 + b1 (splits == 1); 0 = not batched
    long long sd0, sd1;               // element strides of D over the two batch indices
};

// operand maps are 4-D: (inner, outer, batch index 1, batch index 0); the unbatched entry point uses batch extents of 1
__device__ __forceinline__ void tma_load_op(uint32_t dst, const CUtensorMap* map, int c0, int c1, int b1, int b0, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(b1), "r"(b0), "r"(smem_u32(bar)) : "memory");
}

// SWIZZLE_128B descriptors (layout type 2).  K-major: 8-row groups of 128-byte rows at SBO = 1024 B (LBO unused); a K = 16 step is
// 32 bytes inside the swizzle atom.  MN-major: ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO)) — 64-row blocks at LBO = 8 KB (the second TMA
// box), 8-k groups at SBO = 1024 B; a K = 16 step is 16 rows of 128 B.
__device__ __forceinline__ uint64_t gemm_desc(uint32_t addr, bool kmajor) {
    return kmajor ? (umma_desc(addr, 16, 1024) | (2ULL << 61)) : (umma_desc(addr, 8192, 1024) | (2ULL << 61));
}

__global__ void __launch_bounds__(kGThreads, 3)
gemm_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, GemmP p, const float* __restrict__ bias,
               void* __restrict__ dout, float* __restrict__ ws, int* err) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + kGSlots * kGTileA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + kGSlots * kGTileB);
    uint64_t* full = bars;
    uint64_t* empty = bars + kGSlots;
    uint64_t* done = empty + kGSlots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * kGBN, m0 = blockIdx.y * kGBM;
    const int kblocks_all = (p.K + kGBK - 1) / kGBK;
    const int zs = p.nb1 ? 0 : (int)blockIdx.z;                                   // split index
    const int bt0 = p.nb1 ? (int)blockIdx.z / p.nb1 : 0, bt1 = p.nb1 ? (int)blockIdx.z % p.nb1 : 0;
    const int kb_lo = (int)((long long)zs * kblocks_all / p.splits), kb_hi = (int)((long long)(zs + 1) * kblocks_all / p.splits);
    const int kblocks = kb_hi - kb_lo;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kGSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 5) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int slot = kb % kGSlots;
                mbar_wait(&empty[slot], ((kb / kGSlots) & 1) ^ 1, err, 61);
                mbar_expect_tx(&full[slot], (uint32_t)(kGTileA + kGTileB));
                const uint32_t ad = smem_u32(a_s) + slot * kGTileA, bd = smem_u32(b_s) + slot * kGTileB;
                const int k0 = (kb_lo + kb) * kGBK;
                if (p.a_kmajor) tma_load_op(ad, &amap, k0, m0, bt1, bt0, &full[slot]);
                else {
                    tma_load_op(ad, &amap, m0, k0, bt1, bt0, &full[slot]);
                    tma_load_op(ad + kGTileA / 2, &amap, m0 + 64, k0, bt1, bt0, &full[slot]);
                }
                if (p.b_kmajor) tma_load_op(bd, &bmap, k0, n0, bt1, bt0, &full[slot]);
                else tma_load_op(bd, &bmap, n0, k0, bt1, bt0, &full[slot]);           // 64 rows: one box
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(kGBM, kGBN) | (p.a_kmajor ? 0u : (1u << 15)) | (p.b_kmajor ? 0u : (1u << 16));
            for (int kb = 0; kb < kblocks; ++kb) {
                const int slot = kb % kGSlots;
                mbar_wait(&full[slot], (kb / kGSlots) & 1, err, 62);
                tc_fence_after();
                const uint64_t a0 = gemm_desc(smem_u32(a_s) + slot * kGTileA, p.a_kmajor != 0);
                const uint64_t b0 = gemm_desc(smem_u32(b_s) + slot * kGTileB, p.b_kmajor != 0);
                const uint32_t astep = p.a_kmajor ? 2u : 128u, bstep = p.b_kmajor ? 2u : 128u;      // 16-byte units per K = 16 step
#pragma unroll
                for (int ks = 0; ks < kGBK / 16; ++ks)
                    umma_f16(tmem_base, a0 + (uint64_t)(astep * ks), b0 + (uint64_t)(bstep * ks), idesc, (kb | ks) ? 1u : 0u);
                umma_commit(&empty[slot]);
            }
            umma_commit(done);
        }
    }
    __syncwarp();
    if (warp < 4) {
        mbar_wait(done, 0, err, 63);
        tc_fence_after();
        const int m = m0 + warp * 32 + lane;                 // accumulator row = TMEM lane
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const long long doff = (long long)bt0 * p.sd0 + (long long)bt1 * p.sd1;      // batch offset of D (elements)
        const bool vec_ok = p.d_fp32 ? (((p.ldd | doff) & 3) == 0) : (((p.ldd | doff) & 7) == 0);
#pragma unroll 1
        for (int c = 0; c < kGBN; c += 16) {
            float v[16];
            tmem_ld16(taddr + c, v);
            const int n = n0 + c;
            if (m < p.M && n < p.N) {
                if (ws) {                                  // split-K partial: the finalize kernel adds the bias and converts
                    float* dst = ws + (size_t)m * p.N + n;
                    if (n + 16 <= p.N && (p.N & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    } else {
                        for (int j = 0; j < 16 && n + j < p.N; ++j) atomicAdd(dst + j, v[j]);
                    }
                    continue;
                }
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (n + j < p.N) v[j] += __ldg(bias + n + j);
                }
                if (p.d_fp32) {
                    float* dst = reinterpret_cast<float*>(dout) + doff + (size_t)m * p.ldd + n;
                    if (n + 16 <= p.N && vec_ok) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
                        for (int j = 0; j < 16 && n + j < p.N; ++j) dst[j] = v[j];
                    }
                } else {
                    bf16* dst = reinterpret_cast<bf16*>(dout) + doff + (size_t)m * p.ldd + n;
                    if (n + 16 <= p.N && vec_ok) {
                        VecIO<bf16, 8>::store(dst, v);
                        VecIO<bf16, 8>::store(dst + 8, v + 8);
                    } else {
                        for (int j = 0; j < 16 && n + j < p.N; ++j) dst[j] = __float2bfloat16_rn(v[j]);
                    }
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

__global__ void gemm_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ bias, void* __restrict__ dout, int M, int N, int ldd,
                                     int d_fp32) {
    const long long total = (long long)M * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i - (long long)m * N);
        const float v = ws[i] + (bias ? __ldg(bias + n) : 0.f);
        if (d_fp32) reinterpret_cast<float*>(dout)[(size_t)m * ldd + n] = v;
        else reinterpret_cast<bf16*>(dout)[(size_t)m * ldd + n] = __float2bfloat16_rn(v);
    }
}

// split-K factor: enough CTAs to occupy the GPU when the output has few tiles and K is long (FFN down: 250 x 512 x 4096 is 8 tiles)
int gemm_splits(int M, int N, int K) {
    { const char* e = getenv("PB_GEMM_SPLITK"); if (e && atoi(e) == 0) return 1; }
    const int tiles = ((M + kGBM - 1) / kGBM) * ((N + kGBN - 1) / kGBN);
    const int kblocks = (K + kGBK - 1) / kGBK;
    if (tiles >= 296 || kblocks < 32) return 1;                 // a CTA keeps only 72 KB in flight: a long K needs many CTAs to fill the HBM pipe
    int s = (444 + tiles - 1) / tiles;
    if (s > kblocks / 4) s = kblocks / 4;          // at least four K blocks per split
    return s < 1 ? 1 : s;
}

// operand stored [b0][b1][outer][inner] with `ld` elements between outer rows and s0 / s1 elements between batch entries (any order in
// memory: a head is a column block of the qkv rows); box = [1][1][box_outer][64 inner]
int make_operand_map(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int box_outer, int nb0 = 1, int nb1 = 1,
                     long long s0 = 0, long long s1 = 0) {
    EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr) return -1;
    const cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)nb1, (cuuint64_t)nb0};
    const cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(nb1 > 1 ? s1 : ld) * 2, (cuuint64_t)(nb0 > 1 ? s0 : ld) * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)box_outer, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}


// ------------------------------------------------------------------------------------ attention rows: scores fused with their softmax
// The score products of SelfAttention (reference mmformer.py:205-208) never leave the SM as fp32: one CTA owns 128 query rows of one
// (sample, head), keeps its A tile (Q, or dO in the backward; [128 rows][d <= 64] K-major) in shared memory and sweeps the key tiles
// (K, or V; [64 keys][d]) through a 4-slot TMA ring; each 128 x 64 product lands in one of two TMEM buffers (4 MMAs, K = 64) and the
// four epilogue warps — thread = accumulator row, so all row statistics are thread-local — consume it while the next one is computed:
//   MODE 0 (forward), two sweeps: (1) running max and sum of exp(scale (s - max)) per row; (2) the products again (recomputing a
//           K = 64 product is cheaper than storing 4 T^2 bytes of fp32 scores), p = exp(scale (s - max)) / sum -> bf16 P and, with
//           dropout, P' = P keep / (1 - p_drop) (same counter hash as attn_softmax_fwd_kernel in attn.cu);
//   MODE 1 (backward), one sweep: dP' = dO V^T per tile, dS = scale (P' .* dP' - P * delta[row]) -> bf16, with
//           delta[row] = sum_j P'_j dP'_j = sum_c dO[row][c] O[row][c] computed beforehand from the outputs (attn_delta_kernel).
// Against the unfused pb_gemm_tc_batched + row-kernel path this removes the write and the read of two [N, H, T, T] fp32 tensors.
constexpr int kARows = 128, kAKeys = 64, kASlots = 4;
constexpr int kATileA = kARows * 128, kATileB = kAKeys * 128;        // 128-byte rows (d <= 64 bf16; the tensor map zero-fills d..63)

struct AttnP {
    int T, ldp, H, ntiles;
    float scale, keep_scale;
    uint32_t thresh;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__global__ void __launch_bounds__(kGThreads, 3)
attn_rows_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, AttnP p, bf16* __restrict__ out0,
                 bf16* __restrict__ out1, const bf16* __restrict__ pin, const bf16* __restrict__ pdin, const float* __restrict__ delta,
                 const long long* __restrict__ seed_ptr, int* err) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + kATileA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + kASlots * kATileB);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;
    uint64_t* b_empty = b_full + kASlots;
    uint64_t* t_full = b_empty + kASlots;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kARows, hh = blockIdx.y, nn = blockIdx.z;
    const int total = (MODE == 0 ? 2 : 1) * p.ntiles;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int i = 0; i < kASlots; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 5) {
        if (lane == 0) {
            mbar_expect_tx(a_full, (uint32_t)kATileA);
            tma_load_op(smem_u32(a_s), &amap, 0, m0, hh, nn, a_full);
            for (int i = 0; i < total; ++i) {
                const int slot = i % kASlots;
                mbar_wait(&b_empty[slot], ((i / kASlots) & 1) ^ 1, err, 64);
                mbar_expect_tx(&b_full[slot], (uint32_t)kATileB);
                tma_load_op(smem_u32(b_s) + slot * kATileB, &bmap, 0, (i % p.ntiles) * kAKeys, hh, nn, &b_full[slot]);
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(kARows, kAKeys);
            mbar_wait(a_full, 0, err, 65);
            const uint64_t a0 = gemm_desc(smem_u32(a_s), true);
            for (int i = 0; i < total; ++i) {
                const int slot = i % kASlots, buf = i & 1;
                mbar_wait(&b_full[slot], (i / kASlots) & 1, err, 66);
                mbar_wait(&t_empty[buf], ((i >> 1) & 1) ^ 1, err, 67);
                tc_fence_after();
                const uint64_t b0 = gemm_desc(smem_u32(b_s) + slot * kATileB, true);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_f16(tmem_base + buf * kAKeys, a0 + (uint64_t)(2 * ks), b0 + (uint64_t)(2 * ks), idesc, ks ? 1u : 0u);
                umma_commit(&b_empty[slot]);
                umma_commit(&t_full[buf]);
            }
        }
    }
    __syncwarp();
    if (warp < 4) {
        const int m = m0 + warp * 32 + lane;                          // accumulator row = TMEM lane = query
        const bool row_ok = m < p.T;
        const long long rowg = ((long long)nn * p.H + hh) * p.T + m;  // row of P / dS / delta
        const float sc2 = p.scale * 1.4426950408889634f;              // exp(scale x) = 2^(sc2 x)
        float nb = 0.f;                                               // -max * sc2 once the maximum is known
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float mx = -INFINITY, l = 0.f, inv = 0.f;
        const float dl = (MODE == 1 && row_ok) ? __ldg(delta + rowg) : 0.f;
        const unsigned long long seed = (MODE == 0 && out1 != nullptr) ? (unsigned long long)*seed_ptr : 0ULL;
#pragma unroll 1
        for (int i = 0; i < total; ++i) {
            const int buf = i & 1;
            mbar_wait(&t_full[buf], (i >> 1) & 1, err, 68);
            tc_fence_after();
            float v[kAKeys];
#pragma unroll
            for (int c = 0; c < kAKeys; c += 16) tmem_ld16(taddr + buf * kAKeys + c, v + c);
            tc_fence_before();
            mbar_arrive(&t_empty[buf]);
            const int j = i % p.ntiles, c0 = j * kAKeys;
            if (MODE == 0) {
                // keys beyond T (last tile only) are -inf scores: they drop out of the maximum and come out as p = 0 without a
                // per-element predicate in the common path
                if (c0 + kAKeys > p.T) {
#pragma unroll
                    for (int c = 0; c < kAKeys; ++c) if (c0 + c >= p.T) v[c] = -INFINITY;
                }
                if (i < p.ntiles) {
                    float tm = v[0];
#pragma unroll
                    for (int c = 1; c < kAKeys; ++c) tm = fmaxf(tm, v[c]);
                    const float mn = fmaxf(mx, tm);
                    const float b2 = -mn * sc2;
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < kAKeys; ++c) acc += ex2_approx(fmaf(v[c], sc2, b2));
                    l = l * ex2_approx(fmaf(mx, sc2, b2)) + acc;
                    mx = mn;
                    if (i == p.ntiles - 1) { inv = 1.f / l; nb = b2; }
                    continue;
                }
            }
            if (!row_ok) continue;
            bf16* o0 = out0 + rowg * p.ldp + c0;
            if (MODE == 0) {
#pragma unroll
                for (int c = 0; c < kAKeys; ++c) v[c] = ex2_approx(fmaf(v[c], sc2, nb)) * inv;
#pragma unroll
                for (int c8 = 0; c8 < kAKeys; c8 += 8) {
                    if (c0 + c8 >= p.ldp) break;                      // ldp is a multiple of 8: a chunk of 8 is inside the row or beyond it
                    VecIO<bf16, 8>::store(o0 + c8, v + c8);
                }
                if (out1 != nullptr) {
                    const unsigned long long e0 = (unsigned long long)rowg * (unsigned long long)p.T + c0;
                    bf16* o1 = out1 + rowg * p.ldp + c0;
#pragma unroll
                    for (int c8 = 0; c8 < kAKeys; c8 += 8) {
                        if (c0 + c8 >= p.ldp) break;
                        float r[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) r[e] = pb_dropout_bits(seed, e0 + c8 + e) >= p.thresh ? v[c8 + e] * p.keep_scale : 0.f;
                        VecIO<bf16, 8>::store(o1 + c8, r);
                    }
                }
            } else {
                // keys beyond T: dP' = 0 (zero-filled V rows) and P = P' = 0 (pad columns), so dS = 0 without a predicate
                const bf16* pi = pin + rowg * p.ldp + c0;
                const bf16* pdi = pdin + rowg * p.ldp + c0;
                const float sdl = p.scale * dl;
#pragma unroll
                for (int c8 = 0; c8 < kAKeys; c8 += 8) {
                    if (c0 + c8 >= p.ldp) break;
                    float r0[8], r1[8];
                    VecIO<bf16, 8>::load(pi + c8, r0);
                    if (pin != pdin) {
                        VecIO<bf16, 8>::load(pdi + c8, r1);
                    } else {                                          // no dropout: P' is P
#pragma unroll
                        for (int e = 0; e < 8; ++e) r1[e] = r0[e];
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) r0[e] = p.scale * r1[e] * v[c8 + e] - sdl * r0[e];
                    VecIO<bf16, 8>::store(o0 + c8, r0);
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

template <int MODE>
int launch_attn_rows(const void* a, const void* b, int N, int H, int T, int d, int lda, long long sa0, long long sa1, int ldb, long long sb0,
                     long long sb1, const AttnP& p, bf16* out0, bf16* out1, const bf16* pin, const bf16* pdin, const float* delta,
                     const long long* seed, int* err, cudaStream_t st) {
    CUtensorMap amap, bmap;
    const int ra = make_operand_map(&amap, a, d, T, lda, kARows, N, H, sa0, sa1);
    const int rb = make_operand_map(&bmap, b, d, T, ldb, kAKeys, N, H, sb0, sb1);
    if (ra || rb) { pb_set_error("attn_rows: cuTensorMapEncodeTiled failed (a %d, b %d)", ra, rb); return PB_EUNSUPPORTED; }
    const size_t smem = (size_t)kATileA + kASlots * kATileB + (1 + 2 * kASlots + 4) * 8 + 16 + 1024;
    auto kern = attn_rows_kernel<MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("attn_rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    kern<<<dim3((T + kARows - 1) / kARows, H, N), kGThreads, smem, st>>>(amap, bmap, p, out0, out1, pin, pdin, delta, seed, err);
    return 0;
}

}  // namespace

extern "C" long long pb_gemm_tc_workspace_floats(int M, int N, int K) {
    return gemm_splits(M, N, K) > 1 ? (long long)M * N : 0;
}

extern "C" int pb_gemm_tc(const void* a, const void* b, const float* bias, void* d, float* workspace, int M, int N, int K, int lda, int ldb,
                          int ldd, int a_kmajor, int b_kmajor, int d_fp32, int* err_flag, pb_stream_t stream) {
    PB_CHECK_ARG(a && b && d && err_flag, "null pointer");
    PB_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, "empty problem");
    PB_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "operand leading dimensions must be multiples of 8 elements (16-byte TMA strides)");
    PB_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "operands must be 16-byte aligned");
    GemmP p;
    p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.a_kmajor = a_kmajor; p.b_kmajor = b_kmajor; p.d_fp32 = d_fp32;
    p.splits = gemm_splits(M, N, K);
    p.nb1 = 0; p.sd0 = p.sd1 = 0;
    PB_CHECK_ARG(p.splits == 1 || workspace, "this shape runs split-K: pass a zero-filled workspace of pb_gemm_tc_workspace_floats(M, N, K) floats");
    CUtensorMap amap, bmap;
    // K-major: [rows][K] -> inner = K, box [128 rows][64 K];  MN-major: [K][rows] -> inner = rows, box [64 K][64 rows]
    const int ra = a_kmajor ? make_operand_map(&amap, a, K, M, lda, kGBM) : make_operand_map(&amap, a, M, K, lda, kGBK);
    const int rb = b_kmajor ? make_operand_map(&bmap, b, K, N, ldb, kGBN) : make_operand_map(&bmap, b, N, K, ldb, kGBK);
    if (ra || rb) { pb_set_error("pb_gemm_tc: cuTensorMapEncodeTiled failed (a %d, b %d; M %d N %d K %d lda %d ldb %d, a_kmajor %d b_kmajor %d, a %p b %p)", ra, rb, M, N, K, lda, ldb, a_kmajor, b_kmajor, a, b); return PB_EUNSUPPORTED; }
    const size_t smem = (size_t)kGSlots * (kGTileA + kGTileB) + (2 * kGSlots + 1) * 8 + 16 + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("pb_gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    const dim3 grid((N + kGBN - 1) / kGBN, (M + kGBM - 1) / kGBM, p.splits);
    float* ws = p.splits > 1 ? workspace : nullptr;
    gemm_tc_kernel<<<grid, kGThreads, smem, (cudaStream_t)stream>>>(amap, bmap, p, bias, d, ws, err_flag);
    PB_CHECK_LAUNCH();
    if (ws) {
        const long long total = (long long)M * N;
        const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
        gemm_finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ws, bias, d, M, N, ldd, d_fp32);
        PB_CHECK_LAUNCH();
    }
    return PB_OK;
}

// one GEMM per (b0, b1): operand / result element strides over the two batch indices are free (a head may be a column block of a row)
extern "C" int pb_gemm_tc_batched(const void* a, const void* b, void* d, int M, int N, int K, int lda, int ldb, int ldd, int a_kmajor, int b_kmajor,
                                  int d_fp32, int nb0, int nb1, long long sa0, long long sa1, long long sb0, long long sb1, long long sd0,
                                  long long sd1, int* err_flag, pb_stream_t stream) {
    PB_CHECK_ARG(a && b && d && err_flag, "null pointer");
    PB_CHECK_ARG(M >= 1 && N >= 1 && K >= 1 && nb0 >= 1 && nb1 >= 1, "empty problem");
    PB_CHECK_ARG((long long)nb0 * nb1 <= 65535, "more than 65535 batch entries");
    PB_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && sa0 % 8 == 0 && sa1 % 8 == 0 && sb0 % 8 == 0 && sb1 % 8 == 0,
                 "operand leading dimensions and batch strides must be multiples of 8 elements (16-byte TMA strides)");
    PB_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "operands must be 16-byte aligned");
    GemmP p;
    p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.a_kmajor = a_kmajor; p.b_kmajor = b_kmajor; p.d_fp32 = d_fp32;
    p.splits = 1; p.nb1 = nb1; p.sd0 = sd0; p.sd1 = sd1;
    CUtensorMap amap, bmap;
    const int ra = a_kmajor ? make_operand_map(&amap, a, K, M, lda, kGBM, nb0, nb1, sa0, sa1) : make_operand_map(&amap, a, M, K, lda, kGBK, nb0, nb1, sa0, sa1);
    const int rb = b_kmajor ? make_operand_map(&bmap, b, K, N, ldb, kGBN, nb0, nb1, sb0, sb1) : make_operand_map(&bmap, b, N, K, ldb, kGBK, nb0, nb1, sb0, sb1);
    if (ra || rb) { pb_set_error("pb_gemm_tc_batched: cuTensorMapEncodeTiled failed (a %d, b %d; M %d N %d K %d lda %d ldb %d batch %d x %d)", ra, rb, M, N, K, lda, ldb, nb0, nb1); return PB_EUNSUPPORTED; }
    const size_t smem = (size_t)kGSlots * (kGTileA + kGTileB) + (2 * kGSlots + 1) * 8 + 16 + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("pb_gemm_tc_batched: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    const dim3 grid((N + kGBN - 1) / kGBN, (M + kGBM - 1) / kGBM, nb0 * nb1);
    gemm_tc_kernel<<<grid, kGThreads, smem, (cudaStream_t)stream>>>(amap, bmap, p, nullptr, d, nullptr, err_flag);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

// P = softmax(scale * Q K^T) (and P' with dropout) without materialising the scores; see attn_rows_kernel.  q / k: [N][T][H][d] views
// (ld = elements between tokens, s0 = between samples, s1 = between heads), d <= 64 and a multiple of 8; p / p_drop [N][H][T][ldp] bf16.
extern "C" int pb_attn_scores_softmax(const void* q, const void* k, void* p, void* p_drop, int N, int H, int T, int d, int ld, long long s0,
                                      long long s1, int ldp, float scale, float drop_p, const long long* seed, int* err_flag,
                                      pb_stream_t stream) {
    PB_CHECK_ARG(q && k && p && err_flag, "null pointer");
    PB_CHECK_ARG(N >= 1 && H >= 1 && T >= 1 && H <= 65535 && N <= 65535, "bad batch");
    PB_CHECK_ARG(d >= 8 && d <= 64 && d % 8 == 0, "head width must be a multiple of 8 up to 64");
    PB_CHECK_ARG(ld % 8 == 0 && s0 % 8 == 0 && s1 % 8 == 0 && ldp % 8 == 0 && ldp >= T && ldp < T + 8, "strides: multiples of 8, ldp = T rounded up to 8");
    PB_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)p & 15) == 0 && ((uintptr_t)p_drop & 15) == 0, "16-byte alignment");
    PB_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && (p_drop == nullptr) == (drop_p == 0.f) && (p_drop == nullptr || seed), "dropout arguments");
    AttnP ap;
    ap.T = T; ap.ldp = ldp; ap.H = H; ap.ntiles = (T + kAKeys - 1) / kAKeys; ap.scale = scale; ap.keep_scale = 1.f / (1.f - drop_p);
    { const float t = drop_p * 4294967296.f; ap.thresh = (uint32_t)(t < 4294967040.f ? t : 4294967040.f); }
    const int rc = launch_attn_rows<0>(q, k, N, H, T, d, ld, s0, s1, ld, s0, s1, ap, (bf16*)p, (bf16*)p_drop, nullptr, nullptr, nullptr, seed,
                                       err_flag, (cudaStream_t)stream);
    if (rc) return rc;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

// dS = scale * (P' .* (dO V^T) - P * delta) without materialising dO V^T.  d_o [N][T][H][d] (ld_o, so0, so1), v likewise (ld_v, sv0, sv1),
// p / p_drop / ds [N][H][T][ldp] bf16 (without dropout pass p for p_drop), delta [N][H][T] fp32 from pb_attn_delta.
extern "C" int pb_attn_dsoftmax(const void* d_o, const void* v, const void* p, const void* p_drop, const float* delta, void* ds, int N, int H,
                                int T, int d, int ld_o, long long so0, long long so1, int ld_v, long long sv0, long long sv1, int ldp,
                                float scale, int* err_flag, pb_stream_t stream) {
    PB_CHECK_ARG(d_o && v && p && p_drop && delta && ds && err_flag, "null pointer");
    PB_CHECK_ARG(N >= 1 && H >= 1 && T >= 1 && H <= 65535 && N <= 65535, "bad batch");
    PB_CHECK_ARG(d >= 8 && d <= 64 && d % 8 == 0, "head width must be a multiple of 8 up to 64");
    PB_CHECK_ARG(ld_o % 8 == 0 && so0 % 8 == 0 && so1 % 8 == 0 && ld_v % 8 == 0 && sv0 % 8 == 0 && sv1 % 8 == 0 && ldp % 8 == 0 && ldp >= T && ldp < T + 8,
                 "strides: multiples of 8, ldp = T rounded up to 8");
    PB_CHECK_ARG(((uintptr_t)d_o & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)p & 15) == 0 && ((uintptr_t)p_drop & 15) == 0 && ((uintptr_t)ds & 15) == 0,
                 "16-byte alignment");
    AttnP ap;
    ap.T = T; ap.ldp = ldp; ap.H = H; ap.ntiles = (T + kAKeys - 1) / kAKeys; ap.scale = scale; ap.keep_scale = 1.f; ap.thresh = 0;
    const int rc = launch_attn_rows<1>(d_o, v, N, H, T, d, ld_o, so0, so1, ld_v, sv0, sv1, ap, (bf16*)ds, nullptr, (const bf16*)p, (const bf16*)p_drop,
                                       delta, nullptr, err_flag, (cudaStream_t)stream);
    if (rc) return rc;
    PB_CHECK_LAUNCH();
    return PB_OK;
}
