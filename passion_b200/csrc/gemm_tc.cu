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
