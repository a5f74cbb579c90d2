// Row kernels of the mmFormer token path (reference models/mmformer.py:192-313) that sit between the tcgen05 GEMMs of gemm_tc.cu:
//   * softmax over the keys of the fp32 scores S = Q K^T (mmformer.py:206-208: scale, softmax, attention dropout), written as the bf16
//     probabilities P (kept for the backward) and, with dropout, the bf16 P' = P * keep / (1 - p) that multiplies V;
//   * its backward, dS = scale * (P' .* dP' - P * sum_j(P'_j dP'_j)) from the fp32 dP' = dO V^T (dropout folded in: P' .* dP' = P .* dP);
//   * LayerNorm over the embedding (mmformer.py:233-250: PreNorm / PreNormDrop), forward and backward.
// One warp per row; rows are 125..2048 scores (8 KB at most: the three passes of the softmax hit L1) or 512 channels.
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256)
attn_softmax_fwd_kernel(const float* __restrict__ s, bf16* __restrict__ p, bf16* __restrict__ pd, long long rows, int T, int ldp, float scale,
                        float drop_p, const long long* __restrict__ seed_ptr) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* sr = s + row * T;
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, sr[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) sum += __expf((sr[j] - mx) * scale);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    bf16* pr = p + row * ldp;
    bf16* pdr = pd ? pd + row * ldp : nullptr;
    const unsigned long long seed = pdr ? (unsigned long long)*seed_ptr : 0ULL;
    const uint32_t thresh = (uint32_t)fminf(drop_p * 4294967296.f, 4294967040.f);       // dropped when hash < thresh
    const float keep_scale = 1.f / (1.f - drop_p);
    for (int j = lane; j < ldp; j += 32) {
        const float v = j < T ? __expf((sr[j] - mx) * scale) * inv : 0.f;
        pr[j] = __float2bfloat16_rn(v);
        if (pdr) {
            const bool keep = pb_dropout_bits(seed, (unsigned long long)row * (unsigned long long)T + j) >= thresh;
            pdr[j] = __float2bfloat16_rn(keep ? v * keep_scale : 0.f);
        }
    }
}

__global__ void __launch_bounds__(256)
attn_softmax_bwd_kernel(const float* __restrict__ dp, const bf16* __restrict__ p, const bf16* __restrict__ pd, bf16* __restrict__ ds, long long rows,
                        int T, int ldp, float scale) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* dr = dp + row * T;
    const bf16* pr = p + row * ldp;
    const bf16* pdr = pd + row * ldp;
    float dot = 0.f;
    for (int j = lane; j < T; j += 32) dot += __bfloat162float(pdr[j]) * dr[j];
    dot = warp_sum(dot);
    bf16* dsr = ds + row * ldp;
    for (int j = lane; j < ldp; j += 32) {
        const float v = j < T ? scale * (__bfloat162float(pdr[j]) * dr[j] - __bfloat162float(pr[j]) * dot) : 0.f;
        dsr[j] = __float2bfloat16_rn(v);
    }
}

// delta[n][h][t] = sum_c dO[n][t][h][c] * O[n][t][h][c]  (= sum_j P'_j dP'_j, the row term of the softmax backward)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ d_o, const bf16* __restrict__ o, float* __restrict__ delta, int N, int T, int H, int d) {
    const long long total = (long long)N * T * H;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;             // (n, t, h), h fastest: neighbours read neighbours
    if (i >= total) return;
    const int h = (int)(i % H);
    const long long nt = i / H;
    const int t = (int)(nt % T), n = (int)(nt / T);
    const bf16* a = d_o + i * d;
    const bf16* b = o + i * d;
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
        float x[8], y[8];
        VecIO<bf16, 8>::load(a + c, x);
        VecIO<bf16, 8>::load(b + c, y);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += x[e] * y[e];
    }
    delta[((long long)n * H + h) * T + t] = acc;
}

// ---- LayerNorm: C = 256 * NV channels, a lane holds NV vectors of 8 ----
template <typename T, int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, T* __restrict__ y, float* __restrict__ mean,
                     float* __restrict__ rstd, long long rows, float eps) {
    constexpr int C = 256 * NV;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        VecIO<T, 8>::load(x + row * C + i * 256 + lane * 8, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
    const float mu = warp_sum(sum) * (1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mu; sq += d * d; }
    const float rs = rsqrtf(warp_sum(sq) * (1.f / C) + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float wv[8], bv[8], o[8];
        VecIO<float, 8>::load(w + i * 256 + lane * 8, wv);
        VecIO<float, 8>::load(b + i * 256 + lane * 8, bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * wv[j] + bv[j];
        VecIO<T, 8>::store(y + row * C + i * 256 + lane * 8, o);
    }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = dy * w;   dw += dy * xhat, db += dy  (per-CTA partial sums, then atomics)
template <typename T, int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ w, T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long rows) {
    constexpr int C = 256 * NV;
    __shared__ float sw[C], sb[C];
    for (int c = threadIdx.x; c < C; c += blockDim.x) { sw[c] = 0.f; sb[c] = 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    float aw[NV][8], ab[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { aw[i][j] = 0.f; ab[i][j] = 0.f; }
    for (long long row = (long long)blockIdx.x * warps + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * warps) {
        const float mu = mean[row], rs = rstd[row];
        float g[NV][8], xh[NV][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float dv[8], xv[8], wv[8];
            VecIO<T, 8>::load(dy + row * C + i * 256 + lane * 8, dv);
            VecIO<T, 8>::load(x + row * C + i * 256 + lane * 8, xv);
            VecIO<float, 8>::load(w + i * 256 + lane * 8, wv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xh[i][j] = (xv[j] - mu) * rs;
                g[i][j] = dv[j] * wv[j];
                s1 += g[i][j];
                s2 += g[i][j] * xh[i][j];
                aw[i][j] += dv[j] * xh[i][j];
                ab[i][j] += dv[j];
            }
        }
        s1 = warp_sum(s1) * (1.f / C);
        s2 = warp_sum(s2) * (1.f / C);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rs * (g[i][j] - s1 - xh[i][j] * s2);
            VecIO<T, 8>::store(dx + row * C + i * 256 + lane * 8, o);
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sw[i * 256 + lane * 8 + j], aw[i][j]);
            atomicAdd(&sb[i * 256 + lane * 8 + j], ab[i][j]);
        }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) { atomicAdd(dw + c, sw[c]); atomicAdd(db + c, sb[c]); }
}

template <typename T>
int ln_fwd_dispatch(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, long long rows, int C, float eps,
                    cudaStream_t st) {
    const int blocks = (int)((rows + 7) / 8);
#define PB_LN_F(NV) layernorm_fwd_kernel<T, NV><<<blocks, 256, 0, st>>>((const T*)x, w, b, (T*)y, mean, rstd, rows, eps)
    switch (C / 256) {
        case 1: PB_LN_F(1); break;
        case 2: PB_LN_F(2); break;
        case 4: PB_LN_F(4); break;
        default: return 1;
    }
#undef PB_LN_F
    return 0;
}

template <typename T>
int ln_bwd_dispatch(const void* dy, const void* x, const float* mean, const float* rstd, const float* w, void* dx, float* dw, float* db,
                    long long rows, int C, cudaStream_t st) {
    long long want = (rows + 31) / 32;                      // >= 4 rows per warp before the partial sums go out as atomics
    const int blocks = (int)(want < 1 ? 1 : want > 296 ? 296 : want);
#define PB_LN_B(NV) layernorm_bwd_kernel<T, NV><<<blocks, 256, 0, st>>>((const T*)dy, (const T*)x, mean, rstd, w, (T*)dx, dw, db, rows)
    switch (C / 256) {
        case 1: PB_LN_B(1); break;
        case 2: PB_LN_B(2); break;
        case 4: PB_LN_B(4); break;
        default: return 1;
    }
#undef PB_LN_B
    return 0;
}

}  // namespace

extern "C" int pb_attn_softmax_fwd(const float* s, void* p, void* p_drop, long long rows, int T, int ldp, float scale, float drop_p,
                                   const long long* seed, pb_stream_t stream) {
    PB_CHECK_ARG(s && p, "null pointer");
    PB_CHECK_ARG(rows >= 1 && T >= 1 && ldp >= T, "bad shape");
    PB_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "dropout probability outside [0, 1)");
    PB_CHECK_ARG((p_drop == nullptr) == (drop_p == 0.f), "p_drop is given exactly when drop_p > 0");
    PB_CHECK_ARG(p_drop == nullptr || seed != nullptr, "dropout needs a device-side seed");
    PB_CHECK_ARG((rows + 7) / 8 <= 2147483647LL, "too many rows");
    attn_softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(s, (bf16*)p, (bf16*)p_drop, rows, T, ldp, scale, drop_p, seed);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_attn_softmax_bwd(const float* dp, const void* p, const void* p_drop, void* ds, long long rows, int T, int ldp, float scale,
                                   pb_stream_t stream) {
    PB_CHECK_ARG(dp && p && p_drop && ds, "null pointer (without dropout pass p as p_drop)");
    PB_CHECK_ARG(rows >= 1 && T >= 1 && ldp >= T, "bad shape");
    PB_CHECK_ARG((rows + 7) / 8 <= 2147483647LL, "too many rows");
    attn_softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dp, (const bf16*)p, (const bf16*)p_drop, (bf16*)ds, rows, T, ldp, scale);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_layernorm_fwd(int dtype, const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, long long rows, int C,
                                float eps, pb_stream_t stream) {
    PB_CHECK_ARG(x && w && b && y && mean && rstd, "null pointer");
    PB_CHECK_ARG(rows >= 1, "no rows");
    PB_CHECK_ARG(C == 256 || C == 512 || C == 1024, "embedding width must be 256, 512 or 1024");
    PB_CHECK_ARG(dtype == PB_F32 || dtype == PB_BF16, "dtype");
    const int r = dtype == PB_F32 ? ln_fwd_dispatch<float>(x, w, b, y, mean, rstd, rows, C, eps, (cudaStream_t)stream)
                                  : ln_fwd_dispatch<bf16>(x, w, b, y, mean, rstd, rows, C, eps, (cudaStream_t)stream);
    PB_CHECK_ARG(r == 0, "unsupported width");
    PB_CHECK_LAUNCH();
    return PB_OK;
}

// dw / db [C] fp32 are ACCUMULATED into (zero them first)
extern "C" int pb_layernorm_bwd(int dtype, const void* dy, const void* x, const float* mean, const float* rstd, const float* w, void* dx, float* dw,
                                float* db, long long rows, int C, pb_stream_t stream) {
    PB_CHECK_ARG(dy && x && mean && rstd && w && dx && dw && db, "null pointer");
    PB_CHECK_ARG(rows >= 1, "no rows");
    PB_CHECK_ARG(C == 256 || C == 512 || C == 1024, "embedding width must be 256, 512 or 1024");
    PB_CHECK_ARG(dtype == PB_F32 || dtype == PB_BF16, "dtype");
    const int r = dtype == PB_F32 ? ln_bwd_dispatch<float>(dy, x, mean, rstd, w, dx, dw, db, rows, C, (cudaStream_t)stream)
                                  : ln_bwd_dispatch<bf16>(dy, x, mean, rstd, w, dx, dw, db, rows, C, (cudaStream_t)stream);
    PB_CHECK_ARG(r == 0, "unsupported width");
    PB_CHECK_LAUNCH();
    return PB_OK;
}

// delta [N][H][T] fp32 from the contiguous bf16 d_o and o [N][T][H][d] (d a multiple of 8): input of pb_attn_dsoftmax
extern "C" int pb_attn_delta(const void* d_o, const void* o, float* delta, int N, int T, int H, int d, pb_stream_t stream) {
    PB_CHECK_ARG(d_o && o && delta, "null pointer");
    PB_CHECK_ARG(N >= 1 && T >= 1 && H >= 1 && d >= 8 && d % 8 == 0, "bad shape");
    const long long total = (long long)N * T * H;
    PB_CHECK_ARG((total + 255) / 256 <= 2147483647LL, "too many rows");
    attn_delta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)d_o, (const bf16*)o, delta, N, T, H, d);
    PB_CHECK_LAUNCH();
    return PB_OK;
}
