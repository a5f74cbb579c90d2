// InstanceNorm3d(affine=False) + LeakyReLU (+ residual): finalize / apply / backward.
// Memory-bound, vectorised (16 B per thread access), fp32 math, float64 cross-block sums.
// Replaces reference models/blocks.py:18 (nn.InstanceNorm3d), :363 (LeakyReLU) and the encoder
// residual adds models/rfnet.py:37,40,43,46.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kDefaultInormOrder = 5;     // measured: 24.45 (0) -> 24.32 ms/step (5); 1, 4, 2, 3, 7 in between

__global__ void finalize_kernel(const double* __restrict__ stats, float* __restrict__ mr, int nc, double inv_v, float eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    double mean = stats[2 * i] * inv_v;
    double var = stats[2 * i + 1] * inv_v - mean * mean;     // biased variance, as InstanceNorm
    if (var < 0.0) var = 0.0;
    mr[2 * i] = (float)mean;
    mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// Block-level finish of per-thread float64 partials (thread = VEC channels of one voxel lane).  When the lanes-per-voxel
// count is a power of two <= 32 the voxel lanes of a warp are combined by xor-shuffles and only the 8 warp totals go
// through shared memory; otherwise every voxel lane writes its partials and 2c threads sum them (the old, slow tail:
// it made the reduce pass 1.7x slower than the apply pass on the same tensor).
template <int VEC>
__device__ __forceinline__ void block_finish(const double* s1, const double* s2, double* ssum, double* __restrict__ out,
                                             int lanes, int vpb, int c, int c0, int vl, bool active) {
    const bool pow2 = lanes <= 32 && (lanes & (lanes - 1)) == 0;
    if (pow2) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            double a = active ? s1[j] : 0.0, b = active ? s2[j] : 0.0;
            for (int off = lanes; off < 32; off <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, off);
                b += __shfl_xor_sync(0xffffffffu, b, off);
            }
            if (lane < lanes) {                               // lanes divides 32: lane % lanes is this thread's channel lane
                ssum[(size_t)wid * 2 * c + 2 * (c0 + j)] = a;
                ssum[(size_t)wid * 2 * c + 2 * (c0 + j) + 1] = b;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * c; i += 256) {
            double t = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) t += ssum[(size_t)w8 * 2 * c + i];
            atomicAdd(&out[i], t);
        }
    } else {
        if (active) {
            double* r = ssum + (size_t)vl * 2 * c;
#pragma unroll
            for (int j = 0; j < VEC; ++j) { r[2 * (c0 + j)] = s1[j]; r[2 * (c0 + j) + 1] = s2[j]; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * c; i += 256) {
            double t = 0.0;
            for (int v = 0; v < vpb; ++v) t += ssum[(size_t)v * 2 * c + i];
            atomicAdd(&out[i], t);
        }
    }
}

// Traversal order of the streaming passes.  Consecutive kernels of the step sweep the same 65-82 MB tensors; a pass that
// walks its tensor in the OPPOSITE direction of the pass before it starts on the data that pass touched last, i.e. on what
// is still resident in the 126 MB L2.  PB_INORM_ORDER is a bit mask: 1 = forward apply reversed, 2 = backward reduce
// reversed, 4 = backward apply reversed (default chosen by measurement).
int inorm_order() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PB_INORM_ORDER");
        v = e ? atoi(e) : kDefaultInormOrder;
    }
    return v;
}

// Voxel walk shared by the passes: grid = (blocks per sample, n); a thread owns VEC channels (c0) and visits the voxels
// first + k * stride, k = 0..K-1 — in increasing order, or (rev) blocks, samples and k all in decreasing order.
struct Walk {
    int n, first, stride, K;                                  // voxels per sample < 2^31 (checked by the host wrappers)
    __device__ __forceinline__ Walk(long long voxels, int vpb, int vl, bool rev) {
        n = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
        const int bx = rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
        stride = (int)gridDim.x * vpb;
        first = bx * vpb + vl;
        K = first < (int)voxels ? ((int)voxels - 1 - first) / stride + 1 : 0;
    }
    // voxel of the it-th visit (it < K)
    __device__ __forceinline__ long long at(int it, bool rev) const { return first + (long long)(rev ? K - 1 - it : it) * stride; }
};

// out = lrelu((y - mean) * rstd) (+ res).  One thread = VEC channels of one voxel lane with its mean / rstd in registers
// (the first version recomputed sample and channel with two 64-bit divisions per vector and fetched the statistics from
// memory per element: 158 instructions per 16-byte vector, issue-bound at 4.0 TB/s); four independent vectors in flight.
template <typename T, int VEC, bool RES>
__global__ void __launch_bounds__(256) apply_fwd_kernel(const T* __restrict__ y, const float* __restrict__ mr,
                                                        const T* __restrict__ res, T* __restrict__ out,
                                                        long long voxels, int c, float slope, int rev,
                                                        const double* __restrict__ stats, float* __restrict__ mr_out, double inv_v, float eps) {
    constexpr int UNR = 4;
    const int lanes = c / VEC, tpb = (256 / lanes) * lanes, vpb = tpb / lanes;
    if ((int)threadIdx.x >= tpb) return;
    const int cl = threadIdx.x % lanes, vl = threadIdx.x / lanes;
    const int c0 = cl * VEC;
    const Walk wk(voxels, vpb, vl, rev != 0);
    float mean[VEC], rstd[VEC];
    if (stats != nullptr) {
        // finalize fused in (same arithmetic as finalize_kernel): mean / rstd straight from the float64 sums of the producing
        // kernel's epilogue; the first block of the sample also publishes them for the backward pass
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const size_t i = (size_t)wk.n * c + c0 + j;
            const double m = stats[2 * i] * inv_v;
            double var = stats[2 * i + 1] * inv_v - m * m;
            if (var < 0.0) var = 0.0;
            mean[j] = (float)m; rstd[j] = (float)(1.0 / sqrt(var + (double)eps));
            if (blockIdx.x == 0 && vl == 0) { mr_out[2 * i] = mean[j]; mr_out[2 * i + 1] = rstd[j]; }
        }
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            mean[j] = mr[((size_t)wk.n * c + c0 + j) * 2]; rstd[j] = mr[((size_t)wk.n * c + c0 + j) * 2 + 1];
        }
    }
    const size_t base = (size_t)wk.n * voxels * c + c0;
    for (int it = 0; it < wk.K; it += UNR) {
        float v[UNR][VEC], r[RES ? UNR : 1][VEC];
        size_t off[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int k = it + u < wk.K ? it + u : it;                // tail: re-load the first vector, its store is skipped
            off[u] = base + (size_t)wk.at(k, rev != 0) * c;
            VecIO<T, VEC>::load(y + off[u], v[u]);
        }
        if (RES) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) VecIO<T, VEC>::load(res + off[u], r[u]);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float xh = (v[u][j] - mean[j]) * rstd[j];
                xh = xh > 0.f ? xh : xh * slope;
                v[u][j] = RES ? xh + r[RES ? u : 0][j] : xh;
            }
            if (it + u < wk.K) VecIO<T, VEC>::store(out + off[u], v[u]);
        }
    }
}

// per-(n,c) sum and sum of squares of an arbitrary tensor (statistics for a PRE-norm block, reference
// models/blocks.py:312-316, where the normalised tensor is not a conv output of ours).  grid = (blocks_per_sample, n)
template <typename T, int VEC>
__global__ void __launch_bounds__(256, sizeof(T) == 2 && VEC < 8 ? 3 : 2) channel_stats_kernel(const T* __restrict__ x, double* __restrict__ stats, long long voxels, int c) {
    // loads in flight per thread: 8 vectors under bf16 storage (128 B; with 4 the pass ran at 1.0 - 1.5 TB/s on the 128^3 / 80^3 levels
    // of the pre-norm backbone), 4 in the fp32 check mode
    constexpr int UNR = sizeof(T) == 2 ? 8 : 4;
    extern __shared__ double ssum[];                          // [vpb][2*c]
    const int n = blockIdx.y;
    const int lanes = c / VEC, tpb = (256 / lanes) * lanes, vpb = tpb / lanes;
    const bool active = (int)threadIdx.x < tpb;
    const int cl = threadIdx.x % lanes, vl = threadIdx.x / lanes, c0 = cl * VEC;
    double s1[VEC], s2[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { s1[j] = 0.0; s2[j] = 0.0; }
    if (active) {
        const T* xn = x + (size_t)n * voxels * c;
        const long long stride = (long long)gridDim.x * vpb;
        for (long long v0 = (long long)blockIdx.x * vpb + vl; v0 < voxels; v0 += stride * UNR) {
            RawVec<T, VEC> raw[UNR];
            bool ok[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const long long v = v0 + u * stride;
                ok[u] = v < voxels;
                if (ok[u]) raw[u].load(xn + v * c + c0);
            }
            if (sizeof(T) == 4) {                             // fp32 check mode: every element in float64
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    if (!ok[u]) continue;
                    float xv[VEC];
                    raw[u].unpack(xv);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) { s1[j] += (double)xv[j]; s2[j] += (double)xv[j] * (double)xv[j]; }
                }
            } else {                                          // bf16 storage: fp32 over a run of UNR, float64 across runs
                float a[VEC], b[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) { a[j] = 0.f; b[j] = 0.f; }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    if (!ok[u]) continue;
                    float xv[VEC];
                    raw[u].unpack(xv);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) { a[j] += xv[j]; b[j] = fmaf(xv[j], xv[j], b[j]); }
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) { s1[j] += (double)a[j]; s2[j] += (double)b[j]; }
            }
        }
    }
    block_finish<VEC>(s1, s2, ssum, stats + (size_t)n * c * 2, lanes, vpb, c, c0, vl, active);
}

// per-(n,c): sum g, sum g*xhat  with g = dout * lrelu'(xhat).  grid = (blocks_per_sample, n);
// a thread keeps the same VEC channels for all its voxels; block reduction in a fixed order (deterministic).
// Precision: the fp32 check mode (T = float) accumulates every element in float64.  Under bf16 storage the inputs carry
// 2^-9 relative rounding already, so runs of kRun voxels are summed in fp32 and only the run totals go to the float64
// accumulators — per-element float64 conversions would otherwise make this memory-bound pass conversion-bound.
constexpr int kRun = 4;

template <typename T, int VEC>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 3 : 2) bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ y,
                                                         const float* __restrict__ mr, double* __restrict__ sums,
                                                         long long voxels, int c, float slope, int rev) {
    constexpr bool kExact = sizeof(T) == 4;
    constexpr int RUN = kRun;
    extern __shared__ double ssum[];                          // [vpb][2*c]
    const int lanes = c / VEC;                                // threads per voxel
    const int tpb = (256 / lanes) * lanes;                    // active threads
    const int vpb = tpb / lanes;                              // voxels per block-iteration
    const bool active = (int)threadIdx.x < tpb;
    const int cl = threadIdx.x % lanes, vl = threadIdx.x / lanes;
    const int c0 = cl * VEC;
    const Walk wk(voxels, vpb, vl, rev != 0);
    const int n = wk.n;
    double sg[VEC], sgx[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { sg[j] = 0.0; sgx[j] = 0.0; }
    if (active) {
        float mean[VEC], rstd[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            mean[j] = mr[((size_t)n * c + c0 + j) * 2]; rstd[j] = mr[((size_t)n * c + c0 + j) * 2 + 1];
        }
        const T* dn = dout + (size_t)n * voxels * c + c0;
        const T* yn = y + (size_t)n * voxels * c + c0;
        float fg[VEC], fgx[VEC];                               // bf16 mode: a thread sums a few dozen voxels, fp32 is ample
#pragma unroll
        for (int j = 0; j < VEC; ++j) { fg[j] = 0.f; fgx[j] = 0.f; }
        for (int it = 0; it < wk.K; it += RUN) {
            RawVec<T, VEC> rg[RUN], ry[RUN];
#pragma unroll
            for (int u = 0; u < RUN; ++u) {                  // all loads of the run are in flight together, still packed
                const long long v = wk.at(it + u < wk.K ? it + u : it, rev != 0);   // tail: a valid address, weight 0 below
                rg[u].load(dn + v * c);
                ry[u].load(yn + v * c);
            }
#pragma unroll
            for (int u = 0; u < RUN; ++u) {
                float g[VEC], yv[VEC];
                rg[u].unpack(g); ry[u].unpack(yv);
                const float live = it + u < wk.K ? 1.f : 0.f;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float xh = (yv[j] - mean[j]) * rstd[j];
                    const float gg = (xh > 0.f ? g[j] : g[j] * slope) * live;
                    if (kExact) {
                        sg[j] += (double)gg;
                        sgx[j] += (double)gg * (((double)yv[j] - (double)mean[j]) * (double)rstd[j]);
                    } else {
                        fg[j] += gg;
                        fgx[j] = fmaf(gg, xh, fgx[j]);
                    }
                }
            }
        }
        if (!kExact) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) { sg[j] += (double)fg[j]; sgx[j] += (double)fgx[j]; }
        }
    }
    block_finish<VEC>(sg, sgx, ssum, sums + (size_t)n * c * 2, lanes, vpb, c, c0, vl, active);
}

// dy = rstd * (g - mean(g) - xhat * mean(g*xhat)).  The three-term difference cancels heavily when the incoming
// gradient lies mostly in span{1, xhat} (typical right below the loss); in the fp32 check mode it is evaluated in
// float64 from the float64 sums.  Under bf16 storage fp32 arithmetic is already ~2^15 finer than the operands.
// Same thread -> channel mapping as the reduce pass, so the per-channel coefficients are loaded once per thread.
template <typename T, int VEC>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 3 : 1) bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ y,
                                                        const float* __restrict__ mr, const double* __restrict__ sums,
                                                        T* __restrict__ dy, long long voxels, int c, double inv_v, float slope,
                                                        int rev) {
    constexpr bool kExact = sizeof(T) == 4;
    const int lanes = c / VEC, tpb = (256 / lanes) * lanes, vpb = tpb / lanes;
    if ((int)threadIdx.x >= tpb) return;
    const int cl = threadIdx.x % lanes, vl = threadIdx.x / lanes;
    const int c0 = cl * VEC;
    const Walk wk(voxels, vpb, vl, rev != 0);
    const int n = wk.n;
    float mean[VEC], rstd[VEC], af[VEC], bf[VEC];
    double ad[VEC], bd[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        mean[j] = mr[((size_t)n * c + c0 + j) * 2]; rstd[j] = mr[((size_t)n * c + c0 + j) * 2 + 1];
        ad[j] = sums[((size_t)n * c + c0 + j) * 2] * inv_v; bd[j] = sums[((size_t)n * c + c0 + j) * 2 + 1] * inv_v;
        af[j] = (float)ad[j]; bf[j] = (float)bd[j];
    }
    const size_t base = (size_t)n * voxels * c + c0;
    constexpr int RUN = sizeof(T) == 2 ? 3 : 2;                // packed bf16: three vector pairs in flight per thread, 3 CTAs per SM
    for (int it = 0; it < wk.K; it += RUN) {
        RawVec<T, VEC> rg[RUN], ry[RUN];
        size_t off[RUN];
#pragma unroll
        for (int u = 0; u < RUN; ++u) {
            off[u] = base + (size_t)wk.at(it + u < wk.K ? it + u : it, rev != 0) * c;      // tail: re-load, store skipped
            rg[u].load(dout + off[u]);
            ry[u].load(y + off[u]);
        }
#pragma unroll
        for (int u = 0; u < RUN; ++u) {
            float g[VEC], yv[VEC];
            rg[u].unpack(g); ry[u].unpack(yv);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float xh = (yv[j] - mean[j]) * rstd[j];
                const float gg = xh > 0.f ? g[j] : g[j] * slope;
                if (kExact) {
                    const double xd = ((double)yv[j] - (double)mean[j]) * (double)rstd[j];
                    g[j] = (float)((double)rstd[j] * ((double)gg - ad[j] - xd * bd[j]));
                } else {
                    g[j] = rstd[j] * (gg - af[j] - xh * bf[j]);
                }
            }
            if (it + u < wk.K) VecIO<T, VEC>::store(dy + off[u], g);
        }
    }
}

template <typename T, int VEC>
int run_fwd(const void* y, const float* mr, const void* res, void* out, int n, long long voxels, int c, float slope, cudaStream_t st,
            const double* stats = nullptr, float* mr_out = nullptr, float eps = 0.f) {
    const int lanes = c / VEC, vpb = 256 / lanes;
    int bps = (int)((voxels + (long long)vpb * 8 - 1) / ((long long)vpb * 8));       // >= 8 voxels per thread
    const int cap = (148 * 8 + n - 1) / n;
    if (bps > cap) bps = cap;
    if (bps < 1) bps = 1;
    if (res)
        apply_fwd_kernel<T, VEC, true><<<dim3(bps, n), 256, 0, st>>>((const T*)y, mr, (const T*)res, (T*)out, voxels, c, slope,
                                                                     inorm_order() & 1, stats, mr_out, 1.0 / (double)voxels, eps);
    else
        apply_fwd_kernel<T, VEC, false><<<dim3(bps, n), 256, 0, st>>>((const T*)y, mr, nullptr, (T*)out, voxels, c, slope,
                                                                      inorm_order() & 1, stats, mr_out, 1.0 / (double)voxels, eps);
    return 0;
}

template <typename T, int VEC>
int run_bwd(const void* dout, const void* y, const float* mr, double* sums, void* dy, int n, long long voxels, int c, float slope,
            cudaStream_t st) {
    const int lanes = c / VEC;
    const int vpb = 256 / lanes;
    // Large volumes go sample by sample: the apply pass then re-reads dout / y of the sample the reduce pass has just
    // streamed, which is still resident in the 126 MB L2 (2 x voxels x c x sizeof(T) per sample).
    const long long sample_bytes = voxels * c * (long long)sizeof(T);
    // (measured on B200: per-sample launches cost more than the L2 re-read saves, 3.9 -> 5.5 ms/step; kept batched)
    const int group = (false && sample_bytes >= (4LL << 20)) ? 1 : n;
    for (int n0 = 0; n0 < n; n0 += group) {
        const int nn = group;
        const size_t eoff = (size_t)n0 * voxels * c;
        const T* d_ = (const T*)dout + eoff;
        const T* y_ = (const T*)y + eoff;
        T* o_ = (T*)dy + eoff;
        const float* mr_ = mr + (size_t)n0 * c * 2;
        double* s_ = sums + (size_t)n0 * c * 2;
        // grids are capped at whole waves of resident CTAs (the reduce pass ran 1184 CTAs on 444 slots: 2.67 waves)
        const size_t rsmem = (size_t)vpb * 2 * c * sizeof(double);
        static int occ_r = 0, occ_a = 0;                       // per template instantiation
        if (!occ_r && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, bwd_reduce_kernel<T, VEC>, 256, rsmem) != cudaSuccess || occ_r < 1)) occ_r = 2;
        if (!occ_a && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, bwd_apply_kernel<T, VEC>, 256, 0) != cudaSuccess || occ_a < 1)) occ_a = 2;
        int bps = (int)((voxels + (long long)vpb * 8 - 1) / ((long long)vpb * 8));   // >= 8 voxel-iterations per thread
        const int cap = 148 * occ_r / nn;                      // one wave
        if (bps > cap) bps = cap;
        if (bps < 1) bps = 1;
        bwd_reduce_kernel<T, VEC><<<dim3(bps, nn), 256, rsmem, st>>>(d_, y_, mr_, s_, voxels, c, slope, inorm_order() & 2);
        pb_count_launch();
        int bpa = (int)((voxels + (long long)vpb * 4 - 1) / ((long long)vpb * 4));     // >= 4 voxels per thread
        const int capa = 148 * occ_a * 2 / nn;                 // two waves
        if (bpa > capa) bpa = capa;
        if (bpa < 1) bpa = 1;
        bwd_apply_kernel<T, VEC><<<dim3(bpa, nn), 256, 0, st>>>(d_, y_, mr_, s_, o_, voxels, c, 1.0 / (double)voxels, slope,
                                                              inorm_order() & 4);
        if (n0 + group < n) pb_count_launch();
    }
    return 0;
}

#define VEC_SWITCH(T, c, CALL)                                   \
    switch (pb_vec_width(c)) {                                   \
        case 8: CALL(T, 8); break;                               \
        case 4: CALL(T, 4); break;                               \
        case 2: CALL(T, 2); break;                               \
        default: CALL(T, 1); break;                              \
    }

}  // namespace

extern "C" int pb_inorm_finalize(const double* stats, float* mr, int n, int c, long long voxels, float eps, pb_stream_t stream) {
    PB_CHECK_ARG(stats && mr && n > 0 && c > 0 && voxels > 0, "bad argument");
    const int nc = n * c;
    finalize_kernel<<<(nc + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, mr, nc, 1.0 / (double)voxels, eps);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_channel_stats(int dtype, const void* x, double* stats, int n, long long voxels, int c, pb_stream_t stream) {
    PB_CHECK_ARG(x && stats && n > 0 && c > 0 && voxels > 0, "bad argument");
    PB_CHECK_ARG(c <= 256 * pb_vec_width(c), "too many channels");
    cudaStream_t st = (cudaStream_t)stream;
    const int vw = pb_vec_width(c);
    const int lanes = c / vw, vpb = 256 / lanes;
    int bps = (int)((voxels + (long long)vpb * 8 - 1) / ((long long)vpb * 8));
    const int cap = (148 * 8 + n - 1) / n;
    if (bps > cap) bps = cap;
    if (bps < 1) bps = 1;
    const size_t smem = (size_t)vpb * 2 * c * sizeof(double);
#define CS_CALL(T, V) channel_stats_kernel<T, V><<<dim3(bps, n), 256, smem, st>>>((const T*)x, stats, voxels, c)
    if (dtype == PB_BF16) { VEC_SWITCH(bf16, c, CS_CALL) } else { VEC_SWITCH(float, c, CS_CALL) }
#undef CS_CALL
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_inorm_lrelu_fwd(int dtype, const void* y, const float* mr, const void* res, void* out, int n,
                                  long long voxels, int c, float slope, pb_stream_t stream) {
    PB_CHECK_ARG(y && mr && out && n > 0 && c > 0 && voxels > 0 && voxels < (1LL << 31), "bad argument");
    PB_CHECK_ARG(c <= 256 * pb_vec_width(c), "too many channels");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_F(T, V) run_fwd<T, V>(y, mr, res, out, n, voxels, c, slope, st)
    if (dtype == PB_BF16) { VEC_SWITCH(bf16, c, CALL_F) } else { VEC_SWITCH(float, c, CALL_F) }
#undef CALL_F
    PB_CHECK_LAUNCH();
    return PB_OK;
}

// pb_inorm_finalize + pb_inorm_lrelu_fwd in one launch: mean / rstd are computed from the float64 sums in the kernel prologue
// and written to `mr` (for the backward pass) by the first block of each sample.
extern "C" int pb_inorm_lrelu_fwd_stats(int dtype, const void* y, const double* stats, float* mr, const void* res, void* out, int n,
                                        long long voxels, int c, float eps, float slope, pb_stream_t stream) {
    PB_CHECK_ARG(y && stats && mr && out && n > 0 && c > 0 && voxels > 0 && voxels < (1LL << 31), "bad argument");
    PB_CHECK_ARG(c <= 256 * pb_vec_width(c), "too many channels");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_FS(T, V) run_fwd<T, V>(y, nullptr, res, out, n, voxels, c, slope, st, stats, mr, eps)
    if (dtype == PB_BF16) { VEC_SWITCH(bf16, c, CALL_FS) } else { VEC_SWITCH(float, c, CALL_FS) }
#undef CALL_FS
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_inorm_lrelu_bwd(int dtype, const void* dout, const void* y, const float* mr, double* sums, void* dy, int n,
                                  long long voxels, int c, float slope, pb_stream_t stream) {
    PB_CHECK_ARG(dout && y && mr && sums && dy && n > 0 && c > 0 && voxels > 0 && voxels < (1LL << 31), "bad argument");
    PB_CHECK_ARG(c <= 256 * pb_vec_width(c), "too many channels");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_B(T, V) run_bwd<T, V>(dout, y, mr, sums, dy, n, voxels, c, slope, st)
    if (dtype == PB_BF16) { VEC_SWITCH(bf16, c, CALL_B) } else { VEC_SWITCH(float, c, CALL_B) }
#undef CALL_B
    PB_CHECK_LAUNCH();
    return PB_OK;
}
