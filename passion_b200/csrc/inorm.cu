// InstanceNorm3d(affine=False) + LeakyReLU (+ residual): finalize / apply / backward.
// Memory-bound, vectorised (16 B per thread access), fp32 math, float64 cross-block sums.
// Replaces reference models/blocks.py:18 (nn.InstanceNorm3d), :363 (LeakyReLU) and the encoder
// residual adds models/rfnet.py:37,40,43,46.
#include "common.cuh"

namespace {

__global__ void finalize_kernel(const double* __restrict__ stats, float* __restrict__ mr, int nc, double inv_v, float eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    double mean = stats[2 * i] * inv_v;
    double var = stats[2 * i + 1] * inv_v - mean * mean;     // biased variance, as InstanceNorm
    if (var < 0.0) var = 0.0;
    mr[2 * i] = (float)mean;
    mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// out = lrelu((y - mean) * rstd) (+ res);  one thread = VEC channels of one voxel, two independent vectors per
// iteration (all loads are issued before the first use, doubling the bytes in flight per thread)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) apply_fwd_kernel(const T* __restrict__ y, const float* __restrict__ mr,
                                                        const T* __restrict__ res, T* __restrict__ out,
                                                        long long total_vec, long long vox_c, int c, float slope) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += 2 * stride) {
        const long long i2 = i + stride;
        const bool has2 = i2 < total_vec;
        const long long e = i * VEC, e2 = (has2 ? i2 : i) * VEC;
        float v[VEC], r[VEC], v2[VEC], r2[VEC];
        VecIO<T, VEC>::load(y + e, v);
        VecIO<T, VEC>::load(y + e2, v2);
        if (res) { VecIO<T, VEC>::load(res + e, r); VecIO<T, VEC>::load(res + e2, r2); }
        const float* m = mr + ((size_t)(e / vox_c) * c + (int)(e % c)) * 2;
        const float* m2 = mr + ((size_t)(e2 / vox_c) * c + (int)(e2 % c)) * 2;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float xh = (v[j] - m[2 * j]) * m[2 * j + 1];
            xh = xh > 0.f ? xh : xh * slope;
            v[j] = res ? xh + r[j] : xh;
            float xg = (v2[j] - m2[2 * j]) * m2[2 * j + 1];
            xg = xg > 0.f ? xg : xg * slope;
            v2[j] = res ? xg + r2[j] : xg;
        }
        VecIO<T, VEC>::store(out + e, v);
        if (has2) VecIO<T, VEC>::store(out + e2, v2);
    }
}

// per-(n,c) sum and sum of squares of an arbitrary tensor (statistics for a PRE-norm block, reference
// models/blocks.py:312-316, where the normalised tensor is not a conv output of ours).  grid = (blocks_per_sample, n)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) channel_stats_kernel(const T* __restrict__ x, double* __restrict__ stats, long long voxels, int c) {
    extern __shared__ double ssum[];                          // [vpb][2*c]
    const int n = blockIdx.y;
    const int lanes = c / VEC, tpb = (256 / lanes) * lanes, vpb = tpb / lanes;
    if ((int)threadIdx.x < tpb) {
        const int cl = threadIdx.x % lanes, vl = threadIdx.x / lanes, c0 = cl * VEC;
        double s1[VEC], s2[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s1[j] = 0.0; s2[j] = 0.0; }
        const T* xn = x + (size_t)n * voxels * c;
        for (long long v = (long long)blockIdx.x * vpb + vl; v < voxels; v += (long long)gridDim.x * vpb) {
            float xv[VEC];
            VecIO<T, VEC>::load(xn + v * c + c0, xv);
#pragma unroll
            for (int j = 0; j < VEC; ++j) { s1[j] += (double)xv[j]; s2[j] += (double)xv[j] * (double)xv[j]; }
        }
        double* r = ssum + (size_t)vl * 2 * c;
#pragma unroll
        for (int j = 0; j < VEC; ++j) { r[2 * (c0 + j)] = s1[j]; r[2 * (c0 + j) + 1] = s2[j]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * c; i += 256) {
        double s = 0.0;
        for (int v = 0; v < vpb; ++v) s += ssum[(size_t)v * 2 * c + i];
        atomicAdd(&stats[(size_t)n * c * 2 + i], s);
    }
}

// per-(n,c): sum g, sum g*xhat  with g = dout * lrelu'(xhat).  grid = (blocks_per_sample, n);
// a thread keeps the same VEC channels for all its voxels; block reduction in a fixed order (deterministic).
template <typename T, int VEC>
__global__ void __launch_bounds__(256) bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ y,
                                                         const float* __restrict__ mr, double* __restrict__ sums,
                                                         long long voxels, int c, float slope) {
    extern __shared__ double ssum[];                          // [vpb][2*c]
    const int n = blockIdx.y;
    const int lanes = c / VEC;                                // threads per voxel
    const int tpb = (256 / lanes) * lanes;                    // active threads
    const int vpb = tpb / lanes;                              // voxels per block-iteration
    if ((int)threadIdx.x < tpb) {
        const int cl = threadIdx.x % lanes, vl = threadIdx.x / lanes;
        const int c0 = cl * VEC;
        float mean[VEC], rstd[VEC];
        double sg[VEC], sgx[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            mean[j] = mr[((size_t)n * c + c0 + j) * 2]; rstd[j] = mr[((size_t)n * c + c0 + j) * 2 + 1];
            sg[j] = 0.0; sgx[j] = 0.0;
        }
        const T* dn = dout + (size_t)n * voxels * c;
        const T* yn = y + (size_t)n * voxels * c;
        for (long long v = (long long)blockIdx.x * vpb + vl; v < voxels; v += (long long)gridDim.x * vpb) {
            float g[VEC], yv[VEC];
            VecIO<T, VEC>::load(dn + v * c + c0, g);
            VecIO<T, VEC>::load(yn + v * c + c0, yv);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float xh = (yv[j] - mean[j]) * rstd[j];
                float gg = xh > 0.f ? g[j] : g[j] * slope;
                sg[j] += (double)gg; sgx[j] += (double)gg * (((double)yv[j] - (double)mean[j]) * (double)rstd[j]);
            }
        }
        double* r = ssum + (size_t)vl * 2 * c;
#pragma unroll
        for (int j = 0; j < VEC; ++j) { r[2 * (c0 + j)] = sg[j]; r[2 * (c0 + j) + 1] = sgx[j]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * c; i += 256) {
        double s = 0.0;
        for (int v = 0; v < vpb; ++v) s += ssum[(size_t)v * 2 * c + i];
        atomicAdd(&sums[(size_t)n * c * 2 + i], s);
    }
}

// dy = rstd * (g - mean(g) - xhat * mean(g*xhat)).  The three-term difference cancels heavily when the incoming
// gradient lies mostly in span{1, xhat} (typical right below the loss), so it is evaluated in float64 from the
// float64 sums; the LeakyReLU branch uses the same fp32 xhat as the forward pass.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ y,
                                                        const float* __restrict__ mr, const double* __restrict__ sums,
                                                        T* __restrict__ dy, long long total_vec, long long vox_c, int c,
                                                        double inv_v, float slope) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * VEC;
        const int n = (int)(e / vox_c);
        const int c0 = (int)(e % c);
        float g[VEC], yv[VEC];
        VecIO<T, VEC>::load(dout + e, g);
        VecIO<T, VEC>::load(y + e, yv);
        const float* m = mr + ((size_t)n * c + c0) * 2;
        const double* s = sums + ((size_t)n * c + c0) * 2;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float rstd = m[2 * j + 1];
            const float xh = (yv[j] - m[2 * j]) * rstd;
            const float gg = xh > 0.f ? g[j] : g[j] * slope;
            const double xd = ((double)yv[j] - (double)m[2 * j]) * (double)rstd;
            g[j] = (float)((double)rstd * ((double)gg - s[2 * j] * inv_v - xd * (s[2 * j + 1] * inv_v)));
        }
        VecIO<T, VEC>::store(dy + e, g);
    }
}

int grid_for(long long work, int per_block) {
    long long b = (work + per_block - 1) / per_block;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T, int VEC>
int run_fwd(const void* y, const float* mr, const void* res, void* out, int n, long long voxels, int c, float slope, cudaStream_t st) {
    const long long total_vec = (long long)n * voxels * c / VEC;
    apply_fwd_kernel<T, VEC><<<grid_for(total_vec, 256), 256, 0, st>>>((const T*)y, mr, (const T*)res, (T*)out, total_vec,
                                                                       voxels * c, c, slope);
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
        int bps = (int)((voxels + (long long)vpb * 8 - 1) / ((long long)vpb * 8));   // >= 8 voxel-iterations per thread
        const int cap = (148 * 8 + nn - 1) / nn;
        if (bps > cap) bps = cap;
        if (bps < 1) bps = 1;
        bwd_reduce_kernel<T, VEC><<<dim3(bps, nn), 256, (size_t)vpb * 2 * c * sizeof(double), st>>>(d_, y_, mr_, s_, voxels, c, slope);
        pb_count_launch();
        const long long total_vec = (long long)nn * voxels * c / VEC;
        bwd_apply_kernel<T, VEC><<<grid_for(total_vec, 256), 256, 0, st>>>(d_, y_, mr_, s_, o_, total_vec, voxels * c, c,
                                                                          1.0 / (double)voxels, slope);
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
    PB_CHECK_ARG(y && mr && out && n > 0 && c > 0 && voxels > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_F(T, V) run_fwd<T, V>(y, mr, res, out, n, voxels, c, slope, st)
    if (dtype == PB_BF16) { VEC_SWITCH(bf16, c, CALL_F) } else { VEC_SWITCH(float, c, CALL_F) }
#undef CALL_F
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_inorm_lrelu_bwd(int dtype, const void* dout, const void* y, const float* mr, double* sums, void* dy, int n,
                                  long long voxels, int c, float slope, pb_stream_t stream) {
    PB_CHECK_ARG(dout && y && mr && sums && dy && n > 0 && c > 0 && voxels > 0, "bad argument");
    PB_CHECK_ARG(c <= 256 * pb_vec_width(c), "too many channels");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_B(T, V) run_bwd<T, V>(dout, y, mr, sums, dy, n, voxels, c, slope, st)
    if (dtype == PB_BF16) { VEC_SWITCH(bf16, c, CALL_B) } else { VEC_SWITCH(float, c, CALL_B) }
#undef CALL_B
    PB_CHECK_LAUNCH();
    return PB_OK;
}
