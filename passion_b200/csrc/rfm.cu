// Region-aware modal fusion, the part that is not a convolution:
//   pooled region statistics, gated modality mixing, and their adjoints.
// Replaces reference models/blocks.py:504-517 (modal_fusion.forward) and :597-616
// (region_aware_modal_fusion.forward) without materialising the [B,K,cls,C,D,H,W]
// broadcast product: with y[k] the (masked) modality features and p_i the class probabilities,
//     feat_avg_i[k,c] = mean_v(y[k,c] p_i) / (mean_v p_i + 1e-7)        -> pb_rfm_pool (the two sums)
//     region_i[c]     = p_i * sum_k gate_i[k] y[k,c]                     -> pb_rfm_mix
// The 4C+1 -> 128 -> 4 gate MLP on the pooled vector is host-side glue on a [B, 4C+1] tensor.
#include <cstring>
#include "common.cuh"

namespace {

// S[n][i][kc] += sum_v y[kc] p_i ; Psum[n][i] += sum_v p_i.   grid = (blocks_per_sample, n)
// Block reduction is order-deterministic (per-thread partials in smem, summed in a fixed order), so two
// samples with identical inputs produce bit-identical block partials.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) pool_kernel(const T* __restrict__ y, const float* __restrict__ p, double* __restrict__ S,
                                                   double* __restrict__ Psum, long long voxels, int kc) {
    extern __shared__ float ssum[];                           // [vpb][4*kc + 4]
    const int n = blockIdx.y;
    const int lanes = kc / VEC, tpb = (256 / lanes) * lanes, vpb = tpb / lanes;
    const int row = 4 * kc + 4;
    if ((int)threadIdx.x < tpb) {
        const int cl = threadIdx.x % lanes, vl = threadIdx.x / lanes, c0 = cl * VEC;
        float acc[4][VEC], pacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[i][j] = 0.f;
        const T* yn = y + (size_t)n * voxels * kc + c0;
        const float4* pn = reinterpret_cast<const float4*>(p) + (size_t)n * voxels;
        for (long long v = (long long)blockIdx.x * vpb + vl; v < voxels; v += (long long)gridDim.x * vpb) {
            float yv[VEC];
            VecIO<T, VEC>::load(yn + v * kc, yv);
            const float4 pv = __ldg(pn + v);
            const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                pacc[i] += pp[i];
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[i][j] = fmaf(yv[j], pp[i], acc[i][j]);
            }
        }
        float* r = ssum + (size_t)vl * row;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) r[i * kc + c0 + j] = acc[i][j];
            if (cl == 0) r[4 * kc + i] = pacc[i];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < row; i += 256) {
        float s = 0.f;
        for (int v = 0; v < vpb; ++v) s += ssum[(size_t)v * row + i];
        if (i < 4 * kc) atomicAdd(&S[(size_t)n * 4 * kc + i], (double)s);
        else            atomicAdd(&Psum[(size_t)n * 4 + (i - 4 * kc)], (double)s);
    }
}

// R[n][v][i*C+c] = p_i * sum_k gate[n][i][k] y[n][v][k*C+c]   (K = 4 modalities, 4 classes)
template <typename T, int VEC, int K>
__global__ void __launch_bounds__(256) mix_kernel(const T* __restrict__ y, const float* __restrict__ p, const float* __restrict__ gate,
                                                  T* __restrict__ r, long long voxels, int c, long long total) {
    const int cv = c / VEC;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int cl = (int)(t % cv);
        const long long nv = t / cv;                          // n*voxels + v
        const int n = (int)(nv / voxels);
        const float4 pv = __ldg(reinterpret_cast<const float4*>(p) + nv);
        const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
        const float* g = gate + (size_t)n * 4 * K;
        float yv[K][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k) VecIO<T, VEC>::load(y + nv * K * c + k * c + cl * VEC, yv[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float o[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) s = fmaf(__ldg(g + i * K + k), yv[k][j], s);
                o[j] = s * pp[i];
            }
            VecIO<T, VEC>::store(r + nv * 4 * c + i * c + cl * VEC, o);
        }
    }
}

// dgate[n][i][k] += sum_{v,c} p_i y[k*C+c] dR[i*C+c].   grid = (blocks_per_sample, n)
template <typename T, int VEC, int K>
__global__ void __launch_bounds__(256) mix_bwd_gate_kernel(const T* __restrict__ y, const float* __restrict__ p, const T* __restrict__ dr,
                                                           double* __restrict__ dgate, long long voxels, int c) {
    __shared__ float red[8][4 * K];
    const int n = blockIdx.y, cv = c / VEC;
    float acc[4 * K];
#pragma unroll
    for (int i = 0; i < 4 * K; ++i) acc[i] = 0.f;
    const long long total = voxels * cv;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const int cl = (int)(t % cv);
        const long long nv = (size_t)n * voxels + t / cv;
        const float4 pv = __ldg(reinterpret_cast<const float4*>(p) + nv);
        const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
        float yv[K][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k) VecIO<T, VEC>::load(y + nv * K * c + k * c + cl * VEC, yv[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float g[VEC];
            VecIO<T, VEC>::load(dr + nv * 4 * c + i * c + cl * VEC, g);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < VEC; ++j) s = fmaf(yv[k][j], g[j], s);
                acc[i * K + k] = fmaf(s, pp[i], acc[i * K + k]);
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4 * K; ++i) {
        const float s = warp_sum(acc[i]);
        if (lane == 0) red[wid][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 4 * K) {
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
        atomicAdd(&dgate[(size_t)n * 4 * K + threadIdx.x], (double)s);
    }
}

// dy[n][v][k*C+c] = sum_i p_i (gate[n][i][k] dR[n][v][i*C+c] + dS[n][i][k*C+c]).  grid = (blocks_per_sample, n)
template <typename T, int VEC, int K>
__global__ void __launch_bounds__(256) bwd_y_kernel(const float* __restrict__ p, const float* __restrict__ gate, const T* __restrict__ dr,
                                                    const float* __restrict__ dS, T* __restrict__ dy, long long voxels, int c) {
    extern __shared__ __align__(16) float sds[];              // dS[n] : [4][K*c], then gate[4*K]
    const int n = blockIdx.y, cv = c / VEC, kc = K * c;
    for (int i = threadIdx.x; i < 4 * kc; i += 256) sds[i] = dS[(size_t)n * 4 * kc + i];
    float* sg = sds + 4 * kc;
    if (threadIdx.x < 4 * K) sg[threadIdx.x] = gate[(size_t)n * 4 * K + threadIdx.x];
    __syncthreads();
    const long long total = voxels * cv;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const int cl = (int)(t % cv);
        const long long nv = (size_t)n * voxels + t / cv;
        const float4 pv = __ldg(reinterpret_cast<const float4*>(p) + nv);
        const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
        float g[4][VEC];
#pragma unroll
        for (int i = 0; i < 4; ++i) VecIO<T, VEC>::load(dr + nv * 4 * c + i * c + cl * VEC, g[i]);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float o[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) s = fmaf(pp[i], fmaf(sg[i * K + k], g[i][j], sds[i * kc + k * c + cl * VEC + j]), s);
                o[j] = s;
            }
            VecIO<T, VEC>::store(dy + nv * kc + k * c + cl * VEC, o);
        }
    }
}

// ------------------------------------------------------------------------------------ gate MLP of modal_fusion, fused
// Reference models/blocks.py:507-513 per class i: feat = [ mean_v(y p_i) / (mean_v p_i + 1e-7) (K*C values), mean_v p_i + 1e-7 ]
//   -> Conv1x1(K*C+1 -> 128) -> LeakyReLU(0.2) -> Conv1x1(128 -> 4) -> sigmoid = gate_i[k].
// One block per (sample, class): thread h owns hidden unit h.  For the single-modality passes (K = 1) the pooled vector has C
// values, sitting in slot m = sample / b of the 4C-wide MLP input (zeros elsewhere), and only gate column m is used.
// The host-side version of this was ~12 tiny tensor-op launches forward and ~40 backward per call (6 calls per step).
struct GateW {
    const float* w0[4]; const float* b0[4]; const float* w2[4]; const float* b2[4];     // per class: [128][4C+1], [128], [4][128], [4]
    float* dw0[4]; float* db0[4]; float* dw2[4]; float* db2[4];                           // gradient accumulators (zero-filled), backward only
};

constexpr int kGateH = 128;

__global__ void __launch_bounds__(kGateH) rfm_gate_fwd_kernel(GateW gw, const double* __restrict__ S, const double* __restrict__ Ps,
                                                             float* __restrict__ z1buf, float* __restrict__ gate, int b, double inv_v,
                                                             int K, int c) {
    __shared__ float feat[4 * 64 + 1 + 3];
    __shared__ float red[4][kGateH / 32];
    const int n = blockIdx.x, i = blockIdx.y, h = threadIdx.x;
    const int kc = K * c, F = 4 * c + 1;
    const int slot0 = K == 4 ? 0 : (n / b) * c;                 // first MLP input column fed by S
    const float pavg = (float)(Ps[(size_t)n * 4 + i] * inv_v) + 1e-7f;
    for (int f = h; f < kc; f += kGateH) feat[f] = (float)(S[((size_t)n * 4 + i) * kc + f] * inv_v) / pavg;
    __syncthreads();
    const float* w = gw.w0[i] + (size_t)h * F;
    float z = gw.b0[i][h] + w[F - 1] * pavg;
    for (int f = 0; f < kc; ++f) z = fmaf(feat[f], w[slot0 + f], z);
    z1buf[((size_t)n * 4 + i) * kGateH + h] = z;
    const float a = z > 0.f ? z : 0.2f * z;
    float part[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) part[k] = warp_sum(a * gw.w2[i][k * kGateH + h]);
    if ((h & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[k][h >> 5] = part[k];
    }
    __syncthreads();
    if (h < 4) {
        const float zz = gw.b2[i][h] + red[h][0] + red[h][1] + red[h][2] + red[h][3];
        const float g = 1.f / (1.f + expf(-zz));
        if (K == 4) gate[((size_t)n * 4 + i) * 4 + h] = g;
        else if (h == n / b) gate[(size_t)n * 4 + i] = g;
    }
}

__global__ void __launch_bounds__(kGateH) rfm_gate_bwd_kernel(GateW gw, const double* __restrict__ S, const double* __restrict__ Ps,
                                                             const float* __restrict__ z1buf, const double* __restrict__ dgate,
                                                             float* __restrict__ dS, int b, double inv_v, int K, int c) {
    __shared__ float feat[4 * 64 + 1 + 3];
    __shared__ float dz1s[kGateH];
    __shared__ float dz2[4];
    __shared__ float red[4][kGateH / 32];
    const int n = blockIdx.x, i = blockIdx.y, h = threadIdx.x;
    const int kc = K * c, F = 4 * c + 1;
    const int slot0 = K == 4 ? 0 : (n / b) * c;
    const float pavg = (float)(Ps[(size_t)n * 4 + i] * inv_v) + 1e-7f;
    for (int f = h; f < kc; f += kGateH) feat[f] = (float)(S[((size_t)n * 4 + i) * kc + f] * inv_v) / pavg;
    const float z = z1buf[((size_t)n * 4 + i) * kGateH + h];
    const float a = z > 0.f ? z : 0.2f * z;
    // recompute the gate pre-activations (4 block reductions) to get sigmoid'
    float part[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) part[k] = warp_sum(a * gw.w2[i][k * kGateH + h]);
    if ((h & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[k][h >> 5] = part[k];
    }
    __syncthreads();
    if (h < 4) {
        const float zz = gw.b2[i][h] + red[h][0] + red[h][1] + red[h][2] + red[h][3];
        const float g = 1.f / (1.f + expf(-zz));
        float dg;
        if (K == 4) dg = (float)dgate[((size_t)n * 4 + i) * 4 + h];
        else dg = h == n / b ? (float)dgate[(size_t)n * 4 + i] : 0.f;
        const float d = dg * g * (1.f - g);
        dz2[h] = d;
        if (d != 0.f) atomicAdd(gw.db2[i] + h, d);
    }
    __syncthreads();
    float dh = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dh = fmaf(dz2[k], gw.w2[i][k * kGateH + h], dh);
        if (dz2[k] != 0.f) atomicAdd(gw.dw2[i] + k * kGateH + h, dz2[k] * a);
    }
    const float dz1 = dh * (z > 0.f ? 1.f : 0.2f);
    dz1s[h] = dz1;
    atomicAdd(gw.db0[i] + h, dz1);
    float* dw = gw.dw0[i] + (size_t)h * F;
    atomicAdd(dw + F - 1, dz1 * pavg);
    for (int f = 0; f < kc; ++f) atomicAdd(dw + slot0 + f, dz1 * feat[f]);
    __syncthreads();
    // d feat[f] = sum_h dz1[h] w0[h][slot0 + f]  ->  dS = d feat / (voxels * pavg)    (p, hence pavg, is detached)
    for (int f = h; f < kc; f += kGateH) {
        float acc = 0.f;
        const float* w = gw.w0[i] + slot0 + f;
        for (int hh = 0; hh < kGateH; ++hh) acc = fmaf(dz1s[hh], w[(size_t)hh * F], acc);
        dS[((size_t)n * 4 + i) * kc + f] = acc * (float)inv_v / pavg;
    }
}

int blocks_per_sample(long long work_items, int n) {
    long long b = (work_items + 256 * 4 - 1) / (256 * 4);
    const long long cap = (148LL * 8 + n - 1) / n;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

#define RFM_DISPATCH(c_expr, ...)                                                        \
    do {                                                                                  \
        const int vw_ = pb_vec_width(c_expr);                                             \
        if (dtype == PB_BF16) {                                                           \
            typedef bf16 T;                                                               \
            if (vw_ == 8) { constexpr int VEC = 8; __VA_ARGS__ } else if (vw_ == 4) { constexpr int VEC = 4; __VA_ARGS__ } \
            else if (vw_ == 2) { constexpr int VEC = 2; __VA_ARGS__ } else { constexpr int VEC = 1; __VA_ARGS__ }          \
        } else {                                                                          \
            typedef float T;                                                              \
            if (vw_ == 8) { constexpr int VEC = 8; __VA_ARGS__ } else if (vw_ == 4) { constexpr int VEC = 4; __VA_ARGS__ } \
            else if (vw_ == 2) { constexpr int VEC = 2; __VA_ARGS__ } else { constexpr int VEC = 1; __VA_ARGS__ }          \
        }                                                                                 \
    } while (0)

extern "C" int pb_rfm_pool(int dtype, const void* y, const float* p, double* S, double* Psum, int n, long long voxels, int kc,
                           pb_stream_t stream) {
    PB_CHECK_ARG(y && p && S && Psum && n > 0 && voxels > 0 && kc > 0, "bad argument");
    PB_CHECK_ARG(kc / pb_vec_width(kc) <= 256, "too many channels");
    cudaStream_t st = (cudaStream_t)stream;
    RFM_DISPATCH(kc, {
        const int lanes = kc / VEC, vpb = 256 / lanes;
        int bps = blocks_per_sample(voxels * lanes / 2, n);
        if ((long long)bps * vpb > voxels) bps = (int)((voxels + vpb - 1) / vpb);
        pool_kernel<T, VEC><<<dim3(bps, n), 256, (size_t)vpb * (4 * kc + 4) * sizeof(float), st>>>((const T*)y, p, S, Psum, voxels, kc);
    });
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_rfm_mix(int dtype, const void* y, const float* p, const float* gate, void* r, int n, long long voxels, int k,
                          int c, pb_stream_t stream) {
    PB_CHECK_ARG(y && p && gate && r && n > 0 && voxels > 0 && (k == 4 || k == 1) && c > 0, "bad argument (k must be 1 or 4)");
    cudaStream_t st = (cudaStream_t)stream;
    RFM_DISPATCH(c, {
        const long long total = (long long)n * voxels * (c / VEC);
        long long blocks = (total + 255) / 256;
        if (blocks > 148LL * 16) blocks = 148LL * 16;
        if (k == 4) mix_kernel<T, VEC, 4><<<(int)blocks, 256, 0, st>>>((const T*)y, p, gate, (T*)r, voxels, c, total);
        else        mix_kernel<T, VEC, 1><<<(int)blocks, 256, 0, st>>>((const T*)y, p, gate, (T*)r, voxels, c, total);
    });
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_rfm_mix_bwd_gate(int dtype, const void* y, const float* p, const void* dr, double* dgate, int n,
                                   long long voxels, int k, int c, pb_stream_t stream) {
    PB_CHECK_ARG(y && p && dr && dgate && n > 0 && voxels > 0 && (k == 4 || k == 1) && c > 0, "bad argument (k must be 1 or 4)");
    cudaStream_t st = (cudaStream_t)stream;
    RFM_DISPATCH(c, {
        const int bps = blocks_per_sample(voxels * (c / VEC), n);
        if (k == 4) mix_bwd_gate_kernel<T, VEC, 4><<<dim3(bps, n), 256, 0, st>>>((const T*)y, p, (const T*)dr, dgate, voxels, c);
        else        mix_bwd_gate_kernel<T, VEC, 1><<<dim3(bps, n), 256, 0, st>>>((const T*)y, p, (const T*)dr, dgate, voxels, c);
    });
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_rfm_bwd_y(int dtype, const float* p, const float* gate, const void* dr, const float* dS, void* dy, int n,
                            long long voxels, int k, int c, pb_stream_t stream) {
    PB_CHECK_ARG(p && gate && dr && dS && dy && n > 0 && voxels > 0 && (k == 4 || k == 1) && c > 0, "bad argument (k must be 1 or 4)");
    cudaStream_t st = (cudaStream_t)stream;
    RFM_DISPATCH(c, {
        const int bps = blocks_per_sample(voxels * (c / VEC), n);
        const size_t sm = (size_t)(4 * k * c + 16) * sizeof(float);
        if (k == 4) bwd_y_kernel<T, VEC, 4><<<dim3(bps, n), 256, sm, st>>>(p, gate, (const T*)dr, dS, (T*)dy, voxels, c);
        else        bwd_y_kernel<T, VEC, 1><<<dim3(bps, n), 256, sm, st>>>(p, gate, (const T*)dr, dS, (T*)dy, voxels, c);
    });
    PB_CHECK_LAUNCH();
    return PB_OK;
}


// Gate MLP of the four modal_fusion modules (blocks.py:495-513) on the pooled sums of pb_rfm_pool: S [n][4][K*c], Psum [n][4]
// (float64) -> gate [n][4][K] and the hidden pre-activations z1 [n][4][128] (kept for the backward pass).  `w` = 16 parameter
// pointers (4 classes x {w0 [128][4c+1], b0 [128], w2 [4][128], b2 [4]}).  K = 1: sample n holds modality n / b only.
extern "C" int pb_rfm_gate_fwd(const float* const* w, const double* S, const double* Psum, float* z1, float* gate, int n, int b,
                               long long voxels, int k, int c, pb_stream_t stream) {
    PB_CHECK_ARG(w && S && Psum && z1 && gate && n > 0 && b > 0 && voxels > 0 && (k == 4 || k == 1) && c > 0 && c <= 64, "bad argument");
    GateW gw;
    memset(&gw, 0, sizeof(gw));
    for (int i = 0; i < 4; ++i) { gw.w0[i] = w[i]; gw.b0[i] = w[4 + i]; gw.w2[i] = w[8 + i]; gw.b2[i] = w[12 + i]; }
    rfm_gate_fwd_kernel<<<dim3(n, 4), kGateH, 0, (cudaStream_t)stream>>>(gw, S, Psum, z1, gate, b, 1.0 / (double)voxels, k, c);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

// Adjoint: dgate [n][4][K] (float64, from pb_rfm_mix_bwd_gate) -> dS [n][4][K*c] fp32 and the parameter gradients, ACCUMULATED
// into the 16 zero-filled buffers `dw` (same order as `w`).
extern "C" int pb_rfm_gate_bwd(const float* const* w, float* const* dw, const double* S, const double* Psum, const float* z1,
                               const double* dgate, float* dS, int n, int b, long long voxels, int k, int c, pb_stream_t stream) {
    PB_CHECK_ARG(w && dw && S && Psum && z1 && dgate && dS && n > 0 && b > 0 && voxels > 0 && (k == 4 || k == 1) && c > 0 && c <= 64,
                 "bad argument");
    GateW gw;
    for (int i = 0; i < 4; ++i) {
        gw.w0[i] = w[i]; gw.b0[i] = w[4 + i]; gw.w2[i] = w[8 + i]; gw.b2[i] = w[12 + i];
        gw.dw0[i] = dw[i]; gw.db0[i] = dw[4 + i]; gw.dw2[i] = dw[8 + i]; gw.db2[i] = dw[12 + i];
    }
    rfm_gate_bwd_kernel<<<dim3(n, 4), kGateH, 0, (cudaStream_t)stream>>>(gw, S, Psum, z1, dgate, dS, b, 1.0 / (double)voxels, k, c);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

// ------------------------------------------------------------------------------------ masked modality stacking
// MaskModal of the reference (models/rfnet.py:154-163, 239-242; mmformer.py:316-326) for P decoder passes at once:
//   out[p*B + b][v][m*C + c] = enc[m*B + b][v][c] * ms[p][b][m]
// enc is the modality-major encoder output of the grouped pass; each element is read once and written P times.
namespace {

template <typename T, int VEC>
__global__ void __launch_bounds__(256) masked_stack_fwd_kernel(const T* __restrict__ enc, const float* __restrict__ ms,
                                                               T* __restrict__ out, int P, int B, long long V, int C) {
    const int cl = C / VEC;
    const long long total = 4LL * B * V * cl;
    for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        // channel chunk fastest, then the modality: the threads of a warp cover whole 4C-channel rows of the stacked tensor
        // (full sectors on the 5x larger side of this op); the per-modality reads stay 16 B x consecutive voxels per warp
        const int c = (int)(t % cl) * VEC;
        long long r = t / cl;
        const int m = (int)(r & 3); r >>= 2;
        const long long v = r % V;
        const int b = (int)(r / V);
        float x[VEC], y[VEC];
        VecIO<T, VEC>::load(enc + (((size_t)m * B + b) * V + v) * C + c, x);
        for (int p = 0; p < P; ++p) {
            const float s = __ldg(ms + ((size_t)p * B + b) * 4 + m);
#pragma unroll
            for (int i = 0; i < VEC; ++i) y[i] = x[i] * s;
            VecIO<T, VEC>::store(out + (((size_t)p * B + b) * V + v) * (4 * C) + m * C + c, y);
        }
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) masked_stack_bwd_kernel(const T* __restrict__ dout, const float* __restrict__ ms,
                                                               T* __restrict__ denc, int P, int B, long long V, int C) {
    const int cl = C / VEC;
    const long long total = 4LL * B * V * cl;
    for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        // channel chunk fastest, then the modality: the threads of a warp cover whole 4C-channel rows of the stacked tensor
        // (full sectors on the 5x larger side of this op); the per-modality reads stay 16 B x consecutive voxels per warp
        const int c = (int)(t % cl) * VEC;
        long long r = t / cl;
        const int m = (int)(r & 3); r >>= 2;
        const long long v = r % V;
        const int b = (int)(r / V);
        float acc[VEC], g[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        for (int p = 0; p < P; ++p) {
            const float s = __ldg(ms + ((size_t)p * B + b) * 4 + m);
            if (s == 0.f) continue;
            VecIO<T, VEC>::load(dout + (((size_t)p * B + b) * V + v) * (4 * C) + m * C + c, g);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = fmaf(g[i], s, acc[i]);
        }
        VecIO<T, VEC>::store(denc + (((size_t)m * B + b) * V + v) * C + c, acc);
    }
}

template <bool BWD>
int masked_stack_launch(int dtype, const void* in, const float* ms, void* out, int P, int B, long long V, int C, cudaStream_t st) {
    const int vec = dtype == PB_BF16 ? 8 : 4;
    if (C % vec) { pb_set_error("masked_stack: C %d not a multiple of %d", C, vec); return PB_EUNSUPPORTED; }
    const long long total = 4LL * B * V * (C / vec);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (dtype == PB_BF16) {
        if (BWD) masked_stack_bwd_kernel<bf16, 8><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)in, ms, (bf16*)out, P, B, V, C);
        else masked_stack_fwd_kernel<bf16, 8><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)in, ms, (bf16*)out, P, B, V, C);
    } else {
        if (BWD) masked_stack_bwd_kernel<float, 4><<<(unsigned)blocks, 256, 0, st>>>((const float*)in, ms, (float*)out, P, B, V, C);
        else masked_stack_fwd_kernel<float, 4><<<(unsigned)blocks, 256, 0, st>>>((const float*)in, ms, (float*)out, P, B, V, C);
    }
    return 0;
}

}  // namespace

extern "C" int pb_masked_stack_fwd(int dtype, const void* enc, const float* ms, void* out, int passes, int b, long long voxels,
                                   int c, pb_stream_t stream) {
    PB_CHECK_ARG(enc && ms && out && passes >= 1 && b >= 1 && voxels >= 1 && c >= 1, "bad arguments");
    if (int e = masked_stack_launch<false>(dtype, enc, ms, out, passes, b, voxels, c, (cudaStream_t)stream)) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_masked_stack_bwd(int dtype, const void* dout, const float* ms, void* denc, int passes, int b, long long voxels,
                                   int c, pb_stream_t stream) {
    PB_CHECK_ARG(dout && ms && denc && passes >= 1 && b >= 1 && voxels >= 1 && c >= 1, "bad arguments");
    if (int e = masked_stack_launch<true>(dtype, dout, ms, denc, passes, b, voxels, c, (cudaStream_t)stream)) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}
