// PASSION objective kernels (reference utils/criterions.py), channels-last, fp32 math, float64 cross-block sums.
//   softmax4      : F.softmax(logit / T, dim=1) over the 4 classes                           (:93-94, rfnet.py:286,379)
//   cedice        : the three per-class sums behind dice_loss_bs (:25-38) and softmax_weighted_loss_bs (:59-76)
//   kl            : sum_v,c pt (log pt - log ps) with both clamped to [0.005, 1]                (:98-101)
//   proto_*       : masked class-mean prototypes, cosine-similarity maps, (s-t)^2 and |s-t|      (:144-180)
// All of them read class LABELS (uint8) instead of the float64 one-hot target; prediction sample n uses the
// labels of sample n % b (batched decoder passes).  Reductions: per-thread registers -> warp shuffle -> smem ->
// one float64 atomic per block and quantity.
#include "common.cuh"

namespace {

constexpr float kClampMin = 0.005f;

template <typename T> __device__ __forceinline__ void load4(const T* p, float* o) { VecIO<T, 4>::load(p, o); }

template <typename T>
__global__ void __launch_bounds__(256) softmax4_kernel(const T* __restrict__ logits, float* __restrict__ probs, long long rows, float inv_temp) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < rows; i += (long long)gridDim.x * 256) {
        float v[4];
        load4(logits + i * 4, v);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] *= inv_temp;
        const float m = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) { v[c] = expf(v[c] - m); s += v[c]; }
        const float r = 1.f / s;
        reinterpret_cast<float4*>(probs)[i] = make_float4(v[0] * r, v[1] * r, v[2] * r, v[3] * r);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) softmax4_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ dprobs,
                                                           T* __restrict__ dlogits, long long rows, float inv_temp) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < rows; i += (long long)gridDim.x * 256) {
        const float4 p = reinterpret_cast<const float4*>(probs)[i], g = reinterpret_cast<const float4*>(dprobs)[i];
        const float dot = p.x * g.x + p.y * g.y + p.z * g.z + p.w * g.w;
        float o[4] = {inv_temp * p.x * (g.x - dot), inv_temp * p.y * (g.y - dot), inv_temp * p.z * (g.z - dot), inv_temp * p.w * (g.w - dot)};
        VecIO<T, 4>::store(dlogits + i * 4, o);
    }
}

// block-wide sum of NV per-thread values, result valid in threads < NV of warp 0 ... returned through smem `red`
template <int NV>
__device__ __forceinline__ void block_reduce_atomic(float* vals, double* dst) {
    __shared__ float red[8][NV];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float s = warp_sum(vals[i]);
        if (lane == 0) red[wid][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        atomicAdd(dst + threadIdx.x, (double)s);
    }
}

// sums[n][0..3] = A_c = sum p_c t_c ; [4..7] = L_c = sum p_c ; [8..11] = E_c = sum t_c log(clamp(p_c))
__global__ void __launch_bounds__(256) cedice_fwd_kernel(const float* __restrict__ probs, const uint8_t* __restrict__ labels,
                                                         double* __restrict__ sums, long long voxels, int b) {
    const int n = blockIdx.y;
    const float4* pn = reinterpret_cast<const float4*>(probs) + (size_t)n * voxels;
    const uint8_t* ln = labels + (size_t)(n % b) * voxels;
    float acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.f;
    for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < voxels; v += (long long)gridDim.x * 256) {
        const float4 p4 = __ldg(pn + v);
        const float p[4] = {p4.x, p4.y, p4.z, p4.w};
        const int t = ln[v];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[4 + c] += p[c];
            if (t == c) { acc[c] += p[c]; acc[8 + c] += logf(fminf(fmaxf(p[c], kClampMin), 1.f)); }
        }
    }
    block_reduce_atomic<12>(acc, sums + (size_t)n * 12);
}

// dP[n][v][c] = coef[n][4+c] + [t==c] (coef[n][c] + coef[n][8+c] * (clamp passes ? 1/p_c : 0))
__global__ void __launch_bounds__(256) cedice_bwd_kernel(const float* __restrict__ probs, const uint8_t* __restrict__ labels,
                                                         const float* __restrict__ coef, float* __restrict__ dprobs,
                                                         long long voxels, int b, long long total) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int n = (int)(i / voxels);
        const long long v = i - (long long)n * voxels;
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(probs) + i);
        const float p[4] = {p4.x, p4.y, p4.z, p4.w};
        const int t = labels[(size_t)(n % b) * voxels + v];
        const float* cf = coef + (size_t)n * 12;
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float g = cf[4 + c];
            if (t == c) g += cf[c] + ((p[c] >= kClampMin && p[c] <= 1.f) ? cf[8 + c] / p[c] : 0.f);
            o[c] = g;
        }
        reinterpret_cast<float4*>(dprobs)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// sums[n] = sum_{v,c} pt (log pt - log ps), both clamped to [0.005, 1]; teacher sample = n % b
__global__ void __launch_bounds__(256) kl_fwd_kernel(const float* __restrict__ ps, const float* __restrict__ pt, double* __restrict__ sums,
                                                     long long voxels, int b) {
    const int n = blockIdx.y;
    const float4* sn = reinterpret_cast<const float4*>(ps) + (size_t)n * voxels;
    const float4* tn = reinterpret_cast<const float4*>(pt) + (size_t)(n % b) * voxels;
    float acc[1] = {0.f};
    for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < voxels; v += (long long)gridDim.x * 256) {
        const float4 s4 = __ldg(sn + v), t4 = __ldg(tn + v);
        const float s[4] = {s4.x, s4.y, s4.z, s4.w}, t[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float tc = fminf(fmaxf(t[c], kClampMin), 1.f), sc = fminf(fmaxf(s[c], kClampMin), 1.f);
            acc[0] += tc * (logf(tc) - logf(sc));
        }
    }
    block_reduce_atomic<1>(acc, sums + n);
}

__global__ void __launch_bounds__(256) kl_bwd_kernel(const float* __restrict__ ps, const float* __restrict__ pt, const float* __restrict__ coef,
                                                     float* __restrict__ dps, long long voxels, int b, long long total) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int n = (int)(i / voxels);
        const long long v = i - (long long)n * voxels;
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(ps) + i);
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(pt) + (size_t)(n % b) * voxels + v);
        const float s[4] = {s4.x, s4.y, s4.z, s4.w}, t[4] = {t4.x, t4.y, t4.z, t4.w};
        const float cf = coef[n];
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float tc = fminf(fmaxf(t[c], kClampMin), 1.f);
            o[c] = (s[c] >= kClampMin && s[c] <= 1.f) ? -cf * tc / s[c] : 0.f;
        }
        reinterpret_cast<float4*>(dps)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------ prototype loss, C = 8 features
constexpr int PC = 8;

// P[n][i][c] += sum_v f[v][c] [t_v == i]
template <typename T>
__global__ void __launch_bounds__(256) proto_sums_kernel(const T* __restrict__ f, const uint8_t* __restrict__ labels, double* __restrict__ P,
                                                         long long voxels, int b) {
    const int n = blockIdx.y;
    const T* fn = f + (size_t)n * voxels * PC;
    const uint8_t* ln = labels + (size_t)(n % b) * voxels;
    float acc[4 * PC];
#pragma unroll
    for (int i = 0; i < 4 * PC; ++i) acc[i] = 0.f;
    for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < voxels; v += (long long)gridDim.x * 256) {
        float x[PC];
        VecIO<T, PC>::load(fn + v * PC, x);
        const int t = ln[v];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (t == i) {
#pragma unroll
                for (int c = 0; c < PC; ++c) acc[i * PC + c] += x[c];
            }
    }
    block_reduce_atomic<4 * PC>(acc, P + (size_t)n * 4 * PC);
}

__device__ __forceinline__ float dot8(const float* a, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < PC; ++c) s = fmaf(a[c], b[c], s);
    return s;
}

// out[n][0] += sum over present classes, voxels of d^2 ; out[n][1] += |d| ; d = cos(fs, Ps_i) - cos(ft, Pt_i)
template <typename T>
__global__ void __launch_bounds__(256) proto_fwd_kernel(const T* __restrict__ fs, const T* __restrict__ ft, const float* __restrict__ protos,
                                                        const float* __restrict__ protot, const float* __restrict__ present,
                                                        double* __restrict__ out, long long voxels, int b, float eps) {
    __shared__ float sp[2][4][PC + 1];                        // prototypes and the reciprocals of their clamped norms
    const int n = blockIdx.y, nb = n % b;
    if (threadIdx.x < 4 * PC) {
        sp[0][threadIdx.x / PC][threadIdx.x % PC] = protos[(size_t)n * 4 * PC + threadIdx.x];
        sp[1][threadIdx.x / PC][threadIdx.x % PC] = protot[(size_t)nb * 4 * PC + threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float* q = sp[threadIdx.x >> 2][threadIdx.x & 3];
        q[PC] = 1.f / fmaxf(sqrtf(dot8(q, q)), eps);
    }
    __syncthreads();
    float acc[2] = {0.f, 0.f};
    const T* fsn = fs + (size_t)n * voxels * PC;
    const T* ftn = ft + (size_t)nb * voxels * PC;
    for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < voxels; v += (long long)gridDim.x * 256) {
        float a[PC], t[PC];
        VecIO<T, PC>::load(fsn + v * PC, a);
        VecIO<T, PC>::load(ftn + v * PC, t);
        const float ina = 1.f / fmaxf(sqrtf(dot8(a, a)), eps), int_ = 1.f / fmaxf(sqrtf(dot8(t, t)), eps);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float d = dot8(a, sp[0][i]) * (ina * sp[0][i][PC]) - dot8(t, sp[1][i]) * (int_ * sp[1][i][PC]);
            const float w = present[i];
            acc[0] += w * d * d;
            acc[1] += w * fabsf(d);
        }
    }
    block_reduce_atomic<2>(acc, out + (size_t)n * 2);
}

// direct gradient wrt the student features and the reduction for the prototype gradient:
//   g_i = 2 d_i coef[n] present_i ;  dfs = sum_i g_i (Ps_i/(na nP) - cos_i fs / na^2 [|fs| > eps])
//   dPs[n][i] += sum_v g_i (fs/(na nP) - cos_i Ps_i / nP^2 [|Ps_i| > eps])
template <typename T>
__global__ void __launch_bounds__(256) proto_bwd1_kernel(const T* __restrict__ fs, const T* __restrict__ ft, const float* __restrict__ protos,
                                                         const float* __restrict__ protot, const float* __restrict__ present,
                                                         const float* __restrict__ coef, T* __restrict__ dfs, double* __restrict__ dprotos,
                                                         long long voxels, int b, float eps) {
    __shared__ float sp[2][4][PC + 3];                        // prototype, clamped norm, raw norm, 1 / clamped norm
    const int n = blockIdx.y, nb = n % b;
    if (threadIdx.x < 4 * PC) {
        sp[0][threadIdx.x / PC][threadIdx.x % PC] = protos[(size_t)n * 4 * PC + threadIdx.x];
        sp[1][threadIdx.x / PC][threadIdx.x % PC] = protot[(size_t)nb * 4 * PC + threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float* q = sp[threadIdx.x >> 2][threadIdx.x & 3];
        const float r = sqrtf(dot8(q, q));
        q[PC] = fmaxf(r, eps); q[PC + 1] = r; q[PC + 2] = 1.f / fmaxf(r, eps);
    }
    __syncthreads();
    const float cf = coef[n];
    float dP[4 * PC];
#pragma unroll
    for (int i = 0; i < 4 * PC; ++i) dP[i] = 0.f;
    const T* fsn = fs + (size_t)n * voxels * PC;
    const T* ftn = ft + (size_t)nb * voxels * PC;
    T* dn = dfs + (size_t)n * voxels * PC;
    for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < voxels; v += (long long)gridDim.x * 256) {
        float a[PC], t[PC], g[PC];
        VecIO<T, PC>::load(fsn + v * PC, a);
        VecIO<T, PC>::load(ftn + v * PC, t);
        const float ra = sqrtf(dot8(a, a));
        // two reciprocals per voxel (and one per prototype, in shared memory) instead of five IEEE divisions per voxel and class: the
        // kernel was bound by its ~20 division sequences per voxel (0.2 ms per step for 0.2 GB of traffic)
        const float ina = 1.f / fmaxf(ra, eps), int_ = 1.f / fmaxf(sqrtf(dot8(t, t)), eps);
#pragma unroll
        for (int c = 0; c < PC; ++c) g[c] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float inP = sp[0][i][PC + 2];
            const float cs = dot8(a, sp[0][i]) * (ina * inP);
            const float d = cs - dot8(t, sp[1][i]) * (int_ * sp[1][i][PC + 2]);
            const float gi = 2.f * d * cf * present[i];
            const float k1 = gi * (ina * inP);
            const float k2 = ra > eps ? gi * cs * (ina * ina) : 0.f;
            const float k3 = sp[0][i][PC + 1] > eps ? gi * cs * (inP * inP) : 0.f;
#pragma unroll
            for (int c = 0; c < PC; ++c) {
                g[c] += k1 * sp[0][i][c] - k2 * a[c];
                dP[i * PC + c] += k1 * a[c] - k3 * sp[0][i][c];
            }
        }
        VecIO<T, PC>::store(dn + v * PC, g);
    }
    block_reduce_atomic<4 * PC>(dP, dprotos + (size_t)n * 4 * PC);
}

// dfs[v] += dproto[n][t_v]   (gradient through the masked class means)
template <typename T>
__global__ void __launch_bounds__(256) proto_bwd2_kernel(const uint8_t* __restrict__ labels, const float* __restrict__ dproto, T* __restrict__ dfs,
                                                         long long voxels, int b, long long total) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int n = (int)(i / voxels);
        const long long v = i - (long long)n * voxels;
        const int t = labels[(size_t)(n % b) * voxels + v];
        float g[PC];
        VecIO<T, PC>::load(dfs + i * PC, g);
        const float* dp = dproto + ((size_t)n * 4 + t) * PC;
#pragma unroll
        for (int c = 0; c < PC; ++c) g[c] += dp[c];
        VecIO<T, PC>::store(dfs + i * PC, g);
    }
}

// ------------------------------------------------------------------------------------ fused logit-level loss pass
// One pass over the logits of P decoder passes (sample index p*B + b) for the losses that live at the logits' own resolution
// (reference rfnet.py:284-377 / criterions.py:25-38, 59-76, 92-103), without materialising any probability tensor:
//   mode 0 (teacher / students): pass 0 -> softmax (T = 1) -> the CE / Dice sums A, L, E (and, optionally, the probabilities
//           themselves: Model.forward returns them); passes 1..P-1 -> KL(clamp(softmax(l_0 / T)) || clamp(softmax(l_p / T)));
//   mode 1 (all supervised): every pass -> the CE / Dice sums (the per-modality decoder_sep predictions).
// The backward kernel reads the logits again and writes d loss / d logits directly (softmax adjoint fused in).
template <int P> struct LogitLossAcc { static constexpr int kCe0 = 12 + (P - 1), kCeAll = 12 * P; };

__device__ __forceinline__ void softmax4_reg(const float* l, float inv_temp, float* p) {
    const float a = l[0] * inv_temp, b = l[1] * inv_temp, c = l[2] * inv_temp, d = l[3] * inv_temp;
    const float m = fmaxf(fmaxf(a, b), fmaxf(c, d));
    p[0] = expf(a - m); p[1] = expf(b - m); p[2] = expf(c - m); p[3] = expf(d - m);
    const float r = 1.f / (p[0] + p[1] + p[2] + p[3]);
    p[0] *= r; p[1] *= r; p[2] *= r; p[3] *= r;
}

template <typename T, int P, int MODE>
__global__ void __launch_bounds__(256) logit_loss_fwd_kernel(const T* __restrict__ logits, const uint8_t* __restrict__ labels,
                                                            float* __restrict__ probs0, double* __restrict__ ce_sums,
                                                            double* __restrict__ kl_sums, long long voxels, int B, float inv_temp) {
    constexpr int NA = MODE == 0 ? 12 + (P - 1) : 12 * P;
    const int b = blockIdx.y;
    float acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.f;
    const uint8_t* lb = labels + (size_t)b * voxels;
    for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < voxels; v += (long long)gridDim.x * 256) {
        const int t = lb[v];
        float l0[4], p[4];
        load4(logits + ((size_t)b * voxels + v) * 4, l0);
        softmax4_reg(l0, 1.f, p);
        if (probs0 != nullptr) reinterpret_cast<float4*>(probs0)[(size_t)b * voxels + v] = make_float4(p[0], p[1], p[2], p[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[4 + c] += p[c];
            if (t == c) { acc[c] += p[c]; acc[8 + c] += logf(fminf(fmaxf(p[c], kClampMin), 1.f)); }
        }
        if (MODE == 0) {
            float tc[4], ltc[4];
            softmax4_reg(l0, inv_temp, p);
#pragma unroll
            for (int c = 0; c < 4; ++c) { tc[c] = fminf(fmaxf(p[c], kClampMin), 1.f); ltc[c] = logf(tc[c]); }
#pragma unroll
            for (int q = 1; q < P; ++q) {
                float lq[4];
                load4(logits + (((size_t)q * B + b) * voxels + v) * 4, lq);
                softmax4_reg(lq, inv_temp, p);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[12 + q - 1] += tc[c] * (ltc[c] - logf(fminf(fmaxf(p[c], kClampMin), 1.f)));
            }
        } else {
#pragma unroll
            for (int q = 1; q < P; ++q) {
                float lq[4];
                load4(logits + (((size_t)q * B + b) * voxels + v) * 4, lq);
                softmax4_reg(lq, 1.f, p);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[q * 12 + 4 + c] += p[c];
                    if (t == c) { acc[q * 12 + c] += p[c]; acc[q * 12 + 8 + c] += logf(fminf(fmaxf(p[c], kClampMin), 1.f)); }
                }
            }
        }
    }
    // block reduction: 12 CE/Dice sums per supervised pass, one KL sum per student pass
    __shared__ float red[8][NA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const float sv = warp_sum(acc[i]);
        if (lane == 0) red[wid][i] = sv;
    }
    __syncthreads();
    if (threadIdx.x < NA) {
        float sv = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sv += red[w][threadIdx.x];
        const int i = threadIdx.x;
        if (MODE == 0) {
            if (i < 12) atomicAdd(ce_sums + (size_t)b * 12 + i, (double)sv);
            else atomicAdd(kl_sums + (size_t)(i - 12) * B + b, (double)sv);
        } else {
            atomicAdd(ce_sums + ((size_t)(i / 12) * B + b) * 12 + (i % 12), (double)sv);
        }
    }
}

// d/d logit of a supervised pass: g_c = coef[4+c] + [t==c] (coef[c] + coef[8+c] / p_c inside the clamp) (+ dprobs_c); dl = p (g - <p, g>)
__device__ __forceinline__ void cedice_logit_grad(const float* p, int t, const float* cf, const float* extra, float* o) {
    float g[4], dot = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float gc = cf[4 + c] + (extra != nullptr ? extra[c] : 0.f);
        if (t == c) gc += cf[c] + ((p[c] >= kClampMin && p[c] <= 1.f) ? cf[8 + c] / p[c] : 0.f);
        g[c] = gc; dot += p[c] * gc;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) o[c] = p[c] * (g[c] - dot);
}

template <typename T, int P, int MODE>
__global__ void __launch_bounds__(256) logit_loss_bwd_kernel(const T* __restrict__ logits, const uint8_t* __restrict__ labels,
                                                            const float* __restrict__ ce_coef, const float* __restrict__ kl_coef,
                                                            const float* __restrict__ dprobs0, T* __restrict__ dlogits,
                                                            long long voxels, int B, float inv_temp, long long total) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int b = (int)(i / voxels);
        const long long v = i - (long long)b * voxels;
        const int t = labels[i];
        float l0[4], p[4], o[4];
        load4(logits + (size_t)i * 4, l0);
        softmax4_reg(l0, 1.f, p);
        float ex[4];
        if (dprobs0 != nullptr) { const float4 e4 = __ldg(reinterpret_cast<const float4*>(dprobs0) + i); ex[0] = e4.x; ex[1] = e4.y; ex[2] = e4.z; ex[3] = e4.w; }
        cedice_logit_grad(p, t, ce_coef + (size_t)b * 12, dprobs0 != nullptr ? ex : nullptr, o);
        VecIO<T, 4>::store(dlogits + (size_t)i * 4, o);
        if (MODE == 0) {
            float tc[4];
            softmax4_reg(l0, inv_temp, p);
#pragma unroll
            for (int c = 0; c < 4; ++c) tc[c] = fminf(fmaxf(p[c], kClampMin), 1.f);
#pragma unroll
            for (int q = 1; q < P; ++q) {
                const size_t row = ((size_t)q * B + b) * voxels + v;
                float lq[4], g[4], dot = 0.f;
                load4(logits + row * 4, lq);
                softmax4_reg(lq, inv_temp, p);
                const float cf = kl_coef[(size_t)(q - 1) * B + b];
#pragma unroll
                for (int c = 0; c < 4; ++c) { g[c] = (p[c] >= kClampMin && p[c] <= 1.f) ? -cf * tc[c] / p[c] : 0.f; dot += p[c] * g[c]; }
#pragma unroll
                for (int c = 0; c < 4; ++c) o[c] = inv_temp * p[c] * (g[c] - dot);
                VecIO<T, 4>::store(dlogits + row * 4, o);
            }
        } else {
#pragma unroll
            for (int q = 1; q < P; ++q) {
                const size_t row = ((size_t)q * B + b) * voxels + v;
                float lq[4];
                load4(logits + row * 4, lq);
                softmax4_reg(lq, 1.f, p);
                cedice_logit_grad(p, t, ce_coef + ((size_t)q * B + b) * 12, nullptr, o);
                VecIO<T, 4>::store(dlogits + row * 4, o);
            }
        }
    }
}

int ew_blocks(long long work) {
    long long bl = (work + 255) / 256;
    if (bl > 148LL * 16) bl = 148LL * 16;
    return (int)(bl < 1 ? 1 : bl);
}
int red_blocks(long long voxels, int n) {
    long long bl = (voxels + 256 * 8 - 1) / (256 * 8);
    const long long cap = (148LL * 8 + n - 1) / n;
    if (bl > cap) bl = cap;
    return (int)(bl < 1 ? 1 : bl);
}

}  // namespace

namespace {
template <typename T, int P, int MODE>
void launch_logit_loss(bool bwd, const void* logits, const uint8_t* labels, float* probs0, double* ce_sums, double* kl_sums,
                       const float* ce_coef, const float* kl_coef, const float* dprobs0, void* dlogits, int b, long long voxels,
                       float inv_temp, cudaStream_t st) {
    if (!bwd) {
        logit_loss_fwd_kernel<T, P, MODE><<<dim3(red_blocks(voxels, b), b), 256, 0, st>>>((const T*)logits, labels, probs0, ce_sums, kl_sums,
                                                                                       voxels, b, inv_temp);
    } else {
        const long long total = (long long)b * voxels;
        logit_loss_bwd_kernel<T, P, MODE><<<ew_blocks(total), 256, 0, st>>>((const T*)logits, labels, ce_coef, kl_coef, dprobs0, (T*)dlogits,
                                                                          voxels, b, inv_temp, total);
    }
}

int dispatch_logit_loss(bool bwd, int dtype, const void* logits, const uint8_t* labels, float* probs0, double* ce_sums, double* kl_sums,
                        const float* ce_coef, const float* kl_coef, const float* dprobs0, void* dlogits, int passes, int b,
                        long long voxels, int mode, float inv_temp, cudaStream_t st) {
#define LL_CASE(P_, M_)                                                                                                                   \
    if (passes == P_ && mode == M_) {                                                                                                     \
        if (dtype == PB_BF16) launch_logit_loss<bf16, P_, M_>(bwd, logits, labels, probs0, ce_sums, kl_sums, ce_coef, kl_coef, dprobs0,   \
                                                              dlogits, b, voxels, inv_temp, st);                                          \
        else launch_logit_loss<float, P_, M_>(bwd, logits, labels, probs0, ce_sums, kl_sums, ce_coef, kl_coef, dprobs0, dlogits, b,       \
                                              voxels, inv_temp, st);                                                                      \
        return 0;                                                                                                                         \
    }
    LL_CASE(5, 0) LL_CASE(1, 0) LL_CASE(4, 1)
#undef LL_CASE
    pb_set_error("logit_loss: unsupported (passes %d, mode %d): 5/0 (teacher + 4 students), 1/0, 4/1 (4 supervised passes)", passes, mode);
    return PB_EUNSUPPORTED;
}
}  // namespace

extern "C" int pb_logit_loss_fwd(int dtype, const void* logits, const uint8_t* labels, float* probs0, double* ce_sums, double* kl_sums,
                                 int passes, int b, long long voxels, int mode, float inv_temp, pb_stream_t stream) {
    PB_CHECK_ARG(logits && labels && ce_sums && passes >= 1 && b > 0 && voxels > 0, "bad argument");
    PB_CHECK_ARG(mode == 1 || passes == 1 || kl_sums, "kl_sums needed for student passes");
    if (int e = dispatch_logit_loss(false, dtype, logits, labels, probs0, ce_sums, kl_sums, nullptr, nullptr, nullptr, nullptr, passes, b,
                                    voxels, mode, inv_temp, (cudaStream_t)stream)) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_logit_loss_bwd(int dtype, const void* logits, const uint8_t* labels, const float* ce_coef, const float* kl_coef,
                                 const float* dprobs0, void* dlogits, int passes, int b, long long voxels, int mode, float inv_temp,
                                 pb_stream_t stream) {
    PB_CHECK_ARG(logits && labels && ce_coef && dlogits && passes >= 1 && b > 0 && voxels > 0, "bad argument");
    PB_CHECK_ARG(mode == 1 || passes == 1 || kl_coef, "kl_coef needed for student passes");
    if (int e = dispatch_logit_loss(true, dtype, logits, labels, nullptr, nullptr, nullptr, ce_coef, kl_coef, dprobs0, dlogits, passes, b,
                                    voxels, mode, inv_temp, (cudaStream_t)stream)) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_softmax4(int dtype, const void* logits, float* probs, long long rows, float inv_temp, pb_stream_t stream) {
    PB_CHECK_ARG(logits && probs && rows > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PB_BF16) softmax4_kernel<bf16><<<ew_blocks(rows), 256, 0, st>>>((const bf16*)logits, probs, rows, inv_temp);
    else softmax4_kernel<float><<<ew_blocks(rows), 256, 0, st>>>((const float*)logits, probs, rows, inv_temp);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_softmax4_bwd(int dtype, const float* probs, const float* dprobs, void* dlogits, long long rows, float inv_temp,
                               pb_stream_t stream) {
    PB_CHECK_ARG(probs && dprobs && dlogits && rows > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PB_BF16) softmax4_bwd_kernel<bf16><<<ew_blocks(rows), 256, 0, st>>>(probs, dprobs, (bf16*)dlogits, rows, inv_temp);
    else softmax4_bwd_kernel<float><<<ew_blocks(rows), 256, 0, st>>>(probs, dprobs, (float*)dlogits, rows, inv_temp);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_cedice_fwd(const float* probs, const uint8_t* labels, double* sums, int n, int b, long long voxels, pb_stream_t stream) {
    PB_CHECK_ARG(probs && labels && sums && n > 0 && b > 0 && voxels > 0, "bad argument");
    cedice_fwd_kernel<<<dim3(red_blocks(voxels, n), n), 256, 0, (cudaStream_t)stream>>>(probs, labels, sums, voxels, b);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_cedice_bwd(const float* probs, const uint8_t* labels, const float* coef, float* dprobs, int n, int b,
                             long long voxels, pb_stream_t stream) {
    PB_CHECK_ARG(probs && labels && coef && dprobs && n > 0 && b > 0 && voxels > 0, "bad argument");
    const long long total = (long long)n * voxels;
    cedice_bwd_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(probs, labels, coef, dprobs, voxels, b, total);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_kl_fwd(const float* ps, const float* pt, double* sums, int n, int b, long long voxels, pb_stream_t stream) {
    PB_CHECK_ARG(ps && pt && sums && n > 0 && b > 0 && voxels > 0, "bad argument");
    kl_fwd_kernel<<<dim3(red_blocks(voxels, n), n), 256, 0, (cudaStream_t)stream>>>(ps, pt, sums, voxels, b);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_kl_bwd(const float* ps, const float* pt, const float* coef, float* dps, int n, int b, long long voxels,
                         pb_stream_t stream) {
    PB_CHECK_ARG(ps && pt && coef && dps && n > 0 && b > 0 && voxels > 0, "bad argument");
    const long long total = (long long)n * voxels;
    kl_bwd_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(ps, pt, coef, dps, voxels, b, total);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_proto_sums(int dtype, const void* f, const uint8_t* labels, double* P, int n, int b, long long voxels, int c,
                             pb_stream_t stream) {
    PB_CHECK_ARG(f && labels && P && n > 0 && b > 0 && voxels > 0, "bad argument");
    PB_CHECK_ARG(c == PC, "prototype kernels are specialised for 8 feature channels (basic_dims)");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(red_blocks(voxels, n), n);
    if (dtype == PB_BF16) proto_sums_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)f, labels, P, voxels, b);
    else proto_sums_kernel<float><<<grid, 256, 0, st>>>((const float*)f, labels, P, voxels, b);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_proto_fwd(int dtype, const void* fs, const void* ft, const float* protos, const float* protot, const float* present,
                            double* out, int n, int b, long long voxels, int c, float eps, pb_stream_t stream) {
    PB_CHECK_ARG(fs && ft && protos && protot && present && out && n > 0 && b > 0 && voxels > 0, "bad argument");
    PB_CHECK_ARG(c == PC, "prototype kernels are specialised for 8 feature channels (basic_dims)");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(red_blocks(voxels, n), n);
    if (dtype == PB_BF16) proto_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)fs, (const bf16*)ft, protos, protot, present, out, voxels, b, eps);
    else proto_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)fs, (const float*)ft, protos, protot, present, out, voxels, b, eps);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_proto_bwd1(int dtype, const void* fs, const void* ft, const float* protos, const float* protot, const float* present,
                             const float* coef, void* dfs, double* dprotos, int n, int b, long long voxels, int c, float eps,
                             pb_stream_t stream) {
    PB_CHECK_ARG(fs && ft && protos && protot && present && coef && dfs && dprotos && n > 0 && b > 0 && voxels > 0, "bad argument");
    PB_CHECK_ARG(c == PC, "prototype kernels are specialised for 8 feature channels (basic_dims)");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(red_blocks(voxels, n), n);
    if (dtype == PB_BF16)
        proto_bwd1_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)fs, (const bf16*)ft, protos, protot, present, coef, (bf16*)dfs, dprotos, voxels, b, eps);
    else
        proto_bwd1_kernel<float><<<grid, 256, 0, st>>>((const float*)fs, (const float*)ft, protos, protot, present, coef, (float*)dfs, dprotos, voxels, b, eps);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_proto_bwd2(int dtype, const uint8_t* labels, const float* dproto, void* dfs, int n, int b, long long voxels, int c,
                             pb_stream_t stream) {
    PB_CHECK_ARG(labels && dproto && dfs && n > 0 && b > 0 && voxels > 0, "bad argument");
    PB_CHECK_ARG(c == PC, "prototype kernels are specialised for 8 feature channels (basic_dims)");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)n * voxels;
    if (dtype == PB_BF16) proto_bwd2_kernel<bf16><<<ew_blocks(total), 256, 0, st>>>(labels, dproto, (bf16*)dfs, voxels, b, total);
    else proto_bwd2_kernel<float><<<ew_blocks(total), 256, 0, st>>>(labels, dproto, (float*)dfs, voxels, b, total);
    PB_CHECK_LAUNCH();
    return PB_OK;
}
