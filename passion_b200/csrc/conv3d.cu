// 3-D convolution, direct FFMA path (fp32 accumulate): forward, dgrad, wgrad.
//
// This is the precision-reference / small-channel path of the library: it serves the fp32
// check mode (rel-L2 <= 1e-4 vs the CPU oracle) and every (Cin, Cout) class, including the
// C in {1,2,4} layers that cannot feed a tensor-core tile.  The bf16 tcgen05 implicit-GEMM
// path for the dominant classes lives in conv3d_tc.cu.
//
// Replaces nn.Conv3d of general_conv3d (reference models/blocks.py:357) incl. its
// padding_mode='reflect', the torch.cat feeding decoder convs (models/rfnet.py:75,79,83,133,139,145)
// and the four separate modality encoders (models/rfnet.py:234-237) as weight groups.
#include "common.cuh"
#include <stdlib.h>

namespace {

struct ConvK {
    int N, Di, Hi, Wi, Do, Ho, Wo, C0, C1, Cin, Cout, K, S, pad, reflect, groups, npg;
    long long Vi, Vo;
};

template <int N> __device__ __forceinline__ void lds_vec(const float* p, float* o) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            float4 v = reinterpret_cast<const float4*>(p)[i];
            o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w;
        }
    } else if constexpr (N == 2) {
        float2 v = *reinterpret_cast<const float2*>(p); o[0] = v.x; o[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = p[i];
    }
}

// acc[j] += sum_i xv[i] * wr[i*NO + j]
template <int NI, int NO> __device__ __forceinline__ void fma_block(const float* xv, const float* wr, float* acc) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        float wv[NO];
        lds_vec<NO>(wr + i * NO, wv);
#pragma unroll
        for (int j = 0; j < NO; ++j) acc[j] = fmaf(xv[i], wv[j], acc[j]);
    }
}

// ------------------------------------------------------------------------------------ forward
// one thread = one output voxel x CO_T output channels; weights of the (group, co-chunk) in smem.
template <typename T, int CI_V, int CO_T>
__global__ void __launch_bounds__(128) conv_fwd_kernel(ConvK p, const T* __restrict__ x0, const T* __restrict__ x1,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       T* __restrict__ y, double* __restrict__ stats) {
    extern __shared__ __align__(16) float wsm[];              // [taps][Cin][CO_T]
    __shared__ float red[4][CO_T * 2];
    const int n = blockIdx.z, coc = blockIdx.y, g = n / p.npg;
    const int taps = p.K * p.K * p.K;
    const float* wg = w + (size_t)g * taps * p.Cin * p.Cout + coc * CO_T;
    for (int i = threadIdx.x; i < taps * p.Cin * CO_T; i += 128) {
        int j = i % CO_T, r = i / CO_T;
        wsm[i] = wg[(size_t)r * p.Cout + j];
    }
    __syncthreads();

    // persistent over the voxel blocks: the weight slice above (up to 200 KB for the wide layers) is staged once per CTA
    float ssum[CO_T], ssq[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
    for (long long o = (long long)blockIdx.x * 128 + threadIdx.x; o < p.Vo; o += (long long)gridDim.x * 128) {
    const int ow = (int)(o % p.Wo);
    const int t1 = (int)(o / p.Wo);
    const int oh = t1 % p.Ho, od = t1 / p.Ho;

    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = bias ? bias[(size_t)g * p.Cout + coc * CO_T + j] : 0.f;

    for (int kd = 0; kd < p.K; ++kd) {
        int id = od * p.S + kd - p.pad;
        if (p.reflect) id = reflect_idx(id, p.Di); else if (id < 0 || id >= p.Di) continue;
        for (int kh = 0; kh < p.K; ++kh) {
            int ih = oh * p.S + kh - p.pad;
            if (p.reflect) ih = reflect_idx(ih, p.Hi); else if (ih < 0 || ih >= p.Hi) continue;
            for (int kw = 0; kw < p.K; ++kw) {
                int iw = ow * p.S + kw - p.pad;
                if (p.reflect) iw = reflect_idx(iw, p.Wi); else if (iw < 0 || iw >= p.Wi) continue;
                const size_t vox = (((size_t)n * p.Di + id) * p.Hi + ih) * p.Wi + iw;
                const float* wt = wsm + ((kd * p.K + kh) * p.K + kw) * p.Cin * CO_T;
                const T* p0 = x0 + vox * p.C0;
                for (int c = 0; c < p.C0; c += CI_V) {
                    float xv[CI_V];
                    VecIO<T, CI_V>::load(p0 + c, xv);
                    fma_block<CI_V, CO_T>(xv, wt + c * CO_T, acc);
                }
                if (p.C1) {
                    const T* p1 = x1 + vox * p.C1;
                    const float* wt1 = wt + p.C0 * CO_T;
                    for (int c = 0; c < p.C1; c += CI_V) {
                        float xv[CI_V];
                        VecIO<T, CI_V>::load(p1 + c, xv);
                        fma_block<CI_V, CO_T>(xv, wt1 + c * CO_T, acc);
                    }
                }
            }
        }
    }
    VecIO<T, CO_T>::store(y + ((size_t)n * p.Vo + o) * p.Cout + coc * CO_T, acc);
#pragma unroll
    for (int j = 0; j < CO_T; ++j) { ssum[j] += acc[j]; ssq[j] += acc[j] * acc[j]; }
    }
    if (stats) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < CO_T; ++j) {
            float s = ssum[j];
            float q = ssq[j];
            s = warp_sum(s); q = warp_sum(q);
            if (lane == 0) { red[wid][2 * j] = s; red[wid][2 * j + 1] = q; }
        }
        __syncthreads();
        if (threadIdx.x < CO_T * 2) {
            float v = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
            atomicAdd(&stats[((size_t)n * p.Cout + coc * CO_T + (threadIdx.x >> 1)) * 2 + (threadIdx.x & 1)], (double)v);
        }
    }
}

// ------------------------------------------------------------------------------------ 1x1x1 forward / data gradient
// Pointwise conv: y[v][co] = b[co] + sum_ci x[v][ci] w[ci][co].  HBM-bound, so the kernel is built around the memory
// system: a fixed grid of CTAs strides over the voxels of one sample, every thread keeps UNR voxels (independent 16-byte
// loads) in flight, the statistics are accumulated in registers over the whole stride loop and reduced ONCE per CTA.
// The data gradient of a 1x1x1 conv is the same operation on dy with the transposed weights and the output split
// across the two sources (y0 | y1).
struct PwK {
    int N, C0, C1, Cin, CO0, CO1, Cout, npg;
    long long V;
};

template <typename T, int CI_V, int CO_T, int UNR>
__global__ void __launch_bounds__(256) pw_conv_kernel(PwK p, const T* __restrict__ x0, const T* __restrict__ x1,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      T* __restrict__ y0, T* __restrict__ y1, double* __restrict__ stats) {
    extern __shared__ __align__(16) float wsm[];              // [Cin][CO_T]
    __shared__ float red[8][CO_T * 2];
    const int n = blockIdx.z, coc = blockIdx.y, g = n / p.npg;
    const float* wg = w + (size_t)g * p.Cin * p.Cout + coc * CO_T;
    for (int i = threadIdx.x; i < p.Cin * CO_T; i += 256) wsm[i] = wg[(size_t)(i / CO_T) * p.Cout + (i % CO_T)];
    __syncthreads();
    float b[CO_T], s1[CO_T], s2[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) { b[j] = bias ? bias[(size_t)g * p.Cout + coc * CO_T + j] : 0.f; s1[j] = 0.f; s2[j] = 0.f; }
    const long long stride = (long long)gridDim.x * 256;
    const int co0 = coc * CO_T;
    const T* xs0 = x0 + (size_t)n * p.V * p.C0;
    const T* xs1 = p.C1 ? x1 + (size_t)n * p.V * p.C1 : nullptr;
    for (long long v0 = (long long)blockIdx.x * 256 + threadIdx.x; v0 < p.V; v0 += stride * UNR) {
        float acc[UNR][CO_T];
        long long vv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long v = v0 + u * stride;
            vv[u] = v < p.V ? v : v0;                         // tail: recompute voxel v0, the store is skipped
#pragma unroll
            for (int j = 0; j < CO_T; ++j) acc[u][j] = b[j];
        }
        for (int c = 0; c < p.C0; c += CI_V) {
            float xv[UNR][CI_V];
#pragma unroll
            for (int u = 0; u < UNR; ++u) VecIO<T, CI_V>::load(xs0 + vv[u] * p.C0 + c, xv[u]);
#pragma unroll
            for (int i = 0; i < CI_V; ++i) {
                float wv[CO_T];
                lds_vec<CO_T>(wsm + (c + i) * CO_T, wv);
#pragma unroll
                for (int u = 0; u < UNR; ++u)
#pragma unroll
                    for (int j = 0; j < CO_T; ++j) acc[u][j] = fmaf(xv[u][i], wv[j], acc[u][j]);
            }
        }
        for (int c = 0; c < p.C1; c += CI_V) {
            float xv[UNR][CI_V];
#pragma unroll
            for (int u = 0; u < UNR; ++u) VecIO<T, CI_V>::load(xs1 + vv[u] * p.C1 + c, xv[u]);
#pragma unroll
            for (int i = 0; i < CI_V; ++i) {
                float wv[CO_T];
                lds_vec<CO_T>(wsm + (p.C0 + c + i) * CO_T, wv);
#pragma unroll
                for (int u = 0; u < UNR; ++u)
#pragma unroll
                    for (int j = 0; j < CO_T; ++j) acc[u][j] = fmaf(xv[u][i], wv[j], acc[u][j]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long v = v0 + u * stride;
            if (v < p.V) {
                const size_t vox = (size_t)n * p.V + v;
                T* dst = co0 < p.CO0 ? y0 + vox * p.CO0 + co0 : y1 + vox * p.CO1 + (co0 - p.CO0);
                VecIO<T, CO_T>::store(dst, acc[u]);
#pragma unroll
                for (int j = 0; j < CO_T; ++j) { s1[j] += acc[u][j]; s2[j] += acc[u][j] * acc[u][j]; }
            }
        }
    }
    if (stats) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < CO_T; ++j) {
            const float s = warp_sum(s1[j]), q = warp_sum(s2[j]);
            if (lane == 0) { red[wid][2 * j] = s; red[wid][2 * j + 1] = q; }
        }
        __syncthreads();
        if (threadIdx.x < CO_T * 2) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) v += (double)red[k][threadIdx.x];
            atomicAdd(&stats[((size_t)n * p.Cout + co0 + (threadIdx.x >> 1)) * 2 + (threadIdx.x & 1)], v);
        }
    }
}

// bf16 classes with 8-channel chunk counts known at compile time (PB_PW2=0 routes them back to pw_conv_kernel for A/B runs;
// measured on a B200 in round 2: 21.16 -> 20.38 ms per step, the 1x1x1 forward / data-gradient families 3.0 -> 2.8 ms).
// Same operation and thread mapping as pw_conv_kernel, built for more bytes in flight: the channel-chunk counts of the two
// sources are template parameters, so a thread first issues ALL the 16-byte loads of its UNR voxels (packed: one register
// per two bf16 values) and only then unpacks chunk by chunk.  ncu on pw_conv_kernel: 127 registers -> two CTAs per SM with
// 2-4 loads per thread in flight (16-32 KB per SM), 2.5-3.5 TB/s.
template <typename T, int CO_T, int NCH0, int NCH1, int UNR, int MINB>
__global__ void __launch_bounds__(256, MINB) pw_conv2_kernel(PwK p, const T* __restrict__ x0, const T* __restrict__ x1,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             T* __restrict__ y0, T* __restrict__ y1, double* __restrict__ stats) {
    constexpr int CI_V = 8;
    extern __shared__ __align__(16) float wsm[];              // [Cin][CO_T]
    __shared__ float red[8][CO_T * 2];
    __shared__ float bsm[CO_T];
    const int n = blockIdx.z, coc = blockIdx.y, g = n / p.npg;
    const float* wg = w + (size_t)g * p.Cin * p.Cout + coc * CO_T;
    for (int i = threadIdx.x; i < p.Cin * CO_T; i += 256) wsm[i] = wg[(size_t)(i / CO_T) * p.Cout + (i % CO_T)];
    if (threadIdx.x < CO_T) bsm[threadIdx.x] = bias ? bias[(size_t)g * p.Cout + coc * CO_T + threadIdx.x] : 0.f;
    __syncthreads();
    float s1[CO_T], s2[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    const long long stride = (long long)gridDim.x * 256;
    const int co0 = coc * CO_T;
    const T* xs0 = x0 + (size_t)n * p.V * (NCH0 * CI_V);
    const T* xs1 = NCH1 ? x1 + (size_t)n * p.V * (NCH1 * CI_V) : nullptr;
    for (long long v0 = (long long)blockIdx.x * 256 + threadIdx.x; v0 < p.V; v0 += stride * UNR) {
        RawVec<T, CI_V> r0[UNR][NCH0], r1[UNR][NCH1 ? NCH1 : 1];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long v = v0 + u * stride;
            const long long vv = v < p.V ? v : v0;            // tail: recompute voxel v0, the store is skipped
#pragma unroll
            for (int k = 0; k < NCH0; ++k) r0[u][k].load(xs0 + vv * (NCH0 * CI_V) + k * CI_V);
#pragma unroll
            for (int k = 0; k < NCH1; ++k) r1[u][k].load(xs1 + vv * (NCH1 * CI_V) + k * CI_V);
        }
        float acc[UNR][CO_T];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int j = 0; j < CO_T; ++j) acc[u][j] = bsm[j];
#pragma unroll
        for (int k = 0; k < NCH0 + NCH1; ++k) {
            float xv[UNR][CI_V];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (k < NCH0) r0[u][k < NCH0 ? k : 0].unpack(xv[u]);
                else r1[u][k >= NCH0 ? k - NCH0 : 0].unpack(xv[u]);
            }
#pragma unroll
            for (int i = 0; i < CI_V; ++i) {
                float wv[CO_T];
                lds_vec<CO_T>(wsm + (k * CI_V + i) * CO_T, wv);
#pragma unroll
                for (int u = 0; u < UNR; ++u)
#pragma unroll
                    for (int j = 0; j < CO_T; ++j) acc[u][j] = fmaf(xv[u][i], wv[j], acc[u][j]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long v = v0 + u * stride;
            if (v < p.V) {
                const size_t vox = (size_t)n * p.V + v;
                T* dst = co0 < p.CO0 ? y0 + vox * p.CO0 + co0 : y1 + vox * p.CO1 + (co0 - p.CO0);
                VecIO<T, CO_T>::store(dst, acc[u]);
#pragma unroll
                for (int j = 0; j < CO_T; ++j) { s1[j] += acc[u][j]; s2[j] += acc[u][j] * acc[u][j]; }
            }
        }
    }
    if (stats) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < CO_T; ++j) {
            const float s = warp_sum(s1[j]), q = warp_sum(s2[j]);
            if (lane == 0) { red[wid][2 * j] = s; red[wid][2 * j + 1] = q; }
        }
        __syncthreads();
        if (threadIdx.x < CO_T * 2) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) v += (double)red[k][threadIdx.x];
            atomicAdd(&stats[((size_t)n * p.Cout + co0 + (threadIdx.x >> 1)) * 2 + (threadIdx.x & 1)], v);
        }
    }
}

// ------------------------------------------------------------------------------------ dgrad
// Exact adjoint of the forward gather, as a gather over dy: one thread = one INPUT voxel x CI_T
// input channels.  For every padded position p that the forward's index map sends to this voxel
// (p = i, plus the reflected twins -i / 2(D-1)-i next to the faces) and every tap k, the
// contributing output is o = (p + pad - k) / stride when that is an in-range integer.
__device__ __forceinline__ int dgrad_cand(int c, int i, int D, int pad, int reflect, bool& ok) {
    if (c == 0) { ok = true; return i; }
    if (c == 1) { ok = reflect && i >= 1 && i <= pad; return -i; }
    int q = 2 * (D - 1) - i;
    ok = reflect && q >= D && q <= D - 1 + pad;
    return q;
}

template <typename T, int CO_V, int CI_T, bool MIRROR_ONLY>
__global__ void __launch_bounds__(128) conv_dgrad_kernel(ConvK p, const T* __restrict__ dy, const float* __restrict__ wt,
                                                         T* __restrict__ dx0, T* __restrict__ dx1) {
    extern __shared__ __align__(16) float wsm[];              // [taps][Cout][CI_T]
    const int n = blockIdx.z, cic = blockIdx.y, g = n / p.npg;
    const int taps = p.K * p.K * p.K;
    const float* wg = wt + (size_t)g * taps * p.Cout * p.Cin + cic * CI_T;
    for (int i = threadIdx.x; i < taps * p.Cout * CI_T; i += 128) {
        int j = i % CI_T, r = i / CI_T;
        wsm[i] = wg[(size_t)r * p.Cin + j];
    }
    __syncthreads();

    long long iv = (long long)blockIdx.x * 128 + threadIdx.x;
    int iw, ih, id;
    if (MIRROR_ONLY) {
        // Only voxels one step inside a face receive reflected-halo contributions.  Threads enumerate exactly that
        // shell: (A) id in {1, D-2}; (B) id elsewhere, ih in {1, H-2}; (C) id, ih elsewhere, iw in {1, W-2}.
        // (host guarantees D, H, W >= 4 for this mode)
        const long long nA = 2LL * p.Hi * p.Wi, nB = (long long)(p.Di - 2) * 2 * p.Wi, nC = (long long)(p.Di - 2) * (p.Hi - 2) * 2;
        if (iv >= nA + nB + nC) return;
        auto inner = [](int j, int D) { return j == 0 ? 0 : (j == D - 3 ? D - 1 : j + 1); };   // j-th index not in {1, D-2}
        if (iv < nA) {
            id = iv < (long long)p.Hi * p.Wi ? 1 : p.Di - 2;
            const int r = (int)(iv % ((long long)p.Hi * p.Wi));
            ih = r / p.Wi; iw = r % p.Wi;
        } else if (iv < nA + nB) {
            const long long r = iv - nA;
            id = inner((int)(r / (2 * p.Wi)), p.Di);
            const int r2 = (int)(r % (2 * p.Wi));
            ih = r2 < p.Wi ? 1 : p.Hi - 2; iw = r2 % p.Wi;
        } else {
            const long long r = iv - nA - nB;
            id = inner((int)(r / (2 * (p.Hi - 2))), p.Di);
            const int r2 = (int)(r % (2 * (p.Hi - 2)));
            ih = inner(r2 >> 1, p.Hi); iw = (r2 & 1) ? p.Wi - 2 : 1;
        }
        iv = ((long long)id * p.Hi + ih) * p.Wi + iw;
    } else {
        if (iv >= p.Vi) return;
        iw = (int)(iv % p.Wi);
        const int t1 = (int)(iv / p.Wi);
        ih = t1 % p.Hi; id = t1 / p.Hi;
    }

    float acc[CI_T];
#pragma unroll
    for (int j = 0; j < CI_T; ++j) acc[j] = 0.f;
    const int nc = p.reflect ? 3 : 1;

    for (int cd = 0; cd < nc; ++cd) {
        bool okd; const int pd = dgrad_cand(cd, id, p.Di, p.pad, p.reflect, okd);
        if (!okd) continue;
        for (int kd = 0; kd < p.K; ++kd) {
            const int td = pd + p.pad - kd;
            if (td < 0) continue;
            const int od = td / p.S;
            if (od * p.S != td || od >= p.Do) continue;
            for (int ch = 0; ch < nc; ++ch) {
                bool okh; const int ph = dgrad_cand(ch, ih, p.Hi, p.pad, p.reflect, okh);
                if (!okh) continue;
                for (int kh = 0; kh < p.K; ++kh) {
                    const int th = ph + p.pad - kh;
                    if (th < 0) continue;
                    const int oh = th / p.S;
                    if (oh * p.S != th || oh >= p.Ho) continue;
                    for (int cw = 0; cw < nc; ++cw) {
                        bool okw; const int pw = dgrad_cand(cw, iw, p.Wi, p.pad, p.reflect, okw);
                        if (!okw) continue;
                        if (MIRROR_ONLY && cd == 0 && ch == 0 && cw == 0) continue;   // the zero-padding part is already there
                        for (int kw = 0; kw < p.K; ++kw) {
                            const int tw = pw + p.pad - kw;
                            if (tw < 0) continue;
                            const int ow = tw / p.S;
                            if (ow * p.S != tw || ow >= p.Wo) continue;
                            const T* pdy = dy + ((((size_t)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.Cout;
                            const float* wr = wsm + ((kd * p.K + kh) * p.K + kw) * p.Cout * CI_T;
                            for (int c = 0; c < p.Cout; c += CO_V) {
                                float gv[CO_V];
                                VecIO<T, CO_V>::load(pdy + c, gv);
                                fma_block<CO_V, CI_T>(gv, wr + c * CI_T, acc);
                            }
                        }
                    }
                }
            }
        }
    }
    const int ci0 = cic * CI_T;
    const size_t vox = (size_t)n * p.Vi + iv;
    T* dst = ci0 < p.C0 ? dx0 + vox * p.C0 + ci0 : dx1 + vox * p.C1 + (ci0 - p.C0);
    if (MIRROR_ONLY) {
        float old[CI_T];
        VecIO<T, CI_T>::load(dst, old);
#pragma unroll
        for (int j = 0; j < CI_T; ++j) acc[j] += old[j];
    }
    VecIO<T, CI_T>::store(dst, acc);
}

// Same adjoint with the per-axis (output index, tap) pairs enumerated up front: along one axis an input voxel i is hit by
// at most 6 pairs (its own padded position and, next to a face, its reflected twin; for each the taps k with
// (p + pad - k) divisible by the stride).  They are packed 10 bits apiece (o: 8 bits, k: 2 bits) into one 64-bit word per
// axis, so the triple loop below visits only real contributions — the generic kernel above walks all 27 x 27
// (candidate, tap) combinations per voxel and `continue`s through most of them, which made the stride-2 data gradients
// (27/8 useful taps per voxel on average) 5-7x slower than their FLOPs.  Requires output extents <= 256.
__device__ __forceinline__ int axis_pairs(int i, int Din, int Dout, int K, int S, int pad, int reflect, unsigned long long& packed) {
    int n = 0;
    packed = 0ull;
    // 3-tap, stride 2, pad 1 away from the two indices that have a reflected twin (1 and Din - 2): closed form — an odd index is
    // reached by taps 0 and 2, an even one by tap 1.  (The general enumeration below costs ~250 instructions per axis, with
    // divisions by a run-time stride: more than the FMAs of an average voxel of the stride-2 data gradients.)
    if (K == 3 && S == 2 && pad == 1 && !(reflect && (i == 1 || i == Din - 2))) {
        if (i & 1) {
            const int o0 = (i + 1) >> 1, o1 = (i - 1) >> 1;
            if (o0 < Dout) { packed |= (unsigned long long)((o0 << 2) | 0) << (10 * n); ++n; }
            if (o1 < Dout) { packed |= (unsigned long long)((o1 << 2) | 2) << (10 * n); ++n; }
        } else {
            const int o = i >> 1;
            if (o < Dout) { packed = (unsigned long long)((o << 2) | 1); n = 1; }
        }
        return n;
    }
    const int nc = reflect ? 3 : 1;
    for (int c = 0; c < nc; ++c) {
        bool ok;
        const int pp = dgrad_cand(c, i, Din, pad, reflect, ok);
        if (!ok) continue;
        for (int k = 0; k < K; ++k) {
            const int t = pp + pad - k;
            if (t < 0) continue;
            const int o = t / S;
            if (o * S != t || o >= Dout) continue;
            packed |= (unsigned long long)((o << 2) | k) << (10 * n);
            ++n;
        }
    }
    return n;
}

template <typename T, int CO_V, int CI_T>
__global__ void __launch_bounds__(256) conv_dgrad_pairs_kernel(ConvK p, const T* __restrict__ dy, const float* __restrict__ wt,
                                                               T* __restrict__ dx0, T* __restrict__ dx1) {
    extern __shared__ __align__(16) float wsm[];              // [taps][Cout][CI_T]
    const int n = blockIdx.z, cic = blockIdx.y, g = n / p.npg;
    const int taps = p.K * p.K * p.K;
    const float* wg = wt + (size_t)g * taps * p.Cout * p.Cin + cic * CI_T;
    if constexpr (CI_T % 4 == 0) {                            // rows of CI_T floats, 16-byte aligned on both sides
        constexpr int Q = CI_T / 4;
        for (int i = threadIdx.x; i < taps * p.Cout * Q; i += blockDim.x) {
            const int j = i % Q, r = i / Q;
            reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(wg + (size_t)r * p.Cin) + j);
        }
    } else {
        for (int i = threadIdx.x; i < taps * p.Cout * CI_T; i += blockDim.x) {
            int j = i % CI_T, r = i / CI_T;
            wsm[i] = wg[(size_t)r * p.Cin + j];
        }
    }
    __syncthreads();
    // persistent over the voxels: the weight slice above is staged once per CTA.
    // Stride 2: an input index receives ONE (output, tap) pair per axis when it is even and TWO when it is odd, so the
    // voxels are visited parity class by parity class (all-odd class first): the lanes of a warp then run the same number
    // of pairs with the same taps — no divergence (a mixed warp always paid for 8 pairs, the average is 27/8) and the
    // weight rows are shared-memory broadcasts instead of 8-way bank conflicts (ncu: 14.4 K instructions per warp-voxel,
    // short-scoreboard stalls 3-5 per issue).
    const bool by_parity = p.S == 2;
    const int qd = by_parity ? (p.Di + 1) / 2 : p.Di, qh = by_parity ? (p.Hi + 1) / 2 : p.Hi, qw = by_parity ? (p.Wi + 1) / 2 : p.Wi;
    const long long cls_sz = (long long)qd * qh * qw;
    const long long n_virtual = by_parity ? 8 * cls_sz : p.Vi;
    for (long long jv = (long long)blockIdx.x * blockDim.x + threadIdx.x; jv < n_virtual; jv += (long long)gridDim.x * blockDim.x) {
    int iw, ih, id;
    if (by_parity) {
        const int cls = 7 - (int)(jv / cls_sz);
        const int r = (int)(jv - (long long)(7 - cls) * cls_sz);
        iw = 2 * (r % qw) + (cls & 1);
        const int t1 = r / qw;
        ih = 2 * (t1 % qh) + ((cls >> 1) & 1);
        id = 2 * (t1 / qh) + (cls >> 2);
        if (iw >= p.Wi || ih >= p.Hi || id >= p.Di) continue;
    } else {
        iw = (int)(jv % p.Wi);
        const int t1 = (int)(jv / p.Wi);
        ih = t1 % p.Hi; id = t1 / p.Hi;
    }
    const long long iv = ((long long)id * p.Hi + ih) * p.Wi + iw;
    unsigned long long pd, ph, pw;
    const int nd = axis_pairs(id, p.Di, p.Do, p.K, p.S, p.pad, p.reflect, pd);
    const int nh = axis_pairs(ih, p.Hi, p.Ho, p.K, p.S, p.pad, p.reflect, ph);
    const int nw = axis_pairs(iw, p.Wi, p.Wo, p.K, p.S, p.pad, p.reflect, pw);
    float acc[CI_T];
#pragma unroll
    for (int j = 0; j < CI_T; ++j) acc[j] = 0.f;
    for (int a = 0; a < nd; ++a) {
        const int ea = (int)((pd >> (10 * a)) & 1023), od = ea >> 2, kd = ea & 3;
        for (int b = 0; b < nh; ++b) {
            const int eb = (int)((ph >> (10 * b)) & 1023), oh = eb >> 2, kh = eb & 3;
            for (int c = 0; c < nw; ++c) {
                const int ec = (int)((pw >> (10 * c)) & 1023), ow = ec >> 2, kw = ec & 3;
                const T* pdy = dy + ((((size_t)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.Cout;
                const float* wr = wsm + ((kd * p.K + kh) * p.K + kw) * p.Cout * CI_T;
                for (int co = 0; co < p.Cout; co += CO_V) {
                    float gv[CO_V];
                    VecIO<T, CO_V>::load(pdy + co, gv);
                    fma_block<CO_V, CI_T>(gv, wr + co * CI_T, acc);
                }
            }
        }
    }
    const int ci0 = cic * CI_T;
    const size_t vox = (size_t)n * p.Vi + iv;
    T* dst = ci0 < p.C0 ? dx0 + vox * p.C0 + ci0 : dx1 + vox * p.C1 + (ci0 - p.C0);
    VecIO<T, CI_T>::store(dst, acc);
    }
}

// ------------------------------------------------------------------------------------ wgrad
// Persistent blocks; each stages an input halo tile and a dy tile in smem (fp32) and every thread
// owns one (tap, CIQ input channels) x CO_T slab of dw for a slice of the tile's voxels.
struct WgK {
    int td, th, tw;        // output tile
    int hd, hh, hw;        // halo tile = (t-1)*S + K
    int tiles_d, tiles_h, tiles_w;
    int cic;               // input channels per chunk
    int xs;                // smem voxel stride of the halo tile (floats)
    int nitems, slices;
    int n_cic, n_coc;
};

template <typename T, int CIQ, int CO_T, int LV>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(ConvK p, WgK q, const T* __restrict__ x0, const T* __restrict__ x1,
                                                         const T* __restrict__ dy, float* __restrict__ dw) {
    extern __shared__ __align__(16) float sm[];
    float* xsm = sm;                                          // [halo voxels][xs]
    const int halo = q.hd * q.hh * q.hw, tv = q.td * q.th * q.tw;
    float* dsm = sm + (((size_t)halo * q.xs + 3) & ~(size_t)3);  // [tile voxels][CO_T], 16 B aligned
    const int g = blockIdx.z;
    const int cic_i = blockIdx.y % q.n_cic, coc = blockIdx.y / q.n_cic;
    const int ci0 = cic_i * q.cic;                            // global input-channel offset of this chunk
    const bool from1 = ci0 >= p.C0;
    const T* xsrc = from1 ? x1 : x0;
    const int csrc = from1 ? p.C1 : p.C0, coff = from1 ? ci0 - p.C0 : ci0;
    const int tid = threadIdx.x;
    const int item = tid % q.nitems, slice = tid / q.nitems;
    const bool active = slice < q.slices;
    const int nq = q.cic / CIQ;
    const int tap = item / nq, ciq = item % nq;
    const int kd = tap / (p.K * p.K), kh = (tap / p.K) % p.K, kw = tap % p.K;

    float acc[CIQ][CO_T];
#pragma unroll
    for (int i = 0; i < CIQ; ++i)
#pragma unroll
        for (int j = 0; j < CO_T; ++j) acc[i][j] = 0.f;

    const int tiles_ps = q.tiles_d * q.tiles_h * q.tiles_w;
    const long long ntiles = (long long)p.npg * tiles_ps;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int n = g * p.npg + (int)(t / tiles_ps);
        int r = (int)(t % tiles_ps);
        const int tw_i = r % q.tiles_w; r /= q.tiles_w;
        const int th_i = r % q.tiles_h, td_i = r / q.tiles_h;
        const int od0 = td_i * q.td, oh0 = th_i * q.th, ow0 = tw_i * q.tw;
        __syncthreads();                                      // previous tile fully consumed
        // ---- stage input halo tile (reflect / zero padding resolved here)
        const int lpv = q.cic / LV;
        for (int idx = tid; idx < halo * lpv; idx += 256) {
            const int hv = idx / lpv, lc = idx % lpv;
            const int hw_i = hv % q.hw, r2 = hv / q.hw;
            const int hh_i = r2 % q.hh, hd_i = r2 / q.hh;
            int id = od0 * p.S - p.pad + hd_i, ih = oh0 * p.S - p.pad + hh_i, iw = ow0 * p.S - p.pad + hw_i;
            bool ok = true;
            if (p.reflect) {
                id = reflect_idx(id, p.Di); ih = reflect_idx(ih, p.Hi); iw = reflect_idx(iw, p.Wi);
                ok = id >= 0 && id < p.Di && ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi;   // far overhang of partial tiles
            } else {
                ok = id >= 0 && id < p.Di && ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi;
            }
            float v[LV];
            if (ok) VecIO<T, LV>::load(xsrc + ((((size_t)n * p.Di + id) * p.Hi + ih) * p.Wi + iw) * csrc + coff + lc * LV, v);
            else {
#pragma unroll
                for (int i = 0; i < LV; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < LV; ++i) xsm[(size_t)hv * q.xs + lc * LV + i] = v[i];
        }
        // ---- stage dy tile (zero outside the volume)
        for (int idx = tid; idx < tv; idx += 256) {
            const int vw = idx % q.tw, r2 = idx / q.tw;
            const int vh = r2 % q.th, vd = r2 / q.th;
            const int od = od0 + vd, oh = oh0 + vh, ow = ow0 + vw;
            float v[CO_T];
            if (od < p.Do && oh < p.Ho && ow < p.Wo)
                VecIO<T, CO_T>::load(dy + ((((size_t)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.Cout + coc * CO_T, v);
            else {
#pragma unroll
                for (int j = 0; j < CO_T; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < CO_T; ++j) dsm[(size_t)idx * CO_T + j] = v[j];
        }
        __syncthreads();
        if (active) {
            for (int v = slice; v < tv; v += q.slices) {
                const int vw = v % q.tw, r2 = v / q.tw;
                const int vh = r2 % q.th, vd = r2 / q.th;
                const int hv = ((vd * p.S + kd) * q.hh + vh * p.S + kh) * q.hw + vw * p.S + kw;
                float xv[CIQ], gv[CO_T];
                lds_vec<CIQ>(xsm + (size_t)hv * q.xs + ciq * CIQ, xv);
                lds_vec<CO_T>(dsm + (size_t)v * CO_T, gv);
#pragma unroll
                for (int i = 0; i < CIQ; ++i)
#pragma unroll
                    for (int j = 0; j < CO_T; ++j) acc[i][j] = fmaf(xv[i], gv[j], acc[i][j]);
            }
        }
    }
    if (active) {
        const int taps = p.K * p.K * p.K;
#pragma unroll
        for (int i = 0; i < CIQ; ++i) {
            float* o = dw + (((size_t)g * taps + tap) * p.Cin + ci0 + ciq * CIQ + i) * p.Cout + coc * CO_T;
#pragma unroll
            for (int j = 0; j < CO_T; ++j) atomicAdd(o + j, acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------ wgrad, 3x3x3 stride 2, bf16 (round 2)
// The two longest launches of the RFNet step were the stride-2 data and weight gradients (ncu launch list: 319 / 309 us on the
// 8 -> 16 encoder conv).  The tiled kernel above stages a tile, waits, accumulates, and pays two integer divisions per voxel and
// accumulator row; this variant keeps its work split (thread = (tap, 4 input channels) x 16 output channels for a slice of the
// tile's voxels) but
//   * copies the NEXT tile's input halo and dy tile RAW (bf16) by 16-byte cp.async into the other half of a double buffer while the
//     current tile is consumed (out-of-volume elements are the copy's zero fill, reflection is resolved in the source address);
//   * has compile-time tile extents (2 x 4 x 8 outputs, 5 x 9 x 17 inputs), so voxel -> halo offsets are shifts and adds;
//   * converts the 64 x 16 dy values of a tile to fp32 once per tile instead of once per (thread, voxel).
// Single source (the encoder's down-sampling convs), CIC in {8, 16} input channels per CTA, 16 output channels per CTA.
constexpr int kS2Td = 2, kS2Th = 4, kS2Tw = 8, kS2Tv = kS2Td * kS2Th * kS2Tw;              // output tile
constexpr int kS2Hd = 5, kS2Hh = 9, kS2Hw = 17, kS2Halo = kS2Hd * kS2Hh * kS2Hw;          // input halo = (t - 1) * 2 + 3

template <int CIC>
__device__ __forceinline__ void s2_stage(const ConvK& p, const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ xraw,
                                         bf16* __restrict__ dyraw, int n, int od0, int oh0, int ow0, int ci0, int co0) {
    constexpr int CPV = CIC / 8;                                  // 16-byte copies per halo voxel
    for (int idx = threadIdx.x; idx < kS2Halo * CPV; idx += 256) {
        const int hv = idx / CPV, part = idx - hv * CPV;
        const int hw_i = hv % kS2Hw, r2 = hv / kS2Hw;
        const int hh_i = r2 % kS2Hh, hd_i = r2 / kS2Hh;
        int id = od0 * 2 - 1 + hd_i, ih = oh0 * 2 - 1 + hh_i, iw = ow0 * 2 - 1 + hw_i;
        if (p.reflect) { id = reflect_idx(id, p.Di); ih = reflect_idx(ih, p.Hi); iw = reflect_idx(iw, p.Wi); }
        const bool ok = id >= 0 && id < p.Di && ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi;
        const bf16* src = ok ? x + ((((size_t)n * p.Di + id) * p.Hi + ih) * p.Wi + iw) * p.C0 + ci0 + part * 8 : x;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(xraw + (size_t)hv * CIC + part * 8);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
    }
    for (int idx = threadIdx.x; idx < kS2Tv * 2; idx += 256) {
        const int v = idx >> 1, part = idx & 1;
        const int od = od0 + (v >> 5), oh = oh0 + ((v >> 3) & 3), ow = ow0 + (v & 7);
        const bool ok = od < p.Do && oh < p.Ho && ow < p.Wo;
        const bf16* src = ok ? dy + ((((size_t)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.Cout + co0 + part * 8 : dy;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dyraw + v * 16 + part * 8);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
    }
}

template <int CIC>
__global__ void __launch_bounds__(256, 2) conv_wgrad_s2_kernel(ConvK p, int tiles_d, int tiles_h, int tiles_w, int n_cic,
                                                               const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw) {
    extern __shared__ __align__(16) uint8_t s2_smem[];
    constexpr int NQ = CIC / 4, NITEMS = 27 * NQ, SLICES = 256 / NITEMS;
    constexpr size_t XBYTES = (size_t)kS2Halo * CIC * 2, DBYTES = (size_t)kS2Tv * 16 * 2;
    auto xraw = [&](int b) { return reinterpret_cast<bf16*>(s2_smem + (size_t)b * XBYTES); };
    auto dyraw = [&](int b) { return reinterpret_cast<bf16*>(s2_smem + 2 * XBYTES + (size_t)b * DBYTES); };
    float* dyf = reinterpret_cast<float*>(s2_smem + 2 * XBYTES + 2 * DBYTES);               // [64][16] fp32
    const int g = blockIdx.z;
    const int cic_i = blockIdx.y % n_cic, coc = blockIdx.y / n_cic;
    const int ci0 = cic_i * CIC, co0 = coc * 16;
    const int tid = threadIdx.x;
    const int item = tid % NITEMS, slice = tid / NITEMS;
    const bool active = slice < SLICES;
    const int tap = item / NQ, ciq = item - tap * NQ;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    const int tapoff = ((kd * kS2Hh + kh) * kS2Hw + kw) * CIC + ciq * 4;                  // element offset inside the halo tile

    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;

    const int tiles_ps = tiles_d * tiles_h * tiles_w;
    const long long ntiles = (long long)p.npg * tiles_ps;
    auto coords = [&](long long t, int& n, int& od0, int& oh0, int& ow0) {
        n = g * p.npg + (int)(t / tiles_ps);
        int r = (int)(t % tiles_ps);
        const int tw_i = r % tiles_w; r /= tiles_w;
        od0 = (r / tiles_h) * kS2Td; oh0 = (r % tiles_h) * kS2Th; ow0 = tw_i * kS2Tw;
    };
    int buf = 0;
    if ((long long)blockIdx.x < ntiles) {
        int n, od0, oh0, ow0;
        coords(blockIdx.x, n, od0, oh0, ow0);
        s2_stage<CIC>(p, x, dy, xraw(0), dyraw(0), n, od0, oh0, ow0, ci0, co0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (t + gridDim.x < ntiles) {
            int n, od0, oh0, ow0;
            coords(t + gridDim.x, n, od0, oh0, ow0);
            s2_stage<CIC>(p, x, dy, xraw(buf ^ 1), dyraw(buf ^ 1), n, od0, oh0, ow0, ci0, co0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");          // the current tile has landed
        __syncthreads();
        {   // dy tile -> fp32, once per tile (64 x 16 values, 4 per thread)
            const uint2 u = *reinterpret_cast<const uint2*>(dyraw(buf) + tid * 4);
            float4 f;
            bf2_unpack(u.x, f.x, f.y); bf2_unpack(u.y, f.z, f.w);
            reinterpret_cast<float4*>(dyf)[tid] = f;
        }
        __syncthreads();
        if (active) {
            const bf16* xb = xraw(buf) + tapoff;
#pragma unroll 2
            for (int v = slice; v < kS2Tv; v += SLICES) {
                const int hv = (((v >> 5) * 2) * kS2Hh + ((v >> 3) & 3) * 2) * kS2Hw + (v & 7) * 2;
                const uint2 xu = *reinterpret_cast<const uint2*>(xb + hv * CIC);
                float xv[4];
                bf2_unpack(xu.x, xv[0], xv[1]); bf2_unpack(xu.y, xv[2], xv[3]);
                float gv[16];
                lds_vec<16>(dyf + v * 16, gv);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(xv[i], gv[j], acc[i][j]);
            }
        }
        __syncthreads();                                                // both halves of the next iteration's writes are free now
        buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // the SLICES partial sums of an item are combined in shared memory first (slice after slice into one [item][4][16] buffer that
    // reuses the tile storage), so a CTA sends NITEMS x 64 atomics instead of 256 x 64 — with ~300 persistent CTAs adding into the
    // same 27 x CIC x 16 addresses, the atomics were a visible tail of the launch
    float* red = reinterpret_cast<float*>(s2_smem);
    static_assert((size_t)NITEMS * 64 * sizeof(float) <= 2 * XBYTES + 2 * DBYTES + (size_t)kS2Tv * 16 * sizeof(float), "reduction buffer");
    for (int sl = 0; sl < SLICES; ++sl) {
        if (active && slice == sl) {
            float4* r4 = reinterpret_cast<float4*>(red + (size_t)item * 64);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 a = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
                    if (sl > 0) { const float4 b = r4[i * 4 + j / 4]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
                    r4[i * 4 + j / 4] = a;
                }
        }
        __syncthreads();
    }
    for (int e = tid; e < NITEMS * 64; e += 256) {
        const int it2 = e >> 6, i = (e >> 4) & 3, j = e & 15;
        const int tap2 = it2 / NQ, ciq2 = it2 - tap2 * NQ;
        atomicAdd(dw + (((size_t)g * 27 + tap2) * p.Cin + ci0 + ciq2 * 4 + i) * p.Cout + co0 + j, red[e]);
    }
}

// ------------------------------------------------------------------------------------ wgrad, 1x1x1
// dw[ci][co] = sum_v x[v][ci] * dy[v][co]: every thread streams a strided set of voxels and keeps a CI_B x CO_B
// register tile; one shuffle + smem reduction per block, then fp32 atomics.  grid = (voxel blocks, ci/co tiles, groups)
template <typename T, int CI_B, int CO_B>
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(ConvK p, const T* __restrict__ x0, const T* __restrict__ x1,
                                                          const T* __restrict__ dy, float* __restrict__ dw) {
    __shared__ float red[8][CI_B * CO_B];
    const int g = blockIdx.z;
    const int n_ci = p.Cin / CI_B;
    const int ci0 = (blockIdx.y % n_ci) * CI_B, co0 = (blockIdx.y / n_ci) * CO_B;
    const bool from1 = ci0 >= p.C0;
    const T* xs = from1 ? x1 : x0;
    const int cs = from1 ? p.C1 : p.C0, coff = from1 ? ci0 - p.C0 : ci0;
    float acc[CI_B][CO_B];
#pragma unroll
    for (int i = 0; i < CI_B; ++i)
#pragma unroll
        for (int j = 0; j < CO_B; ++j) acc[i][j] = 0.f;
    const long long v_begin = (long long)g * p.npg * p.Vo, v_end = v_begin + (long long)p.npg * p.Vo;
    for (long long v = v_begin + (long long)blockIdx.x * 256 + threadIdx.x; v < v_end; v += (long long)gridDim.x * 256) {
        float xv[CI_B], gv[CO_B];
        VecIO<T, CI_B>::load(xs + v * cs + coff, xv);
        VecIO<T, CO_B>::load(dy + v * p.Cout + co0, gv);
#pragma unroll
        for (int i = 0; i < CI_B; ++i)
#pragma unroll
            for (int j = 0; j < CO_B; ++j) acc[i][j] = fmaf(xv[i], gv[j], acc[i][j]);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < CI_B; ++i)
#pragma unroll
        for (int j = 0; j < CO_B; ++j) {
            const float s = warp_sum(acc[i][j]);
            if (lane == 0) red[wid][i * CO_B + j] = s;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < CI_B * CO_B; e += 256) {
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) s += red[w8][e];
        const int i = e / CO_B, j = e % CO_B;
        atomicAdd(dw + ((size_t)g * p.Cin + ci0 + i) * p.Cout + co0 + j, s);
    }
}

// ------------------------------------------------------------------------------------ host side
int fill(const pb_conv_desc* d, ConvK& k) {
    if (!d) return -1;
    if (d->ksize != 1 && d->ksize != 3) return -1;
    if (d->stride != 1 && d->stride != 2) return -1;
    if (d->groups < 1 || d->n % d->groups) return -1;
    if (d->c0 < 1 || d->c1 < 0 || d->cout < 1) return -1;
    k.N = d->n; k.Di = d->di; k.Hi = d->hi; k.Wi = d->wi; k.Do = d->dout; k.Ho = d->ho; k.Wo = d->wo;
    k.C0 = d->c0; k.C1 = d->c1; k.Cin = d->c0 + d->c1; k.Cout = d->cout; k.K = d->ksize; k.S = d->stride;
    k.pad = d->ksize / 2; k.reflect = (d->pad_mode == PB_PAD_REFLECT && d->ksize == 3) ? 1 : 0;
    k.groups = d->groups; k.npg = d->n / d->groups;
    k.Vi = (long long)d->di * d->hi * d->wi; k.Vo = (long long)d->dout * d->ho * d->wo;
    // output size must match the conv arithmetic
    auto osz = [&](int i) { return (i + 2 * k.pad - k.K) / k.S + 1; };
    if (osz(k.Di) != k.Do || osz(k.Hi) != k.Ho || osz(k.Wi) != k.Wo) return -1;
    if (k.reflect && (k.Di < 2 || k.Hi < 2 || k.Wi < 2)) return -1;
    return 0;
}

constexpr size_t kSmemBudget = 200 * 1024;   // dynamic shared memory the FFMA kernels may ask for (227 KB per CTA on sm_100a)

int chunk_of(int c, int cap) {           // largest power of two <= cap dividing c
    int v = cap;
    while (v > 1 && c % v) v >>= 1;
    return v;
}

template <typename K> int set_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { pb_set_error("cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e)); return PB_ECUDA; }
    }
    return 0;
}

template <typename T, int CI_V, int CO_T>
int launch_fwd(const ConvK& k, const void* x0, const void* x1, const float* w, const float* bias, void* y, double* stats,
               cudaStream_t st) {
    const size_t smem = (size_t)k.K * k.K * k.K * k.Cin * CO_T * sizeof(float);
    auto kern = conv_fwd_kernel<T, CI_V, CO_T>;
    if (int e = set_smem(kern, smem)) return e;
    long long bx = (k.Vo + 127) / 128;
    const long long cap = (148LL * 8 + (long long)(k.Cout / CO_T) * k.N - 1) / ((long long)(k.Cout / CO_T) * k.N);
    if (smem > 16 * 1024 && bx > cap) bx = cap;                 // wide layers: amortise the weight staging over many voxels
    dim3 grid((unsigned)bx, k.Cout / CO_T, k.N);
    kern<<<grid, 128, smem, st>>>(k, (const T*)x0, (const T*)x1, w, bias, (T*)y, stats);
    return 0;
}

template <typename T, int CI_V>
int dispatch_fwd_co(int co_t, const ConvK& k, const void* x0, const void* x1, const float* w, const float* bias, void* y,
                    double* stats, cudaStream_t st) {
    switch (co_t) {
        case 16: return launch_fwd<T, CI_V, 16>(k, x0, x1, w, bias, y, stats, st);
        case 8:  return launch_fwd<T, CI_V, 8>(k, x0, x1, w, bias, y, stats, st);
        case 4:  return launch_fwd<T, CI_V, 4>(k, x0, x1, w, bias, y, stats, st);
        case 2:  return launch_fwd<T, CI_V, 2>(k, x0, x1, w, bias, y, stats, st);
        default: return launch_fwd<T, CI_V, 1>(k, x0, x1, w, bias, y, stats, st);
    }
}

template <typename T, int CI_V, int CO_T>
int launch_pw(const PwK& p, const void* x0, const void* x1, const float* w, const float* bias, void* y0, void* y1, double* stats,
              cudaStream_t st) {
    constexpr int UNR = (CO_T >= 16 || CI_V * CO_T > 64) ? 2 : 4;
    const size_t smem = (size_t)p.Cin * CO_T * sizeof(float);
    auto kern = pw_conv_kernel<T, CI_V, CO_T, UNR>;
    if (int e = set_smem(kern, smem)) return e;
    const int chunks = p.Cout / CO_T;
    long long bx = (p.V + 256LL * UNR - 1) / (256LL * UNR);
    const long long cap = (148LL * 8 + (long long)p.N * chunks - 1) / ((long long)p.N * chunks);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    kern<<<dim3((unsigned)bx, chunks, p.N), 256, smem, st>>>(p, (const T*)x0, (const T*)x1, w, bias, (T*)y0, (T*)y1, stats);
    return 0;
}

template <typename T, int CI_V>
int dispatch_pw_co(int co_t, const PwK& p, const void* x0, const void* x1, const float* w, const float* bias, void* y0, void* y1,
                   double* stats, cudaStream_t st) {
    switch (co_t) {
        case 16: return launch_pw<T, CI_V, 16>(p, x0, x1, w, bias, y0, y1, stats, st);
        case 8:  return launch_pw<T, CI_V, 8>(p, x0, x1, w, bias, y0, y1, stats, st);
        case 4:  return launch_pw<T, CI_V, 4>(p, x0, x1, w, bias, y0, y1, stats, st);
        case 2:  return launch_pw<T, CI_V, 2>(p, x0, x1, w, bias, y0, y1, stats, st);
        default: return launch_pw<T, CI_V, 1>(p, x0, x1, w, bias, y0, y1, stats, st);
    }
}

template <int CO_T, int NCH0, int NCH1>
int launch_pw2(const PwK& p, const void* x0, const void* x1, const float* w, const float* bias, void* y0, void* y1, double* stats,
               cudaStream_t st) {
    constexpr int NCH = NCH0 + NCH1;
    // register budgets checked with ptxas -v: no spills at 80 registers (3 CTAs/SM) for CO_T <= 8, 128 (2 CTAs/SM) for CO_T = 16
    constexpr int UNR = (CO_T >= 8 || NCH >= 4) ? 2 : 4;
    constexpr int MINB = CO_T >= 16 ? 2 : 3;
    const size_t smem = (size_t)p.Cin * CO_T * sizeof(float);
    auto kern = pw_conv2_kernel<bf16, CO_T, NCH0, NCH1, UNR, MINB>;
    if (int e = set_smem(kern, smem)) return e;
    const int chunks = p.Cout / CO_T;
    long long bx = (p.V + 256LL * UNR - 1) / (256LL * UNR);
    const long long cap = (148LL * MINB * 2) / ((long long)p.N * chunks);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    kern<<<dim3((unsigned)bx, chunks, p.N), 256, smem, st>>>(p, (const bf16*)x0, (const bf16*)x1, w, bias, (bf16*)y0, (bf16*)y1, stats);
    return 0;
}

// route the bf16 classes pw_conv2_kernel was instantiated for to it (returns -1 when the class is not covered)
template <int CO_T>
int dispatch_pw2(const PwK& p, const void* x0, const void* x1, const float* w, const float* bias, void* y0, void* y1, double* stats,
                 cudaStream_t st) {
    const int a = p.C0 / 8, b = p.C1 / 8;
    if (a == 1 && b == 0) return launch_pw2<CO_T, 1, 0>(p, x0, x1, w, bias, y0, y1, stats, st);
    if (a == 2 && b == 0) return launch_pw2<CO_T, 2, 0>(p, x0, x1, w, bias, y0, y1, stats, st);
    if (a == 4 && b == 0) return launch_pw2<CO_T, 4, 0>(p, x0, x1, w, bias, y0, y1, stats, st);
    if (a == 1 && b == 1) return launch_pw2<CO_T, 1, 1>(p, x0, x1, w, bias, y0, y1, stats, st);
    if (a == 2 && b == 2) return launch_pw2<CO_T, 2, 2>(p, x0, x1, w, bias, y0, y1, stats, st);
    return -1;
}

bool pw2_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PB_PW2"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// p.C0|C1 = input split, p.CO0|CO1 = output split
template <typename T>
int dispatch_pw(const PwK& p, const void* x0, const void* x1, const float* w, const float* bias, void* y0, void* y1,
                double* stats, cudaStream_t st) {
    int ci_v = chunk_of(p.C0, 8);
    if (p.C1) ci_v = chunk_of(p.C1, ci_v);
    int co_t = chunk_of(p.CO0, 16);
    if (p.CO1) co_t = chunk_of(p.CO1, co_t);
    // small volumes (deep levels): narrower Cout tiles give more CTAs; x is re-read from L2, which is cheap at these sizes
    while (co_t > 4 && ((p.V + 511) / 512) * (p.Cout / co_t) * p.N < 148) co_t >>= 1;
    if (sizeof(T) == 2 && ci_v == 8 && pw2_enabled()) {
        int e = -1;
        if (co_t == 16) e = dispatch_pw2<16>(p, x0, x1, w, bias, y0, y1, stats, st);
        else if (co_t == 8) e = dispatch_pw2<8>(p, x0, x1, w, bias, y0, y1, stats, st);
        else if (co_t == 4) e = dispatch_pw2<4>(p, x0, x1, w, bias, y0, y1, stats, st);
        if (e >= 0) return e;
    }
    switch (ci_v) {
        case 8:  return dispatch_pw_co<T, 8>(co_t, p, x0, x1, w, bias, y0, y1, stats, st);
        case 4:  return dispatch_pw_co<T, 4>(co_t, p, x0, x1, w, bias, y0, y1, stats, st);
        case 2:  return dispatch_pw_co<T, 2>(co_t, p, x0, x1, w, bias, y0, y1, stats, st);
        default: return dispatch_pw_co<T, 1>(co_t, p, x0, x1, w, bias, y0, y1, stats, st);
    }
}

template <typename T>
int dispatch_fwd(const ConvK& k, const void* x0, const void* x1, const float* w, const float* bias, void* y, double* stats,
                 cudaStream_t st) {
    if (k.K == 1 && k.S == 1) {
        PwK p{k.N, k.C0, k.C1, k.Cin, k.Cout, 0, k.Cout, k.npg, k.Vo};
        return dispatch_pw<T>(p, x0, x1, w, bias, y, nullptr, stats, st);
    }
    int ci_v = chunk_of(k.C0, 8);
    if (k.C1) ci_v = chunk_of(k.C1, ci_v);
    int co_t = chunk_of(k.Cout, 16);
    // the CTA's weight slice (taps x Cin x co_t floats) lives in shared memory: narrow the tile for wide layers
    while (co_t > 1 && (size_t)k.K * k.K * k.K * k.Cin * co_t * sizeof(float) > kSmemBudget) co_t >>= 1;
    switch (ci_v) {
        case 8:  return dispatch_fwd_co<T, 8>(co_t, k, x0, x1, w, bias, y, stats, st);
        case 4:  return dispatch_fwd_co<T, 4>(co_t, k, x0, x1, w, bias, y, stats, st);
        case 2:  return dispatch_fwd_co<T, 2>(co_t, k, x0, x1, w, bias, y, stats, st);
        default: return dispatch_fwd_co<T, 1>(co_t, k, x0, x1, w, bias, y, stats, st);
    }
}

template <typename T, int CO_V, int CI_T>
int launch_dgrad(const ConvK& k, const void* dy, const float* wt, void* dx0, void* dx1, cudaStream_t st, bool mirror_only) {
    const size_t smem = (size_t)k.K * k.K * k.K * k.Cout * CI_T * sizeof(float);
    if (!mirror_only && k.Do <= 256 && k.Ho <= 256 && k.Wo <= 256) {
        auto kp = conv_dgrad_pairs_kernel<T, CO_V, CI_T>;
        if (int e = set_smem(kp, smem)) return e;
        long long bx = (k.Vi + 127) / 128;
        // every CTA stages its weight slice first: more CTAs than fit the SMs at once only multiply that staging (the coarse
        // levels moved 25x more weight bytes than activation bytes: 0.27 ms for the 10^3 layer)
        long long resident = smem ? (long long)(220 * 1024 / smem) : 16;
        if (resident > 6) resident = 6;
        if (resident < 1) resident = 1;
        const long long cap = (148LL * resident + (long long)(k.Cin / CI_T) * k.N - 1) / ((long long)(k.Cin / CI_T) * k.N);
        if (bx > cap) bx = cap;
        dim3 gridp((unsigned)bx, k.Cin / CI_T, k.N);
        // (256 threads per CTA for the layers whose weight slice leaves one or two CTAs per SM were measured: 16 -> 32 at 20^3 0.168 ->
        // 0.201 ms, 32 -> 64 at 10^3 unchanged — the same grid then has half the voxels per thread behind the same staging prologue)
        kp<<<gridp, 128, smem, st>>>(k, (const T*)dy, wt, (T*)dx0, (T*)dx1);
        return 0;
    }
    auto kern = mirror_only ? conv_dgrad_kernel<T, CO_V, CI_T, true> : conv_dgrad_kernel<T, CO_V, CI_T, false>;
    if (int e = set_smem(kern, smem)) return e;
    long long work = k.Vi;
    if (mirror_only) work = 2LL * k.Hi * k.Wi + (long long)(k.Di - 2) * 2 * k.Wi + (long long)(k.Di - 2) * (k.Hi - 2) * 2;
    dim3 grid((unsigned)((work + 127) / 128), k.Cin / CI_T, k.N);
    kern<<<grid, 128, smem, st>>>(k, (const T*)dy, wt, (T*)dx0, (T*)dx1);
    return 0;
}

template <typename T, int CO_V>
int dispatch_dgrad_ci(int ci_t, const ConvK& k, const void* dy, const float* wt, void* dx0, void* dx1, cudaStream_t st, bool mo) {
    switch (ci_t) {
        case 16: return launch_dgrad<T, CO_V, 16>(k, dy, wt, dx0, dx1, st, mo);
        case 8:  return launch_dgrad<T, CO_V, 8>(k, dy, wt, dx0, dx1, st, mo);
        case 4:  return launch_dgrad<T, CO_V, 4>(k, dy, wt, dx0, dx1, st, mo);
        case 2:  return launch_dgrad<T, CO_V, 2>(k, dy, wt, dx0, dx1, st, mo);
        default: return launch_dgrad<T, CO_V, 1>(k, dy, wt, dx0, dx1, st, mo);
    }
}

template <typename T>
int dispatch_dgrad(const ConvK& k, const void* dy, const float* wt, void* dx0, void* dx1, cudaStream_t st, bool mo = false) {
    if (k.K == 1 && k.S == 1 && !mo) {
        // dx[v][ci] = sum_co dy[v][co] wt[co][ci]: the pointwise kernel with the roles of the channels swapped
        PwK p{k.N, k.Cout, 0, k.Cout, k.C0, k.C1, k.Cin, k.npg, k.Vi};
        return dispatch_pw<T>(p, dy, nullptr, wt, nullptr, dx0, dx1, nullptr, st);
    }
    const int co_v = chunk_of(k.Cout, 8);
    int ci_t = chunk_of(k.C0, 16);
    if (k.C1) ci_t = chunk_of(k.C1, ci_t);
    while (ci_t > 1 && (size_t)k.K * k.K * k.K * k.Cout * ci_t * sizeof(float) > kSmemBudget) ci_t >>= 1;
    switch (co_v) {
        case 8:  return dispatch_dgrad_ci<T, 8>(ci_t, k, dy, wt, dx0, dx1, st, mo);
        case 4:  return dispatch_dgrad_ci<T, 4>(ci_t, k, dy, wt, dx0, dx1, st, mo);
        case 2:  return dispatch_dgrad_ci<T, 2>(ci_t, k, dy, wt, dx0, dx1, st, mo);
        default: return dispatch_dgrad_ci<T, 1>(ci_t, k, dy, wt, dx0, dx1, st, mo);
    }
}

template <typename T, int CIQ, int CO_T, int LV>
int launch_wgrad(const ConvK& k, WgK q, const void* x0, const void* x1, const void* dy, float* dw, cudaStream_t st) {
    const size_t smem = ((((size_t)q.hd * q.hh * q.hw * q.xs + 3) & ~(size_t)3) + (size_t)q.td * q.th * q.tw * CO_T) * sizeof(float);
    auto kern = conv_wgrad_kernel<T, CIQ, CO_T, LV>;
    if (int e = set_smem(kern, smem)) return e;
    const long long ntiles = (long long)k.npg * q.tiles_d * q.tiles_h * q.tiles_w;
    const int pairs = q.n_cic * q.n_coc;
    // persistent CTAs: exactly as many as are resident at once (registers / shared memory allow 2-3 per SM); 1.5 waves of
    // 444 CTAs left every SM half idle for the last third of the kernel
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem) != cudaSuccess || occ < 1) occ = 2;
    long long nblk = (148LL * occ) / ((long long)pairs * k.groups);
    if (nblk < 1) nblk = 1;
    if (nblk > ntiles) nblk = ntiles;
    dim3 grid((unsigned)nblk, pairs, k.groups);
    kern<<<grid, 256, smem, st>>>(k, q, (const T*)x0, (const T*)x1, (const T*)dy, dw);
    return 0;
}

template <typename T, int CIQ, int LV>
int dispatch_wgrad_co(int co_t, const ConvK& k, const WgK& q, const void* x0, const void* x1, const void* dy, float* dw,
                      cudaStream_t st) {
    switch (co_t) {
        case 16: return launch_wgrad<T, CIQ, 16, LV>(k, q, x0, x1, dy, dw, st);
        case 8:  return launch_wgrad<T, CIQ, 8, LV>(k, q, x0, x1, dy, dw, st);
        case 4:  return launch_wgrad<T, CIQ, 4, LV>(k, q, x0, x1, dy, dw, st);
        case 2:  return launch_wgrad<T, CIQ, 2, LV>(k, q, x0, x1, dy, dw, st);
        default: return launch_wgrad<T, CIQ, 1, LV>(k, q, x0, x1, dy, dw, st);
    }
}

// Row-cooperative variant: the L = Cin/CI_B lanes of a lane group share one voxel, each owning one CI_B-channel chunk
// of its x row (so a warp reads whole rows, every sector fully used) and all reading the same CO_B dy values (one
// broadcast sector).  x and dy are read exactly once per Cout tile.  Accumulators are combined by xor-shuffles across
// the lane groups, then across the 8 warps through shared memory, then one atomicAdd per element and CTA.
template <typename T, int CI_B, int CO_B, int UNR>
__global__ void __launch_bounds__(256) conv1_wgrad_rows_kernel(ConvK p, int L, const T* __restrict__ x0, const T* __restrict__ x1,
                                                               const T* __restrict__ dy, float* __restrict__ dw) {
    extern __shared__ float red[];                            // [8][Cin * CO_B]
    const int g = blockIdx.z, co0 = blockIdx.y * CO_B;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sub = lane % L, vl = lane / L, vpw = 32 / L;
    const int ci0 = sub * CI_B;
    const bool from1 = ci0 >= p.C0;
    const T* xs = from1 ? x1 : x0;
    const int cs = from1 ? p.C1 : p.C0, coff = from1 ? ci0 - p.C0 : ci0;
    float acc[CI_B][CO_B];
#pragma unroll
    for (int i = 0; i < CI_B; ++i)
#pragma unroll
        for (int j = 0; j < CO_B; ++j) acc[i][j] = 0.f;
    const long long v_begin = (long long)g * p.npg * p.Vo, v_end = v_begin + (long long)p.npg * p.Vo;
    const long long stride = (long long)gridDim.x * 8 * vpw;
    for (long long v0 = v_begin + ((long long)blockIdx.x * 8 + wid) * vpw + vl; v0 < v_end; v0 += stride * UNR) {
        float xv[UNR][CI_B], gv[UNR][CO_B];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long v = v0 + u * stride;
            if (v < v_end) {
                VecIO<T, CI_B>::load(xs + v * cs + coff, xv[u]);
                VecIO<T, CO_B>::load(dy + v * p.Cout + co0, gv[u]);
            } else {
#pragma unroll
                for (int i = 0; i < CI_B; ++i) xv[u][i] = 0.f;
#pragma unroll
                for (int j = 0; j < CO_B; ++j) gv[u][j] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int i = 0; i < CI_B; ++i)
#pragma unroll
                for (int j = 0; j < CO_B; ++j) acc[i][j] = fmaf(xv[u][i], gv[u][j], acc[i][j]);
    }
    const int row = p.Cin * CO_B;
#pragma unroll
    for (int i = 0; i < CI_B; ++i)
#pragma unroll
        for (int j = 0; j < CO_B; ++j) {
            float s = acc[i][j];
            for (int off = L; off < 32; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane < L) red[wid * row + (ci0 + i) * CO_B + j] = s;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < row; e += 256) {
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * row + e];
        const int ci = e / CO_B, j = e % CO_B;
        atomicAdd(dw + ((size_t)g * p.Cin + ci) * p.Cout + co0 + j, s);
    }
}

template <typename T, int CI_B, int CO_B>
int launch_wgrad1_rows(const ConvK& k, const void* x0, const void* x1, const void* dy, float* dw, cudaStream_t st) {
    constexpr int UNR = CI_B * CO_B >= 64 ? 2 : 4;
    const int L = k.Cin / CI_B;
    const int tiles = k.Cout / CO_B;
    const size_t smem = (size_t)8 * k.Cin * CO_B * sizeof(float);
    auto kern = conv1_wgrad_rows_kernel<T, CI_B, CO_B, UNR>;
    if (int e = set_smem(kern, smem)) return e;
    const long long vox = (long long)k.npg * k.Vo;
    long long nblk = (148LL * 6 + (long long)tiles * k.groups - 1) / ((long long)tiles * k.groups);
    const long long need = (vox * L + 256 * 8 - 1) / (256 * 8);         // >= 8 voxels per lane group
    if (nblk > need) nblk = need;
    if (nblk < 1) nblk = 1;
    kern<<<dim3((unsigned)nblk, tiles, k.groups), 256, smem, st>>>(k, L, (const T*)x0, (const T*)x1, (const T*)dy, dw);
    return 0;
}

template <typename T, int CI_B, int CO_B>
int launch_wgrad1(const ConvK& k, const void* x0, const void* x1, const void* dy, float* dw, cudaStream_t st) {
    const int tiles = (k.Cin / CI_B) * (k.Cout / CO_B);
    const long long vox = (long long)k.npg * k.Vo;
    long long nblk = (148LL * 8 + (long long)tiles * k.groups - 1) / ((long long)tiles * k.groups);
    const long long need = (vox + 256 * 4 - 1) / (256 * 4);            // >= 4 voxels per thread
    if (nblk > need) nblk = need;
    if (nblk < 1) nblk = 1;
    conv1_wgrad_kernel<T, CI_B, CO_B><<<dim3((unsigned)nblk, tiles, k.groups), 256, 0, st>>>(k, (const T*)x0, (const T*)x1,
                                                                                            (const T*)dy, dw);
    return 0;
}

template <typename T>
int dispatch_wgrad1(const ConvK& k, const void* x0, const void* x1, const void* dy, float* dw, cudaStream_t st) {
    {
        // row-cooperative kernel: needs Cin / CI_B to be a power of two <= 32 and the reduction buffer to fit; measured
        // faster than the tile kernel below only for rows of >= 4 chunks (c32->8 80^3: 0.27 -> 0.23 ms)
        const int co_r = chunk_of(k.Cout, 8);
        int ci_r = chunk_of(k.C0, 8);
        if (k.C1) ci_r = chunk_of(k.C1, ci_r);
        while (k.Cin / ci_r > 32 && ci_r < 8) ci_r <<= 1;
        const int L = k.Cin / ci_r;
        const bool pow2 = L >= 1 && L <= 32 && (L & (L - 1)) == 0 && L * ci_r == k.Cin && k.C0 % ci_r == 0 && k.C1 % ci_r == 0;
        if (pow2 && L >= 4 && (size_t)8 * k.Cin * co_r * sizeof(float) <= 96 * 1024 && (ci_r == 8 || ci_r == 4 || ci_r == 2 || ci_r == 1)) {
#define W1R(CI, CO) return launch_wgrad1_rows<T, CI, CO>(k, x0, x1, dy, dw, st)
            switch (co_r) {
                case 8:  switch (ci_r) { case 8: W1R(8, 8); case 4: W1R(4, 8); case 2: W1R(2, 8); default: W1R(1, 8); }
                case 4:  switch (ci_r) { case 8: W1R(8, 4); case 4: W1R(4, 4); case 2: W1R(2, 4); default: W1R(1, 4); }
                case 2:  switch (ci_r) { case 8: W1R(8, 2); case 4: W1R(4, 2); case 2: W1R(2, 2); default: W1R(1, 2); }
                default: switch (ci_r) { case 8: W1R(8, 1); case 4: W1R(4, 1); case 2: W1R(2, 1); default: W1R(1, 1); }
            }
#undef W1R
        }
    }
    const int co_b = chunk_of(k.Cout, 16);
    int ci_b = chunk_of(k.C0, 64 / co_b > 8 ? 8 : 64 / co_b);
    if (k.C1) ci_b = chunk_of(k.C1, ci_b);
#define W1(CI, CO) return launch_wgrad1<T, CI, CO>(k, x0, x1, dy, dw, st)
    switch (co_b) {
        case 16: switch (ci_b) { case 4: W1(4, 16); case 2: W1(2, 16); default: W1(1, 16); }
        case 8:  switch (ci_b) { case 8: W1(8, 8); case 4: W1(4, 8); case 2: W1(2, 8); default: W1(1, 8); }
        case 4:  switch (ci_b) { case 8: W1(8, 4); case 4: W1(4, 4); case 2: W1(2, 4); default: W1(1, 4); }
        case 2:  switch (ci_b) { case 8: W1(8, 2); case 4: W1(4, 2); case 2: W1(2, 2); default: W1(1, 2); }
        default: switch (ci_b) { case 8: W1(8, 1); case 4: W1(4, 1); case 2: W1(2, 1); default: W1(1, 1); }
    }
#undef W1
}

template <int CIC>
int launch_wgrad_s2(const ConvK& k, const void* x, const void* dy, float* dw, cudaStream_t st) {
    const int tiles_d = (k.Do + kS2Td - 1) / kS2Td, tiles_h = (k.Ho + kS2Th - 1) / kS2Th, tiles_w = (k.Wo + kS2Tw - 1) / kS2Tw;
    const size_t smem = 2 * ((size_t)kS2Halo * CIC * 2 + (size_t)kS2Tv * 16 * 2) + (size_t)kS2Tv * 16 * sizeof(float);
    auto kern = conv_wgrad_s2_kernel<CIC>;
    if (int e = set_smem(kern, smem)) return e;
    const int n_cic = k.Cin / CIC, pairs = n_cic * (k.Cout / 16);
    const long long ntiles = (long long)k.npg * tiles_d * tiles_h * tiles_w;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem) != cudaSuccess || occ < 1) occ = 2;
    long long nblk = (148LL * occ) / ((long long)pairs * k.groups);
    if (nblk < 1) nblk = 1;
    if (nblk > ntiles) nblk = ntiles;
    kern<<<dim3((unsigned)nblk, pairs, k.groups), 256, smem, st>>>(k, tiles_d, tiles_h, tiles_w, n_cic, (const bf16*)x, (const bf16*)dy, dw);
    return 0;
}

int wgrad_s2_mode() {          // PB_WGRAD_S2=0: the generic tiled kernel for the stride-2 weight gradients (A/B measurements)
    static const int mode = [] { const char* e = getenv("PB_WGRAD_S2"); return e != nullptr && e[0] == '0' ? 0 : 1; }();
    return mode;
}

template <typename T>
int dispatch_wgrad(const ConvK& k, const void* x0, const void* x1, const void* dy, float* dw, cudaStream_t st) {
    if (k.K == 1) return dispatch_wgrad1<T>(k, x0, x1, dy, dw, st);
    if constexpr (sizeof(T) == 2) {
        if (k.K == 3 && k.S == 2 && k.C1 == 0 && k.Cin % 8 == 0 && k.Cout % 16 == 0 && wgrad_s2_mode()) {
            if (k.Cin % 16 == 0) return launch_wgrad_s2<16>(k, x0, dy, dw, st);
            return launch_wgrad_s2<8>(k, x0, dy, dw, st);
        }
    }
    WgK q;
    if (k.S == 1) { q.td = 4; q.th = 4; q.tw = 8; } else { q.td = 2; q.th = 4; q.tw = 8; }
    q.hd = (q.td - 1) * k.S + k.K; q.hh = (q.th - 1) * k.S + k.K; q.hw = (q.tw - 1) * k.S + k.K;
    q.tiles_d = (k.Do + q.td - 1) / q.td; q.tiles_h = (k.Ho + q.th - 1) / q.th; q.tiles_w = (k.Wo + q.tw - 1) / q.tw;
    int cic = chunk_of(k.C0, k.K == 3 ? 16 : 64);
    if (k.C1) cic = chunk_of(k.C1, cic);
    q.cic = cic;
    const int ciq = chunk_of(cic, 4);
    q.xs = cic + (cic >= 4 ? 4 : 0);                          // +4 floats: keeps 16 B alignment, spreads banks
    const int taps = k.K * k.K * k.K;
    q.nitems = taps * (cic / ciq);
    if (q.nitems > 256) { pb_set_error("wgrad: nitems %d > 256", q.nitems); return PB_EUNSUPPORTED; }
    q.slices = 256 / q.nitems;
    const int co_t = chunk_of(k.Cout, 16);
    q.n_cic = k.Cin / cic; q.n_coc = k.Cout / co_t;
    switch (ciq) {
        case 4:
            if (cic % 8 == 0) return dispatch_wgrad_co<T, 4, 8>(co_t, k, q, x0, x1, dy, dw, st);
            return dispatch_wgrad_co<T, 4, 4>(co_t, k, q, x0, x1, dy, dw, st);
        case 2:  return dispatch_wgrad_co<T, 2, 2>(co_t, k, q, x0, x1, dy, dw, st);
        default: return dispatch_wgrad_co<T, 1, 1>(co_t, k, q, x0, x1, dy, dw, st);
    }
}

}  // namespace

extern "C" int pb_conv3d_fwd(const pb_conv_desc* d, const void* x0, const void* x1, const float* w, const float* bias,
                             void* y, double* stats, pb_stream_t stream) {
    ConvK k;
    PB_CHECK_ARG(fill(d, k) == 0, "bad descriptor");
    PB_CHECK_ARG(x0 && w && y && (d->c1 == 0 || x1), "null pointer");
    int e = d->dtype == PB_BF16 ? dispatch_fwd<bf16>(k, x0, x1, w, bias, y, stats, (cudaStream_t)stream)
                                : dispatch_fwd<float>(k, x0, x1, w, bias, y, stats, (cudaStream_t)stream);
    if (e) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_conv3d_dgrad(const pb_conv_desc* d, const void* dy, const float* wt, void* dx0, void* dx1,
                               pb_stream_t stream) {
    ConvK k;
    PB_CHECK_ARG(fill(d, k) == 0, "bad descriptor");
    PB_CHECK_ARG(dy && wt && dx0 && (d->c1 == 0 || dx1), "null pointer");
    int e = d->dtype == PB_BF16 ? dispatch_dgrad<bf16>(k, dy, wt, dx0, dx1, (cudaStream_t)stream)
                                : dispatch_dgrad<float>(k, dy, wt, dx0, dx1, (cudaStream_t)stream);
    if (e) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_conv3d_dgrad_reflect_fix(const pb_conv_desc* d, const void* dy, const float* wt, void* dx0, void* dx1,
                                           pb_stream_t stream) {
    ConvK k;
    PB_CHECK_ARG(fill(d, k) == 0, "bad descriptor");
    PB_CHECK_ARG(dy && wt && dx0 && (d->c1 == 0 || dx1), "null pointer");
    if (!k.reflect) return PB_OK;
    if (k.Di < 4 || k.Hi < 4 || k.Wi < 4 || k.S != 1) { pb_set_error("reflect_fix: needs stride 1 and sizes >= 4"); return PB_EUNSUPPORTED; }
    int e = d->dtype == PB_BF16 ? dispatch_dgrad<bf16>(k, dy, wt, dx0, dx1, (cudaStream_t)stream, true)
                                : dispatch_dgrad<float>(k, dy, wt, dx0, dx1, (cudaStream_t)stream, true);
    if (e) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_conv3d_wgrad(const pb_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw,
                               pb_stream_t stream) {
    ConvK k;
    PB_CHECK_ARG(fill(d, k) == 0, "bad descriptor");
    PB_CHECK_ARG(x0 && dy && dw && (d->c1 == 0 || x1), "null pointer");
    int e = d->dtype == PB_BF16 ? dispatch_wgrad<bf16>(k, x0, x1, dy, dw, (cudaStream_t)stream)
                                : dispatch_wgrad<float>(k, x0, x1, dy, dw, (cudaStream_t)stream);
    if (e) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}
