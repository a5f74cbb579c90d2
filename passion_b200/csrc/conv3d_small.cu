// 3x3x3 stride-1 convolutions with VERY few channels (Cin, Cout <= 8, Cin*Cout <= 16): the first encoder conv
// (1 -> 8, reference models/rfnet.py:24 / mmformer.py:28) and the middle convs of the PRM embedding layers
// (C/4 -> C/4 with C/4 in {2, 4}, models/blocks.py:399-401).  These layers move a few tens of MB but ran 0.3-0.8 ms each
// on the generic FFMA kernels (one thread per voxel, 27 x Cin global loads through L1): together 2.7 ms of a 32 ms step.
//
// Here a CTA stages the input halo of a tile in shared memory once (fp32; reflect / zero padding resolved by the loader)
// and every thread produces 8 consecutive output planes of one (h, w) position, so an input value is read from shared
// memory once per (kh, kw) and used for up to 3 x Cout FMAs, and a weight vector is loaded once per 8 planes.
//   forward        y  = conv(x, w) (+ bias) (+ per-(n,c) sum / sum of squares)
//   data gradient  the SAME kernel on dy with mirrored taps / transposed channels; for reflect padding in "full" mode
//                  (output = input grown by one voxel per side) followed by small_fold_kernel (see conv3d_tc.cu)
//   weight grad.   every thread keeps the (kd x) 9 x Cin x Cout partial sums of its voxels in registers across all the
//                  tiles of a persistent CTA; one block reduction and one atomicAdd per element at the end
#include <cstdlib>
#include "common.cuh"

namespace {

template <int N> __device__ __forceinline__ void lds_vecN(const float* p, float* o) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 v = reinterpret_cast<const float4*>(p)[i];
            o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w;
        }
    } else if constexpr (N == 2) {
        const float2 v = *reinterpret_cast<const float2*>(p);
        o[0] = v.x; o[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = p[i];
    }
}

constexpr int kTD = 8;            // output planes per tile
constexpr int kTQ = 256;          // (h, w) positions per tile = threads per CTA

struct SmK {
    int N, D, H, W;               // output extent
    int Di, Hi, Wi;               // input extent (= output, or output - 2 in full mode)
    int pad;                      // 1 = same-size conv, 2 = full correlation
    int reflect;                  // reflect padding (pad 1 only)
    int npg;                      // samples per weight group
    int mirror;                   // weights are [27][Cout_conv][Cin_conv] of the forward conv, taps mirrored (data gradient)
    int rows;                     // input rows staged per tile
    int tiles_q, tiles_d;
};

// input halo of tile (n, dt, qt): planes d0-pad .. d0+kTD+1-pad, rows hb-pad .. , all columns -pad .. W-1+pad(+...)
template <typename T, int CIN>
__device__ __forceinline__ void load_tile(const SmK& p, const T* __restrict__ x, float* __restrict__ tile, int n, int d0, int hb) {
    // one warp per staged row (dz, r): the plane / row indices (and their reflection) are resolved once per row, the lanes
    // stride over the columns — no per-element divisions (they made the first version of this kernel instruction-bound)
    const int WP = p.W + 2;                                     // staged columns: output w + kw, kw in 0..2
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nrows = (kTD + 2) * p.rows;
    for (int row = wid; row < nrows; row += kTQ / 32) {
        const int dz = row / p.rows, r = row - dz * p.rows;
        int id = d0 + dz - p.pad, ih = hb + r - p.pad;
        bool ok_row;
        if (p.reflect) { id = reflect_idx(id, p.Di); ih = reflect_idx(ih, p.Hi); }
        ok_row = id >= 0 && id < p.Di && ih >= 0 && ih < p.Hi;   // false: zero padding, or planes / rows beyond the last tile
        const T* src = x + (((size_t)n * p.Di + (ok_row ? id : 0)) * p.Hi + (ok_row ? ih : 0)) * p.Wi * CIN;
        float* dst = tile + (size_t)row * WP * CIN;
        for (int c0 = 0; c0 < WP; c0 += 128) {
            float v[4][CIN];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * 32 + lane;
                int iw = c - p.pad;
                bool ok = ok_row && c < WP;
                if (p.reflect) iw = reflect_idx(iw, p.Wi); else ok = ok && iw >= 0 && iw < p.Wi;
                if (ok) VecIO<T, CIN>::load(src + (size_t)iw * CIN, v[u]);
                else {
#pragma unroll
                    for (int k = 0; k < CIN; ++k) v[u][k] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * 32 + lane;
                if (c < WP) {
#pragma unroll
                    for (int k = 0; k < CIN; ++k) dst[(size_t)c * CIN + k] = v[u][k];
                }
            }
        }
    }
}

template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(kTQ) conv3_small_fwd_kernel(SmK p, const T* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, T* __restrict__ y,
                                                              double* __restrict__ stats) {
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                                          // [27][CIN][COUT]
    float* tile = smem + 27 * CIN * COUT;                       // [kTD+2][rows][W+2][CIN]
    __shared__ float red[8][2 * COUT];
    const int qt = blockIdx.x % p.tiles_q, dt = blockIdx.x / p.tiles_q, n = blockIdx.y, g = n / p.npg;
    const float* wg = w + (size_t)g * 27 * CIN * COUT;
    for (int i = threadIdx.x; i < 27 * CIN * COUT; i += kTQ) {
        if (p.mirror) {                                         // source [27][CIN(rows)][COUT] already in that order, taps mirrored
            const int tap = i / (CIN * COUT), rest = i % (CIN * COUT);
            wsm[i] = wg[(26 - tap) * CIN * COUT + rest];
        } else wsm[i] = wg[i];
    }
    const int d0 = dt * kTD, q0 = qt * kTQ;
    const int hb = q0 / p.W;
    load_tile<T, CIN>(p, x, tile, n, d0, hb);
    __syncthreads();

    const int q = q0 + threadIdx.x;
    const bool valid = q < p.H * p.W;
    const int h = valid ? q / p.W : hb, wq = valid ? q - h * p.W : 0;
    const int WP = p.W + 2;
    float acc[kTD][COUT];
#pragma unroll
    for (int o = 0; o < kTD; ++o)
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[o][j] = bias ? bias[g * COUT + j] : 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const float* tp = tile + ((size_t)(h - hb + kh) * WP + wq + kw) * CIN;
            float xin[kTD + 2][CIN];
#pragma unroll
            for (int dz = 0; dz < kTD + 2; ++dz) lds_vecN<CIN>(tp + (size_t)dz * p.rows * WP * CIN, xin[dz]);
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                float wv[3][COUT];
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) lds_vecN<COUT>(wsm + (((kd * 3 + kh) * 3 + kw) * CIN + ci) * COUT, wv[kd]);
#pragma unroll
                for (int o = 0; o < kTD; ++o)
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd)
#pragma unroll
                        for (int j = 0; j < COUT; ++j) acc[o][j] = fmaf(xin[o + kd][ci], wv[kd][j], acc[o][j]);
            }
        }
    float s1[COUT], s2[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    if (valid) {
#pragma unroll
        for (int o = 0; o < kTD; ++o) {
            const int d = d0 + o;
            if (d < p.D) {
                VecIO<T, COUT>::store(y + ((((size_t)n * p.D + d) * p.H + h) * p.W + wq) * COUT, acc[o]);
#pragma unroll
                for (int j = 0; j < COUT; ++j) { s1[j] += acc[o][j]; s2[j] += acc[o][j] * acc[o][j]; }
            }
        }
    }
    if (stats) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
            const float a = warp_sum(s1[j]), b = warp_sum(s2[j]);
            if (lane == 0) { red[wid][2 * j] = a; red[wid][2 * j + 1] = b; }
        }
        __syncthreads();
        if (threadIdx.x < 2 * COUT) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += (double)red[k][threadIdx.x];
            atomicAdd(&stats[((size_t)n * COUT + (threadIdx.x >> 1)) * 2 + (threadIdx.x & 1)], t);
        }
    }
}

// dw[g][tap][ci][co] += sum_v x[v + tap - 1][ci] * dy[v][co].  grid = (ctas, 3 / KDS, groups); a CTA owns the taps with
// kd in [kd0, kd0 + KDS) and strides over the tiles of its group.
template <typename T, int CIN, int COUT, int KDS>
__global__ void __launch_bounds__(kTQ, (KDS * 9 * CIN * COUT <= 80) ? 2 : 1)
conv3_small_wgrad_kernel(SmK p, const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;                                         // [kTD+2][rows][W+2][CIN]
    constexpr int NACC = KDS * 9 * CIN * COUT;
    const int g = blockIdx.z, kd0 = blockIdx.y * KDS;
    const int WP = p.W + 2;
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    const int per_sample = p.tiles_q * p.tiles_d;
    const int items = p.npg * per_sample;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int n = g * p.npg + it / per_sample;
        const int r = it % per_sample;
        const int qt = r % p.tiles_q, dt = r / p.tiles_q;
        const int d0 = dt * kTD, q0 = qt * kTQ;
        const int hb = q0 / p.W;
        const int q = q0 + threadIdx.x;
        const bool valid = q < p.H * p.W;
        const int h = valid ? q / p.W : hb, wq = valid ? q - h * p.W : 0;
        const T* dyp = dy + ((((size_t)n * p.D + d0) * p.H + h) * p.W + wq) * COUT;
        const size_t plane = (size_t)p.H * p.W * COUT;
        // Few output channels: ALL kTD dy vectors of this thread are requested before the tile is staged, so their latency hides behind
        // the staging (with one plane of look-ahead the loop was a chain of kTD global-load round trips: ncu showed 60 % of the samples
        // on the long scoreboard at 0.7 instructions per clock and SM).  Wide outputs keep the one-plane look-ahead (register budget).
        constexpr bool PRE = COUT * kTD <= 32;
        float gall[PRE ? kTD : 1][COUT];
        if constexpr (PRE) {
#pragma unroll
            for (int o = 0; o < kTD; ++o) {
                if (valid && d0 + o < p.D) VecIO<T, COUT>::load(dyp + (size_t)o * plane, gall[o]);
                else {
#pragma unroll
                    for (int j = 0; j < COUT; ++j) gall[o][j] = 0.f;
                }
            }
        }
        __syncthreads();                                        // previous tile fully consumed
        load_tile<T, CIN>(p, x, tile, n, d0, hb);
        __syncthreads();
        float gc[COUT], gn[COUT];
        if constexpr (!PRE) {
            if (valid && d0 < p.D) VecIO<T, COUT>::load(dyp, gc);
            else {
#pragma unroll
                for (int j = 0; j < COUT; ++j) gc[j] = 0.f;
            }
        }
#pragma unroll
        for (int o = 0; o < kTD; ++o) {
            if constexpr (PRE) {
#pragma unroll
                for (int j = 0; j < COUT; ++j) gc[j] = gall[o][j];
            } else if (o + 1 < kTD) {
                if (valid && d0 + o + 1 < p.D) VecIO<T, COUT>::load(dyp + (size_t)(o + 1) * plane, gn);
                else {
#pragma unroll
                    for (int j = 0; j < COUT; ++j) gn[j] = 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < KDS; ++k)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float xv[CIN];
                        lds_vecN<CIN>(tile + (((size_t)(o + kd0 + k) * p.rows + (h - hb + kh)) * WP + wq + kw) * CIN, xv);
#pragma unroll
                        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                            for (int j = 0; j < COUT; ++j)
                                acc[((k * 3 + kh) * 3 + kw) * CIN * COUT + ci * COUT + j] =
                                    fmaf(xv[ci], gc[j], acc[((k * 3 + kh) * 3 + kw) * CIN * COUT + ci * COUT + j]);
                    }
            if constexpr (!PRE) {
                if (o + 1 < kTD) {
#pragma unroll
                    for (int j = 0; j < COUT; ++j) gc[j] = gn[j];
                }
            }
        }
    }
    // block reduction: shuffles inside the warps, the 8 warp totals through shared memory (reusing the tile buffer)
    __syncthreads();
    float* red = smem;                                          // [8][NACC]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        const float s = warp_sum(acc[i]);
        if (lane == 0) red[wid * NACC + i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NACC; i += kTQ) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k * NACC + i];
        const int k = i / (9 * CIN * COUT), rest = i % (9 * CIN * COUT);
        atomicAdd(dw + ((size_t)g * 27 + (kd0 + k) * 9) * CIN * COUT + rest, s);
    }
}

// ---- bf16 weight gradient with the NEXT tile's halo in flight while the current one is consumed (round 2) ----
// ncu on the kernel above (c2->2, 80^3, n = 10: 288 us for 41 MB): one CTA of 8 warps per SM (the 108 partial sums take 255 registers),
// 60 % of the stall samples on the long scoreboard — the CTA stages a tile, waits, computes, and nothing else is resident to hide the wait.
// Here the halo is copied RAW (bf16, no conversion pass, no registers) by cp.async into the other half of a double buffer while the
// threads accumulate from the current half; out-of-volume elements are the copy's zero fill, reflection is resolved in the source
// address.  Needs Cin * 2 bytes >= 4 per copy, i.e. Cin >= 2.
template <int BYTES> __device__ __forceinline__ void cp_async_small(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(src), "n"(BYTES), "r"(src_bytes) : "memory");
}

template <int CIN>
__device__ __forceinline__ void stage_tile_async(const SmK& p, const bf16* __restrict__ x, bf16* __restrict__ tile, int n, int d0, int hb) {
    const int WP = p.W + 2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nrows = (kTD + 2) * p.rows;
    for (int row = wid; row < nrows; row += kTQ / 32) {
        const int dz = row / p.rows, r = row - dz * p.rows;
        int id = d0 + dz - p.pad, ih = hb + r - p.pad;
        if (p.reflect) { id = reflect_idx(id, p.Di); ih = reflect_idx(ih, p.Hi); }
        const bool ok_row = id >= 0 && id < p.Di && ih >= 0 && ih < p.Hi;
        const bf16* src = x + (((size_t)n * p.Di + (ok_row ? id : 0)) * p.Hi + (ok_row ? ih : 0)) * p.Wi * CIN;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + (size_t)row * WP * CIN);
        for (int c = lane; c < WP; c += 32) {
            int iw = c - p.pad;
            bool ok = ok_row;
            if (p.reflect) iw = reflect_idx(iw, p.Wi); else ok = ok && iw >= 0 && iw < p.Wi;
            cp_async_small<CIN * 2>(dst + (uint32_t)c * CIN * 2, ok ? src + (size_t)iw * CIN : x, ok ? CIN * 2u : 0u);
        }
    }
}

template <int CIN> __device__ __forceinline__ void lds_bf16N(const bf16* p, float* o) {
    if constexpr (CIN == 2) {
        bf2_unpack(*reinterpret_cast<const uint32_t*>(p), o[0], o[1]);
    } else if constexpr (CIN == 4) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        bf2_unpack(u.x, o[0], o[1]); bf2_unpack(u.y, o[2], o[3]);
    } else {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        bf2_unpack(u.x, o[0], o[1]); bf2_unpack(u.y, o[2], o[3]); bf2_unpack(u.z, o[4], o[5]); bf2_unpack(u.w, o[6], o[7]);
    }
}

template <int CIN, int COUT, int KDS>
__global__ void __launch_bounds__(kTQ, (KDS * 9 * CIN * COUT <= 80) ? 2 : 1)
conv3_small_wgrad_async_kernel(SmK p, const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NACC = KDS * 9 * CIN * COUT;
    const int g = blockIdx.z, kd0 = blockIdx.y * KDS;
    const int WP = p.W + 2;
    const size_t tile_elems = (((size_t)(kTD + 2) * p.rows * WP * CIN + 7) / 8) * 8;           // 16-byte aligned halves
    bf16* tiles = reinterpret_cast<bf16*>(smem);
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    const int per_sample = p.tiles_q * p.tiles_d;
    const int items = p.npg * per_sample;
    int buf = 0;
    if ((int)blockIdx.x < items) {
        const int it = blockIdx.x, r = it % per_sample;
        stage_tile_async<CIN>(p, x, tiles, g * p.npg + it / per_sample, (r / p.tiles_q) * kTD, ((r % p.tiles_q) * kTQ) / p.W);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int n = g * p.npg + it / per_sample;
        const int r = it % per_sample;
        const int qt = r % p.tiles_q, dt = r / p.tiles_q;
        const int d0 = dt * kTD, q0 = qt * kTQ;
        const int hb = q0 / p.W;
        const int q = q0 + threadIdx.x;
        const bool valid = q < p.H * p.W;
        const int h = valid ? q / p.W : hb, wq = valid ? q - h * p.W : 0;
        const bf16* dyp = dy + ((((size_t)n * p.D + d0) * p.H + h) * p.W + wq) * COUT;
        const size_t plane = (size_t)p.H * p.W * COUT;
        float gall[kTD][COUT];
#pragma unroll
        for (int o = 0; o < kTD; ++o) {
            if (valid && d0 + o < p.D) VecIO<bf16, COUT>::load(dyp + (size_t)o * plane, gall[o]);
            else {
#pragma unroll
                for (int j = 0; j < COUT; ++j) gall[o][j] = 0.f;
            }
        }
        {   // next tile of this CTA into the other half (its previous readers passed the barrier at the end of the last iteration)
            const int nx = it + gridDim.x;
            if (nx < items) {
                const int rn = nx % per_sample;
                stage_tile_async<CIN>(p, x, tiles + (size_t)(buf ^ 1) * tile_elems, g * p.npg + nx / per_sample, (rn / p.tiles_q) * kTD,
                                      ((rn % p.tiles_q) * kTQ) / p.W);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // everything but the group just committed: the current tile has landed
        __syncthreads();
        const bf16* tile = tiles + (size_t)buf * tile_elems;
#pragma unroll
        for (int o = 0; o < kTD; ++o) {
#pragma unroll
            for (int k = 0; k < KDS; ++k)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float xv[CIN];
                        lds_bf16N<CIN>(tile + (((size_t)(o + kd0 + k) * p.rows + (h - hb + kh)) * WP + wq + kw) * CIN, xv);
#pragma unroll
                        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                            for (int j = 0; j < COUT; ++j)
                                acc[((k * 3 + kh) * 3 + kw) * CIN * COUT + ci * COUT + j] =
                                    fmaf(xv[ci], gall[o][j], acc[((k * 3 + kh) * 3 + kw) * CIN * COUT + ci * COUT + j]);
                    }
        }
        __syncthreads();                                        // this half may be overwritten by the copies of the next iteration
        buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float* red = smem;                                          // [8][NACC]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        const float sres = warp_sum(acc[i]);
        if (lane == 0) red[wid * NACC + i] = sres;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NACC; i += kTQ) {
        float sres = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sres += red[k * NACC + i];
        const int k = i / (9 * CIN * COUT), rest = i % (9 * CIN * COUT);
        atomicAdd(dw + ((size_t)g * 27 + (kd0 + k) * 9) * CIN * COUT + rest, sres);
    }
}

// dx[v] = sum of the extended-domain values that the reflect padding maps onto v (every voxel; C channels, C <= 8)
template <typename T, int C>
__global__ void __launch_bounds__(256) small_fold_kernel(const T* __restrict__ ext, T* __restrict__ dx, int N, int D, int H, int W) {
    const long long total = (long long)N * D * H * W;
    for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        long long v = t;
        const int w = (int)(v % W); v /= W;
        const int h = (int)(v % H); v /= H;
        const int d = (int)(v % D);
        const int n = (int)(v / D);
        int sd[3], sh[3], sw[3];
        int nd = 0, nh = 0, nw = 0;
        sd[nd++] = d + 1; if (d == 1) sd[nd++] = 0; if (d == D - 2) sd[nd++] = D + 1;
        sh[nh++] = h + 1; if (h == 1) sh[nh++] = 0; if (h == H - 2) sh[nh++] = H + 1;
        sw[nw++] = w + 1; if (w == 1) sw[nw++] = 0; if (w == W - 2) sw[nw++] = W + 1;
        float acc[C];
#pragma unroll
        for (int i = 0; i < C; ++i) acc[i] = 0.f;
        for (int a = 0; a < nd; ++a)
            for (int b = 0; b < nh; ++b)
                for (int c = 0; c < nw; ++c) {
                    float xv[C];
                    VecIO<T, C>::load(ext + ((((size_t)n * (D + 2) + sd[a]) * (H + 2) + sh[b]) * (W + 2) + sw[c]) * C, xv);
#pragma unroll
                    for (int i = 0; i < C; ++i) acc[i] += xv[i];
                }
        VecIO<T, C>::store(dx + (size_t)t * C, acc);
    }
}

int fill(const pb_conv_desc* d, SmK& p, bool full) {
    if (!d || d->ksize != 3 || d->stride != 1 || d->c1 != 0 || d->groups < 1 || d->n % d->groups) return -1;
    p.N = d->n; p.Di = d->di; p.Hi = d->hi; p.Wi = d->wi;
    p.pad = full ? 2 : 1;
    p.D = d->di + (full ? 2 : 0); p.H = d->hi + (full ? 2 : 0); p.W = d->wi + (full ? 2 : 0);
    p.reflect = (!full && d->pad_mode == PB_PAD_REFLECT) ? 1 : 0;
    if (p.reflect && (p.Di < 2 || p.Hi < 2 || p.Wi < 2)) return -1;
    p.npg = d->n / d->groups;
    p.mirror = 0;
    p.tiles_q = (p.H * p.W + kTQ - 1) / kTQ;
    p.tiles_d = (p.D + kTD - 1) / kTD;
    p.rows = (kTQ + p.W - 2) / p.W + 1 + 2;                     // rows touched by 256 consecutive positions, + 2 halo rows
    return 0;
}

size_t tile_bytes(const SmK& p, int cin) { return (size_t)(kTD + 2) * p.rows * (p.W + 2) * cin * sizeof(float); }

template <typename K> int set_smem_small(K kern, size_t bytes) {
    if (bytes > 200 * 1024) { pb_set_error("conv3d_small: tile of %zu B does not fit shared memory", bytes); return PB_EUNSUPPORTED; }
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { pb_set_error("conv3d_small: cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e)); return PB_ECUDA; }
    }
    return 0;
}

template <typename T, int CIN, int COUT>
int launch_fwd(const SmK& p, const void* x, const float* w, const float* bias, void* y, double* stats, cudaStream_t st) {
    const size_t smem = 27 * CIN * COUT * sizeof(float) + tile_bytes(p, CIN);
    auto kern = conv3_small_fwd_kernel<T, CIN, COUT>;
    if (int e = set_smem_small(kern, smem)) return e;
    kern<<<dim3(p.tiles_q * p.tiles_d, p.N), kTQ, smem, st>>>(p, (const T*)x, w, bias, (T*)y, stats);
    return 0;
}

template <typename T, int CIN, int COUT, int KDS>
int launch_wgrad(const SmK& p, int groups, const void* x, const void* dy, float* dw, cudaStream_t st) {
    constexpr int NACC = KDS * 9 * CIN * COUT;
    size_t smem = tile_bytes(p, CIN);
    if (smem < (size_t)8 * NACC * sizeof(float)) smem = (size_t)8 * NACC * sizeof(float);
    auto kern = conv3_small_wgrad_kernel<T, CIN, COUT, KDS>;
    if (int e = set_smem_small(kern, smem)) return e;
    const int items = p.npg * p.tiles_q * p.tiles_d;
    int ctas = 148 * 2 / (groups * (3 / KDS));
    if (ctas < 1) ctas = 1;
    if (ctas > items) ctas = items;
    kern<<<dim3(ctas, 3 / KDS, groups), kTQ, smem, st>>>(p, (const T*)x, (const T*)dy, dw);
    return 0;
}

// supported (Cin, Cout) classes of the forward kernel (the data gradient uses them with the roles swapped)
bool fwd_class(int cin, int cout) {
    return (cin == 1 && cout == 8) || (cin == 2 && cout == 2) || (cin == 4 && cout == 4) || (cin == 8 && cout == 1);
}

template <typename T>
int dispatch_fwd(const SmK& p, int cin, int cout, const void* x, const float* w, const float* bias, void* y, double* stats,
                 cudaStream_t st) {
    if (cin == 1 && cout == 8) return launch_fwd<T, 1, 8>(p, x, w, bias, y, stats, st);
    if (cin == 2 && cout == 2) return launch_fwd<T, 2, 2>(p, x, w, bias, y, stats, st);
    if (cin == 4 && cout == 4) return launch_fwd<T, 4, 4>(p, x, w, bias, y, stats, st);
    if (cin == 8 && cout == 1) return launch_fwd<T, 8, 1>(p, x, w, bias, y, stats, st);
    pb_set_error("conv3d_small: class c%d->%d not built", cin, cout);
    return PB_EUNSUPPORTED;
}

template <int CIN, int COUT, int KDS>
int launch_wgrad_async(const SmK& p, int groups, const void* x, const void* dy, float* dw, cudaStream_t st) {
    constexpr int NACC = KDS * 9 * CIN * COUT;
    const size_t tile_elems = (((size_t)(kTD + 2) * p.rows * (p.W + 2) * CIN + 7) / 8) * 8;
    size_t smem = 2 * tile_elems * sizeof(bf16);
    if (smem < (size_t)8 * NACC * sizeof(float)) smem = (size_t)8 * NACC * sizeof(float);
    auto kern = conv3_small_wgrad_async_kernel<CIN, COUT, KDS>;
    if (int e = set_smem_small(kern, smem)) return e;
    const int items = p.npg * p.tiles_q * p.tiles_d;
    const int per_sm = (NACC <= 80) ? 2 : 1;                    // resident CTAs (register budget of the partial sums)
    int ctas = 148 * per_sm / (groups * (3 / KDS));
    if (ctas < 1) ctas = 1;
    if (ctas > items) ctas = items;
    kern<<<dim3(ctas, 3 / KDS, groups), kTQ, smem, st>>>(p, (const bf16*)x, (const bf16*)dy, dw);
    return 0;
}

int small_wgrad_async_mode() {          // PB_SMALL_WGRAD_ASYNC=0: the single-buffer kernel everywhere (A/B measurements)
    static const int mode = [] { const char* e = getenv("PB_SMALL_WGRAD_ASYNC"); return e != nullptr && e[0] == '0' ? 0 : 1; }();
    return mode;
}

template <typename T>
int dispatch_wgrad(const SmK& p, int groups, int cin, int cout, const void* x, const void* dy, float* dw, cudaStream_t st) {
    if constexpr (sizeof(T) == 2) {
        if (small_wgrad_async_mode()) {
            if (cin == 2 && cout == 2) return launch_wgrad_async<2, 2, 3>(p, groups, x, dy, dw, st);
            if (cin == 4 && cout == 4) return launch_wgrad_async<4, 4, 1>(p, groups, x, dy, dw, st);
        }
    }
    if (cin == 1 && cout == 8) return launch_wgrad<T, 1, 8, 1>(p, groups, x, dy, dw, st);
    if (cin == 2 && cout == 2) return launch_wgrad<T, 2, 2, 3>(p, groups, x, dy, dw, st);
    if (cin == 4 && cout == 4) return launch_wgrad<T, 4, 4, 1>(p, groups, x, dy, dw, st);
    pb_set_error("conv3d_small: class c%d->%d not built", cin, cout);
    return PB_EUNSUPPORTED;
}

}  // namespace

extern "C" int pb_conv3d_small_supported(int cin, int cout) { return (fwd_class(cin, cout) && cout != 1) ? 1 : 0; }

extern "C" int pb_conv3d_small_fwd(const pb_conv_desc* d, const void* x, const float* w, const float* bias, void* y, double* stats,
                                   pb_stream_t stream) {
    SmK p;
    PB_CHECK_ARG(fill(d, p, false) == 0, "bad descriptor (3x3x3, stride 1, single source)");
    PB_CHECK_ARG(x && w && y, "null pointer");
    PB_CHECK_ARG(d->dout == d->di && d->ho == d->hi && d->wo == d->wi, "same-size output only");
    cudaStream_t st = (cudaStream_t)stream;
    int e = d->dtype == PB_BF16 ? dispatch_fwd<bf16>(p, d->c0, d->cout, x, w, bias, y, stats, st)
                                : dispatch_fwd<float>(p, d->c0, d->cout, x, w, bias, y, stats, st);
    if (e) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

// wt: fp32 [G][27][cout][cin] (pb_weight_prep's `wt`).  Zero padding: dx is written directly.  Reflect padding: the full
// correlation goes to `ext` [n][di+2][hi+2][wi+2][cin] (caller-provided scratch) and is folded into dx.
extern "C" int pb_conv3d_small_dgrad(const pb_conv_desc* d, const void* dy, const float* wt, void* dx, void* ext, pb_stream_t stream) {
    PB_CHECK_ARG(d && dy && wt && dx, "null pointer");
    const bool reflect = d->pad_mode == PB_PAD_REFLECT;
    PB_CHECK_ARG(!reflect || ext, "reflect padding needs the extended scratch buffer");
    PB_CHECK_ARG(!reflect || (d->di >= 4 && d->hi >= 4 && d->wi >= 4), "sizes >= 4");
    SmK p;
    PB_CHECK_ARG(fill(d, p, reflect) == 0, "bad descriptor (3x3x3, stride 1, single source)");
    p.mirror = 1;
    const int cin = d->cout, cout = d->c0;                      // roles swapped
    PB_CHECK_ARG(fwd_class(cin, cout), "channel class not supported");
    cudaStream_t st = (cudaStream_t)stream;
    void* out = reflect ? ext : dx;
    int e = d->dtype == PB_BF16 ? dispatch_fwd<bf16>(p, cin, cout, dy, wt, nullptr, out, nullptr, st)
                                : dispatch_fwd<float>(p, cin, cout, dy, wt, nullptr, out, nullptr, st);
    if (e) return e;
    PB_CHECK_LAUNCH();
    if (reflect) {
        const long long total = (long long)d->n * d->di * d->hi * d->wi;
        long long blocks = (total + 255) / 256;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
#define FOLD(T, C) small_fold_kernel<T, C><<<(unsigned)blocks, 256, 0, st>>>((const T*)ext, (T*)dx, d->n, d->di, d->hi, d->wi)
        if (d->dtype == PB_BF16) {
            switch (cout) { case 1: FOLD(bf16, 1); break; case 2: FOLD(bf16, 2); break; case 4: FOLD(bf16, 4); break; default: FOLD(bf16, 8); break; }
        } else {
            switch (cout) { case 1: FOLD(float, 1); break; case 2: FOLD(float, 2); break; case 4: FOLD(float, 4); break; default: FOLD(float, 8); break; }
        }
#undef FOLD
        PB_CHECK_LAUNCH();
    }
    return PB_OK;
}

extern "C" int pb_conv3d_small_wgrad(const pb_conv_desc* d, const void* x, const void* dy, float* dw, pb_stream_t stream) {
    SmK p;
    PB_CHECK_ARG(fill(d, p, false) == 0, "bad descriptor (3x3x3, stride 1, single source)");
    PB_CHECK_ARG(x && dy && dw, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int e = d->dtype == PB_BF16 ? dispatch_wgrad<bf16>(p, d->groups, d->c0, d->cout, x, dy, dw, st)
                                : dispatch_wgrad<float>(p, d->groups, d->c0, d->cout, x, dy, dw, st);
    if (e) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}
