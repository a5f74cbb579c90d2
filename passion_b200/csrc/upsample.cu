// Trilinear up-sampling with align_corners=True and an integer scale factor, NDHWC, fwd + exact adjoint.
// Replaces nn.Upsample(scale_factor=s, mode='trilinear', align_corners=True)
// (reference models/rfnet.py:54,59,64,110-112).  Index arithmetic mirrors ATen's
// area_pixel_compute_source_index for align_corners: src = dst * (in-1)/(out-1) in float.
#include <cstdlib>
#include "common.cuh"

namespace {

__device__ __forceinline__ void src_index(int o, float ratio, int in, int& i0, int& i1, float& l1) {
    const float s = ratio * (float)o;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

// One output vector per thread (fp32 check mode, see run()).  grid = (blocks per output plane, 1, n * output planes): the plane / sample indices and the d interpolation are uniform
// per CTA, a thread resolves only its (oh, ow, channel chunk) — the flat version spent most of its time in five 64-bit
// divisions per 16-byte output vector (1.1 TB/s).
template <typename T, int VEC>
__global__ void __launch_bounds__(256) up_fwd_point_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int d, int h, int w, int c,
                                                     int scale, float rd, float rh, float rw, long long total_vec) {
    const int od_n = d * scale, oh_n = h * scale, ow_n = w * scale, cv = c / VEC;
    const int nn = blockIdx.z / od_n, od = blockIdx.z - nn * od_n;
    int d0, d1; float ld;
    src_index(od, rd, d, d0, d1, ld);
    const int plane_vec = oh_n * ow_n * cv;
    for (int pi = blockIdx.x * blockDim.x + threadIdx.x; pi < plane_vec; pi += gridDim.x * blockDim.x) {
        const int r = pi / cv, cl = pi - r * cv;
        const int oh = r / ow_n, ow = r - oh * ow_n;
        const long long i = (long long)blockIdx.z * plane_vec + pi;
        int h0, h1, w0, w1; float lh, lw;
        src_index(oh, rh, h, h0, h1, lh); src_index(ow, rw, w, w0, w1, lw);
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const T* xb = x + (size_t)nn * d * h * w * c + cl * VEC;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float wt = (a ? ld : 1.f - ld) * (b ? lh : 1.f - lh) * (e ? lw : 1.f - lw);
                    float v[VEC];
                    VecIO<T, VEC>::load(xb + ((((size_t)(a ? d1 : d0)) * h + (b ? h1 : h0)) * w + (e ? w1 : w0)) * c, v);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[j] = fmaf(wt, v[j], acc[j]);
                }
        VecIO<T, VEC>::store(y + i * VEC, acc);
    }
}

// grid = (row tiles, oh segments, n * output planes); a thread owns one (ow, channel chunk) position and walks SEG output
// rows of its plane.  The d- and w-interpolation of an input row ("column" = sum over the four (d, w) corners) is computed
// once and reused by every output row that reads it: ~2.5 loads and 6 fmas x VEC per output vector instead of 8 and 8 (the
// one-output-per-thread version was bound by instruction issue and the load pipe at 1.2 TB/s).  The threads of a warp hold
// consecutive (ow, chunk) positions, so every store instruction writes one contiguous run of the output row.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) up_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int d, int h, int w, int c,
                                                     int scale, float rd, float rh, float rw, int per, int seg) {
    const int od_n = d * scale, oh_n = h * scale, ow_n = w * scale, cv = c / VEC;
    const int rowvec = ow_n * cv;
    const int nn = blockIdx.z / od_n, od = blockIdx.z - nn * od_n;
    const int ls = threadIdx.x / per;                          // segment slot inside the CTA
    const int pi = blockIdx.x * per + (threadIdx.x - ls * per);
    const int oh_lo = (blockIdx.y * (blockDim.x / per) + ls) * seg;
    if (pi >= rowvec || oh_lo >= oh_n) return;
    const int oh_hi = min(oh_lo + seg, oh_n);
    const int ow = pi / cv, cl = pi - ow * cv;
    int d0, d1, w0, w1; float ld, lw;
    src_index(od, rd, d, d0, d1, ld);
    src_index(ow, rw, w, w0, w1, lw);
    const float k00 = (1.f - ld) * (1.f - lw), k01 = (1.f - ld) * lw, k10 = ld * (1.f - lw), k11 = ld * lw;
    const T* xb = x + (size_t)nn * d * h * w * c + cl * VEC;
    const size_t o00 = ((size_t)d0 * h * w + w0) * c, o01 = ((size_t)d0 * h * w + w1) * c;
    const size_t o10 = ((size_t)d1 * h * w + w0) * c, o11 = ((size_t)d1 * h * w + w1) * c;
    const size_t rowpitch = (size_t)w * c;
    auto column = [&](int hh, float* col) {
        float a[VEC], b[VEC], e[VEC], f[VEC];
        const T* r = xb + (size_t)hh * rowpitch;
        VecIO<T, VEC>::load(r + o00, a); VecIO<T, VEC>::load(r + o01, b);
        VecIO<T, VEC>::load(r + o10, e); VecIO<T, VEC>::load(r + o11, f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) col[j] = fmaf(k11, f[j], fmaf(k10, e[j], fmaf(k01, b[j], k00 * a[j])));
    };
    float c0[VEC], c1[VEC];
    int hc0 = -1, hc1 = -1;                                    // input rows held in c0 / c1
    T* yb = y + (((size_t)blockIdx.z * oh_n) * ow_n + ow) * c + cl * VEC;
    for (int oh = oh_lo; oh < oh_hi; ++oh) {                   // oh, h0, h1 are uniform across the CTA's segment slot
        int h0, h1; float lh;
        src_index(oh, rh, h, h0, h1, lh);
        if (h0 != hc0) {
            if (h0 == hc1) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) c0[j] = c1[j];
            } else {
                column(h0, c0);
            }
            hc0 = h0;
        }
        if (h1 != hc1) {
            if (h1 == hc0) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) c1[j] = c0[j];
            } else {
                column(h1, c1);
            }
            hc1 = h1;
        }
        float o[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = fmaf(lh, c1[j], (1.f - lh) * c0[j]);
        VecIO<T, VEC>::store(yb + (size_t)oh * ow_n * c, o);
    }
}

// weight with which output index o reads input index i along one axis (0 if it does not)
__device__ __forceinline__ float axis_weight(int o, int i, float ratio, int in) {
    int i0, i1; float l1;
    src_index(o, ratio, in, i0, i1, l1);
    float wgt = 0.f;
    if (i0 == i) wgt += 1.f - l1;
    if (i1 == i) wgt += l1;
    return wgt;
}

// gather form of the adjoint: dx[i] = sum_o w(o,i) dy[o]; candidates o in [(i-1)/ratio, (i+1)/ratio]
template <typename T, int VEC>
__global__ void __launch_bounds__(256) up_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int d, int h, int w, int c,
                                                     int scale, float rd, float rh, float rw, long long total_vec) {
    const int od_n = d * scale, oh_n = h * scale, ow_n = w * scale, cv = c / VEC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int cl = (int)(t % cv); t /= cv;
        const int iw = (int)(t % w); t /= w;
        const int ih = (int)(t % h); t /= h;
        const int id = (int)(t % d);
        const int nn = (int)(t / d);
        // candidate ranges (one extra on each side guards float rounding; axis_weight rejects non-contributors)
        int od_lo = 0, od_hi = od_n - 1, oh_lo = 0, oh_hi = oh_n - 1, ow_lo = 0, ow_hi = ow_n - 1;
        if (rd > 0.f) { od_lo = max(0, (int)floorf((id - 1) / rd) - 1); od_hi = min(od_n - 1, (int)ceilf((id + 1) / rd) + 1); }
        if (rh > 0.f) { oh_lo = max(0, (int)floorf((ih - 1) / rh) - 1); oh_hi = min(oh_n - 1, (int)ceilf((ih + 1) / rh) + 1); }
        if (rw > 0.f) { ow_lo = max(0, (int)floorf((iw - 1) / rw) - 1); ow_hi = min(ow_n - 1, (int)ceilf((iw + 1) / rw) + 1); }
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const T* gb = dy + (size_t)nn * od_n * oh_n * ow_n * c + cl * VEC;
        for (int od = od_lo; od <= od_hi; ++od) {
            const float wd = axis_weight(od, id, rd, d);
            if (wd == 0.f) continue;
            for (int oh = oh_lo; oh <= oh_hi; ++oh) {
                const float wh = axis_weight(oh, ih, rh, h);
                if (wh == 0.f) continue;
                for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                    const float ww = axis_weight(ow, iw, rw, w);
                    if (ww == 0.f) continue;
                    float v[VEC];
                    VecIO<T, VEC>::load(gb + (((size_t)od * oh_n + oh) * ow_n + ow) * c, v);
                    const float wt = wd * wh * ww;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[j] = fmaf(wt, v[j], acc[j]);
                }
            }
        }
        VecIO<T, VEC>::store(dx + i * VEC, acc);
    }
}

// One axis of the (separable) adjoint: in [outer][n_big][inner] -> out [outer][n_small][inner],
// out[i] = sum_o w(o, i) in[o].  The 3-D adjoint is the composition over W, H, D; the intermediates shrink by the
// scale factor after every pass, so for x4 / x8 this replaces a (2s)^3-term gather per voxel by three 2s-term ones.
// I = index type: 32-bit whenever the tensors allow it (two divisions per output vector; as 64-bit divisions they cost as much as the loads)
template <typename T, int VEC, typename I>
__global__ void __launch_bounds__(256) up_bwd_axis_kernel(const T* __restrict__ in, T* __restrict__ out, I outer, int n_big,
                                                          int n_small, I inner_vec, float ratio, I total_vec) {
    const float inv_ratio = ratio > 0.f ? 1.f / ratio : 0.f;
    for (I t = (I)blockIdx.x * blockDim.x + threadIdx.x; t < total_vec; t += (I)gridDim.x * blockDim.x) {
        const I r = t / inner_vec;
        const I iv = t - r * inner_vec;
        const I o_idx = r / n_small;
        const int i = (int)(r - o_idx * n_small);
        int lo = 0, hi = n_big - 1;
        if (ratio > 0.f) { lo = max(0, (int)floorf((i - 1) * inv_ratio) - 2); hi = min(n_big - 1, (int)ceilf((i + 1) * inv_ratio) + 2); }
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const T* base = in + ((size_t)o_idx * n_big) * inner_vec * VEC + (size_t)iv * VEC;
        for (int o = lo; o <= hi; ++o) {
            const float wgt = axis_weight(o, i, ratio, n_small);
            if (wgt == 0.f) continue;
            float v[VEC];
            VecIO<T, VEC>::load(base + (size_t)o * inner_vec * VEC, v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = fmaf(wgt, v[j], acc[j]);
        }
        VecIO<T, VEC>::store(out + (size_t)t * VEC, acc);
    }
}

template <typename T, int VEC>
int run_axis(const void* in, void* out, long long outer, int n_big, int n_small, long long inner, cudaStream_t st) {
    const float ratio = n_big > 1 ? (float)(n_small - 1) / (float)(n_big - 1) : 0.f;
    const long long inner_vec = inner / VEC;
    const long long total_vec = outer * n_small * inner_vec;
    long long blocks = (total_vec + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    if (blocks < 1) blocks = 1;
    if (total_vec + 148LL * 32 * 256 < 0x7fffffffLL)
        up_bwd_axis_kernel<T, VEC, int><<<(int)blocks, 256, 0, st>>>((const T*)in, (T*)out, (int)outer, n_big, n_small, (int)inner_vec, ratio,
                                                                    (int)total_vec);
    else
        up_bwd_axis_kernel<T, VEC, long long><<<(int)blocks, 256, 0, st>>>((const T*)in, (T*)out, outer, n_big, n_small, inner_vec, ratio,
                                                                          total_vec);
    return 0;
}

template <typename T, int VEC, bool FWD>
int run(const void* a, void* b, int n, int d, int h, int w, int c, int scale, cudaStream_t st) {
    const float rd = d * scale > 1 ? (float)(d - 1) / (float)(d * scale - 1) : 0.f;
    const float rh = h * scale > 1 ? (float)(h - 1) / (float)(h * scale - 1) : 0.f;
    const float rw = w * scale > 1 ? (float)(w - 1) / (float)(w * scale - 1) : 0.f;
    const long long vox = FWD ? (long long)d * h * w * scale * scale * scale : (long long)d * h * w;
    const long long total_vec = (long long)n * vox * c / VEC;
    long long blocks = (total_vec + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    if (blocks < 1) blocks = 1;
    if (FWD) {
        const long long planes = (long long)n * d * scale;
        static const bool point = [] { const char* e = getenv("PB_UP_POINT"); return e != nullptr && e[0] == '1'; }();
        if (point) {                                                 // one output vector per thread (kept for A/B measurements)
            const long long plane_vec = (long long)h * scale * w * scale * (c / VEC);
            long long bx = (plane_vec + 255) / 256;
            if (bx > 1024) bx = 1024;
            if (planes > 65535 || plane_vec > 0x7fffffffLL) { pb_set_error("upsample_fwd: volume too large"); return PB_EUNSUPPORTED; }
            up_fwd_point_kernel<T, VEC><<<dim3((unsigned)bx, 1, (unsigned)planes), 256, 0, st>>>((const T*)a, (T*)b, n, d, h, w, c, scale,
                                                                                                    rd, rh, rw, total_vec);
            return 0;
        }
        const int rowvec = w * scale * (c / VEC), oh_n = h * scale;
        const int per = rowvec < 256 ? rowvec : 256;                 // row positions per CTA
        const int slots = 256 / per;                                 // output-row segments per CTA
        const int seg = 16;
        const int segs = (oh_n + seg - 1) / seg;
        if (planes > 65535 || (segs + slots - 1) / slots > 65535) { pb_set_error("upsample_fwd: volume too large"); return PB_EUNSUPPORTED; }
        up_fwd_kernel<T, VEC><<<dim3((unsigned)((rowvec + per - 1) / per), (unsigned)((segs + slots - 1) / slots), (unsigned)planes),
                                per * slots, 0, st>>>((const T*)a, (T*)b, d, h, w, c, scale, rd, rh, rw, per, seg);
    }
    else     up_bwd_kernel<T, VEC><<<(int)blocks, 256, 0, st>>>((const T*)a, (T*)b, n, d, h, w, c, scale, rd, rh, rw, total_vec);
    return 0;
}

template <bool FWD>
int dispatch(int dtype, const void* a, void* b, int n, int d, int h, int w, int c, int scale, cudaStream_t st) {
    const int v = pb_vec_width(c);
#define UP_CASE(T)                                                     \
    switch (v) {                                                       \
        case 8: return run<T, 8, FWD>(a, b, n, d, h, w, c, scale, st); \
        case 4: return run<T, 4, FWD>(a, b, n, d, h, w, c, scale, st); \
        case 2: return run<T, 2, FWD>(a, b, n, d, h, w, c, scale, st); \
        default: return run<T, 1, FWD>(a, b, n, d, h, w, c, scale, st); \
    }
    if (dtype == PB_BF16) { UP_CASE(bf16) } else { UP_CASE(float) }
#undef UP_CASE
}

}  // namespace

extern "C" int pb_upsample_fwd(int dtype, const void* x, void* y, int n, int d, int h, int w, int c, int scale, pb_stream_t stream) {
    PB_CHECK_ARG(x && y && n > 0 && d > 0 && h > 0 && w > 0 && c > 0 && scale >= 1, "bad argument");
    if (int e = dispatch<true>(dtype, x, y, n, d, h, w, c, scale, (cudaStream_t)stream)) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_upsample_bwd(int dtype, const void* dy, void* dx, int n, int d, int h, int w, int c, int scale, pb_stream_t stream) {
    PB_CHECK_ARG(dy && dx && n > 0 && d > 0 && h > 0 && w > 0 && c > 0 && scale >= 1, "bad argument");
    dispatch<false>(dtype, dy, dx, n, d, h, w, c, scale, (cudaStream_t)stream);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_upsample_bwd_axis(int dtype, const void* in, void* out, long long outer, int n_big, int n_small, long long inner,
                                    int c, pb_stream_t stream) {
    PB_CHECK_ARG(in && out && outer > 0 && n_big > 0 && n_small > 0 && inner > 0 && c > 0 && inner % c == 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int v = pb_vec_width(c);
#define AX_CASE(T)                                                                      \
    switch (v) {                                                                        \
        case 8: run_axis<T, 8>(in, out, outer, n_big, n_small, inner, st); break;       \
        case 4: run_axis<T, 4>(in, out, outer, n_big, n_small, inner, st); break;       \
        case 2: run_axis<T, 2>(in, out, outer, n_big, n_small, inner, st); break;       \
        default: run_axis<T, 1>(in, out, outer, n_big, n_small, inner, st); break;      \
    }
    if (dtype == PB_BF16) { AX_CASE(bf16) } else { AX_CASE(float) }
#undef AX_CASE
    PB_CHECK_LAUNCH();
    return PB_OK;
}
