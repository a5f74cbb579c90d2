// Weight layout conversions, one launch per conv layer and direction.
//
// The model keeps its parameters in the reference's nn.Conv3d layout [Cout][Cin][kd][kh][kw] fp32 (state_dict
// compatibility, reference models/blocks.py:357).  The kernels want
//   wk   fp32 [G][taps][Cin][Cout]                 FFMA forward / weight-gradient layout
//   wt   fp32 [G][taps][Cout][Cin]                 FFMA data-gradient layout
//   img  bf16 [G][Cout tiles][9][NCH][3 NT][8]     tcgen05 forward weight image: (kh,kw) taps, rows kd = 2,1,0 (zero padded)
//   imgT bf16 [G][Cin tiles][27][NCH'][NT'][8]     the same for the data gradient: taps flipped, channels transposed
// Doing this with tensor ops costs ~10 tiny launches per layer and step (permute / stack / zeros / copy / cast); here
// it is one gather kernel before the layer runs and one scatter kernel after its weight gradient.
#include <vector>
#include "common.cuh"

namespace {

struct PrepK {
    const float* w[4];
    const float* b[4];
    int G, cin, cout, taps;
    float* wk; float* wt; bf16* img; bf16* imgT; float* bias;
    int nt, ntT;
    long long n_wk, n_wt, n_img, n_imgT, n_bias;
    int nch, tiles, nchT, tilesT;
    int wstride;            // cin extent of the PARAMETER a group reads from (= cin unless the group is a cin-slice of a wider weight)
    int kws, kwsT;          // image layout of the kw-stacked kernel: [g][tile][3 kh][chunk][rows: kd = 2,1,0 | kw | nt co][8]
};

// (tap, local output row) of image row `row` of the (kh,kw)-plane / kh-plane `t`
__device__ __forceinline__ void img_row(int kws, int nt, int t, int row, int& tap, int& col) {
    if (kws) { tap = (2 - row / (3 * nt)) * 9 + t * 3 + (row / nt) % 3; col = row % nt; }
    else { tap = (2 - row / nt) * 9 + t; col = row % nt; }
}

__device__ __forceinline__ float ref_w(const PrepK& k, int g, int co, int ci, int tap) {
    return __ldg(k.w[g] + ((size_t)co * k.wstride + ci) * k.taps + tap);
}

__device__ __forceinline__ void prep_element(const PrepK& k, const long long t) {
    {
        long long i = t;
        if (i < k.n_wk) {                                   // [g][tap][ci][co]
            const int co = (int)(i % k.cout); i /= k.cout;
            const int ci = (int)(i % k.cin); i /= k.cin;
            const int tap = (int)(i % k.taps);
            const int g = (int)(i / k.taps);
            k.wk[t] = ref_w(k, g, co, ci, tap);
            return;
        }
        i -= k.n_wk;
        if (i < k.n_wt) {                                   // [g][tap][co][ci]
            const long long o = i;
            const int ci = (int)(i % k.cin); i /= k.cin;
            const int co = (int)(i % k.cout); i /= k.cout;
            const int tap = (int)(i % k.taps);
            const int g = (int)(i / k.taps);
            k.wt[o] = ref_w(k, g, co, ci, tap);
            return;
        }
        i -= k.n_wt;
        if (i < k.n_img) {                                  // [g][tile][9 (kh,kw)][chunk][3 nt rows: kd = 2,1,0][8]
            const long long o = i;
            if (k.taps == 1) {                              // 1x1x1 (csrc/conv1_tc.cu): [g][tile][chunk][nt rows][8]
                const int e = (int)(i % 8); i /= 8;
                const int row = (int)(i % k.nt); i /= k.nt;
                const int chunk = (int)(i % k.nch); i /= k.nch;
                const int tile = (int)(i % k.tiles);
                const int g = (int)(i / k.tiles);
                const int ci = chunk * 8 + e, co = tile * k.nt + row;
                k.img[o] = __float2bfloat16_rn((ci < k.cin && co < k.cout) ? ref_w(k, g, co, ci, 0) : 0.f);
                return;
            }
            const int rows = k.kws ? 9 * k.nt : 3 * k.nt, planes = k.kws ? 3 : 9;
            const int e = (int)(i % 8); i /= 8;
            const int row = (int)(i % rows); i /= rows;
            const int chunk = (int)(i % k.nch); i /= k.nch;
            const int t9 = (int)(i % planes); i /= planes;
            const int tile = (int)(i % k.tiles);
            const int g = (int)(i / k.tiles);
            int tap, col;
            img_row(k.kws, k.nt, t9, row, tap, col);
            const int ci = chunk * 8 + e, co = tile * k.nt + col;
            k.img[o] = __float2bfloat16_rn((ci < k.cin && co < k.cout) ? ref_w(k, g, co, ci, tap) : 0.f);
            return;
        }
        i -= k.n_img;
        if (i < k.n_imgT) {                                 // roles of Cin / Cout swapped, taps mirrored
            const long long o = i;
            if (k.taps == 1) {
                const int e = (int)(i % 8); i /= 8;
                const int row = (int)(i % k.ntT); i /= k.ntT;
                const int chunk = (int)(i % k.nchT); i /= k.nchT;
                const int tile = (int)(i % k.tilesT);
                const int g = (int)(i / k.tilesT);
                const int co = chunk * 8 + e, ci = tile * k.ntT + row;
                k.imgT[o] = __float2bfloat16_rn((ci < k.cin && co < k.cout) ? ref_w(k, g, co, ci, 0) : 0.f);
                return;
            }
            const int rows = k.kwsT ? 9 * k.ntT : 3 * k.ntT, planes = k.kwsT ? 3 : 9;
            const int e = (int)(i % 8); i /= 8;
            const int row = (int)(i % rows); i /= rows;
            const int chunk = (int)(i % k.nchT); i /= k.nchT;
            const int t9 = (int)(i % planes); i /= planes;
            const int tile = (int)(i % k.tilesT);
            const int g = (int)(i / k.tilesT);
            int tap, col;
            img_row(k.kwsT, k.ntT, t9, row, tap, col);
            const int co = chunk * 8 + e, ci = tile * k.ntT + col;
            k.imgT[o] = __float2bfloat16_rn((ci < k.cin && co < k.cout) ? ref_w(k, g, co, ci, 26 - tap) : 0.f);
            return;
        }
        i -= k.n_imgT;
        {
            const int co = (int)(i % k.cout);
            const int g = (int)(i / k.cout);
            k.bias[i] = __ldg(k.b[g] + co);
        }
    }
}

__global__ void weight_prep_kernel(PrepK k) {
    const long long total = k.n_wk + k.n_wt + k.n_img + k.n_imgT + k.n_bias;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        prep_element(k, t);
}

// layer of flat element t: prefix[l] <= t < prefix[l+1]
__device__ __forceinline__ int find_layer(const long long* __restrict__ prefix, int n, long long t) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= t) lo = mid; else hi = mid;
    }
    return lo;
}

// every conv layer of the model in ONE launch (weights change once per step): table = [n+1 prefix sums][n PrepK]
__global__ void weight_prep_batch_kernel(const long long* __restrict__ prefix, const PrepK* __restrict__ tab, int n) {
    const long long total = __ldg(prefix + n);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int l = find_layer(prefix, n, t);
        prep_element(tab[l], t - __ldg(prefix + l));
    }
}

struct UnpackK {
    const float* dw; const float* db; const double* dy_stats; int npg;
    float* gw[4]; float* gb[4];
    int G, cin, cout, taps;
    int wstride;            // see PrepK
    int accumulate;         // 1: add to what gw / gb hold (autograd's accumulation semantics), 0: overwrite
};

__device__ __forceinline__ void unpack_element(const UnpackK& k, const long long t) {
    const long long per = (long long)k.cout * k.cin * k.taps;
    const long long n_w = per * k.G;
    {
        if (t < n_w) {
            const int g = (int)(t / per);
            long long i = t - (long long)g * per;
            const long long o = i;
            const int tap = (int)(i % k.taps); i /= k.taps;
            const int ci = (int)(i % k.cin);
            const int co = (int)(i / k.cin);
            const float v = __ldg(k.dw + (((size_t)g * k.taps + tap) * k.cin + ci) * k.cout + co);
            float* dst = k.gw[g] + ((size_t)co * k.wstride + ci) * k.taps + tap;
            *dst = k.accumulate ? *dst + v : v;
        } else {
            const long long i = t - n_w;
            const int g = (int)(i / k.cout), co = (int)(i % k.cout);
            float v;
            if (k.db) v = __ldg(k.db + i);
            else {
                double acc = 0.0;
                for (int n = 0; n < k.npg; ++n) acc += k.dy_stats[(((size_t)g * k.npg + n) * k.cout + co) * 2];
                v = (float)acc;
            }
            k.gb[g][co] = k.accumulate ? k.gb[g][co] + v : v;
        }
    }
}

__device__ __forceinline__ long long unpack_total(const UnpackK& k) {
    return (long long)k.G * k.cout * ((long long)k.cin * k.taps + ((k.db || k.dy_stats) ? 1 : 0));
}

__global__ void weight_unpack_kernel(UnpackK k) {
    const long long total = unpack_total(k);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
        unpack_element(k, t);
}

__global__ void weight_unpack_batch_kernel(const long long* __restrict__ prefix, const UnpackK* __restrict__ tab, int n) {
    const long long total = __ldg(prefix + n);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int l = find_layer(prefix, n, t);
        unpack_element(tab[l], t - __ldg(prefix + l));
    }
}

}  // namespace

namespace {
int make_prep(const pb_weight_prep_desc* d, PrepK& k, long long& total);
int make_unpack(const pb_weight_unpack_desc* d, UnpackK& k, long long& total);
constexpr size_t kEntryBytes = sizeof(PrepK) > sizeof(UnpackK) ? sizeof(PrepK) : sizeof(UnpackK);

template <typename K, typename D, typename F>
int run_batch(const D* descs, int n, void* table, int upload, cudaStream_t st, F make, void (*kern)(const long long*, const K*, int),
              const char* what) {
    static thread_local std::vector<uint8_t> host;
    const size_t pre = ((size_t)(n + 1) * sizeof(long long) + 15) & ~(size_t)15;
    long long grand = 0;
    if (upload) {
        host.assign(pre + (size_t)n * sizeof(K), 0);
        long long* prefix = reinterpret_cast<long long*>(host.data());
        K* tab = reinterpret_cast<K*>(host.data() + pre);
        for (int i = 0; i < n; ++i) {
            long long total = 0;
            if (int e = make(descs + i, tab[i], total)) return e;
            prefix[i] = grand;
            grand += total;
        }
        prefix[n] = grand;
        cudaError_t e = cudaMemcpyAsync(table, host.data(), host.size(), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { pb_set_error("%s: table upload: %s", what, cudaGetErrorString(e)); return PB_ECUDA; }
    } else {
        for (int i = 0; i < n; ++i) {               // only the element count is needed to size the grid
            K tmp; long long total = 0;
            if (int e = make(descs + i, tmp, total)) return e;
            grand += total;
        }
    }
    if (grand == 0) return PB_OK;
    long long blocks = (grand + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    kern<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(table),
                                           reinterpret_cast<const K*>(reinterpret_cast<const uint8_t*>(table) + pre), n);
    return PB_OK;
}
}  // namespace

extern "C" size_t pb_weight_batch_table_bytes(int n) {
    return (((size_t)(n + 1) * sizeof(long long) + 15) & ~(size_t)15) + (size_t)n * kEntryBytes;
}

extern "C" int pb_weight_prep_batch(const pb_weight_prep_desc* descs, int n, void* table, int upload, pb_stream_t stream) {
    PB_CHECK_ARG(descs && n > 0 && table, "bad argument");
    if (int e = run_batch<PrepK>(descs, n, table, upload, (cudaStream_t)stream, make_prep, weight_prep_batch_kernel, __func__)) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_weight_grad_unpack_batch(const pb_weight_unpack_desc* descs, int n, void* table, int upload, pb_stream_t stream) {
    PB_CHECK_ARG(descs && n > 0 && table, "bad argument");
    if (int e = run_batch<UnpackK>(descs, n, table, upload, (cudaStream_t)stream, make_unpack, weight_unpack_batch_kernel, __func__)) return e;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

namespace {
int make_prep(const pb_weight_prep_desc* d, PrepK& k, long long& total) {
    PB_CHECK_ARG(d && d->groups >= 1 && d->groups <= 4, "1..4 weight groups");
    PB_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "ksize 1 or 3");
    PB_CHECK_ARG(d->cin >= 1 && d->cout >= 1, "bad channels");
    k.G = d->groups; k.cin = d->cin; k.cout = d->cout; k.taps = d->ksize * d->ksize * d->ksize;
    k.wstride = d->w_cin_stride > 0 ? d->w_cin_stride : d->cin;
    PB_CHECK_ARG(k.wstride >= k.cin, "w_cin_stride < cin");
    for (int g = 0; g < 4; ++g) {
        k.w[g] = g < k.G ? d->w[g] : nullptr;
        k.b[g] = g < k.G ? d->b[g] : nullptr;
        PB_CHECK_ARG(g >= k.G || k.w[g], "null weight");
        PB_CHECK_ARG(g >= k.G || !d->bias || k.b[g], "null bias");
    }
    k.wk = d->wk; k.wt = d->wt; k.img = (bf16*)d->img; k.imgT = (bf16*)d->imgT; k.bias = d->bias;
    k.nt = d->nt; k.ntT = d->ntT;
    const long long nw = (long long)k.G * k.taps * k.cin * k.cout;
    k.n_wk = k.wk ? nw : 0;
    k.n_wt = k.wt ? nw : 0;
    k.nch = k.tiles = k.nchT = k.tilesT = 1;
    k.kws = (k.img && k.taps == 27) ? pb_conv3d_tc_kws(k.cin, k.cout) : 0;
    k.kwsT = (k.imgT && k.taps == 27) ? pb_conv3d_tc_kws(k.cout, k.cin) : 0;
    k.n_img = k.n_imgT = 0;
    if (k.img) {
        PB_CHECK_ARG(k.nt > 0 && (k.taps == 27 ? k.cin % 8 == 0 : k.cin % 8 == 0), "weight image: nt > 0, cin % 8 == 0");
        if (k.taps == 27) {
            k.nch = k.cin / 8 < 2 ? 2 : k.cin / 8;
            k.tiles = (k.cout + k.nt - 1) / k.nt;
            k.n_img = (long long)k.G * k.tiles * 27 * k.nch * k.nt * 8;
        } else {                                            // 1x1x1: chunk planes rounded up to even (K = 16 per MMA)
            k.nch = (k.cin / 8 + 1) & ~1;
            k.tiles = (k.cout + k.nt - 1) / k.nt;
            k.n_img = (long long)k.G * k.tiles * k.nch * k.nt * 8;
        }
    }
    if (k.imgT) {
        PB_CHECK_ARG(k.ntT > 0 && k.cout % 8 == 0, "data-gradient weight image: ntT > 0, cout % 8 == 0");
        if (k.taps == 27) {
            k.nchT = k.cout / 8 < 2 ? 2 : k.cout / 8;
            k.tilesT = (k.cin + k.ntT - 1) / k.ntT;
            k.n_imgT = (long long)k.G * k.tilesT * 27 * k.nchT * k.ntT * 8;
        } else {
            k.nchT = (k.cout / 8 + 1) & ~1;
            k.tilesT = (k.cin + k.ntT - 1) / k.ntT;
            k.n_imgT = (long long)k.G * k.tilesT * k.nchT * k.ntT * 8;
        }
    }
    k.n_bias = k.bias ? (long long)k.G * k.cout : 0;
    total = k.n_wk + k.n_wt + k.n_img + k.n_imgT + k.n_bias;
    return PB_OK;
}
}  // namespace

extern "C" int pb_weight_prep(const pb_weight_prep_desc* d, pb_stream_t stream) {
    PrepK k;
    long long total = 0;
    if (int e = make_prep(d, k, total)) return e;
    if (total == 0) return PB_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    weight_prep_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(k);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

namespace {
int make_unpack(const pb_weight_unpack_desc* d, UnpackK& k, long long& total) {
    PB_CHECK_ARG(d && d->dw && d->groups >= 1 && d->groups <= 4, "1..4 weight groups");
    k.dw = d->dw; k.db = d->db; k.dy_stats = d->dy_stats; k.npg = d->npg;
    k.G = d->groups; k.cin = d->cin; k.cout = d->cout; k.taps = d->ksize * d->ksize * d->ksize;
    k.wstride = d->w_cin_stride > 0 ? d->w_cin_stride : d->cin;
    PB_CHECK_ARG(k.wstride >= k.cin, "w_cin_stride < cin");
    for (int g = 0; g < 4; ++g) {
        k.gw[g] = g < k.G ? d->gw[g] : nullptr;
        k.gb[g] = g < k.G ? d->gb[g] : nullptr;
        PB_CHECK_ARG(g >= k.G || k.gw[g], "null gradient pointer");
        PB_CHECK_ARG(g >= k.G || !(d->db || d->dy_stats) || k.gb[g], "null bias-gradient pointer");
    }
    k.accumulate = d->accumulate;
    total = (long long)k.G * k.cout * ((long long)k.cin * k.taps + ((k.db || k.dy_stats) ? 1 : 0));
    return PB_OK;
}
}  // namespace

extern "C" int pb_weight_grad_unpack(const pb_weight_unpack_desc* d, pb_stream_t stream) {
    UnpackK k;
    long long total = 0;
    if (int e = make_unpack(d, k, total)) return e;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    weight_unpack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(k);
    PB_CHECK_LAUNCH();
    return PB_OK;
}
