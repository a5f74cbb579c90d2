// 1x1x1 convolution (forward, and — with transposed weights — its data gradient) as a TMA-fed tcgen05 GEMM (sm_100a).
//
// Replaces nn.Conv3d(k=1) of general_conv3d (reference models/blocks.py:357,401-407,448-455,590-596; rfnet.py:69,107) for bf16
// NDHWC tensors whose channel counts are multiples of 8.  A 1x1x1 conv is a skinny GEMM  Y[voxel, co] = X[voxel, ci] W[ci, co]
// with 2..16 FLOP per byte: HBM-bound at 80^3 / 40^3, launch- and latency-bound at the coarse levels, where the FFMA kernel
// (pw_conv_kernel) ran 30-120 us launches over a few MB.  Here:
//   * tile = 128 consecutive voxels of one sample; the A operand is fetched by TMA: one box [128 voxels][8 ch] per 8-channel
//     chunk lands as a K-major no-swizzle chunk plane (the layout the 3x3x3 kernels stage by hand); a tile that runs past the
//     sample's last voxel is zero-filled by the tensor map and its rows are not stored.  Two-source inputs (torch.cat feeding the
//     conv, blocks.py:462) use one tensor map per source.
//   * the weights of the CTA's (group, Cout tile) arrive once by bulk copy: image [chunk][NT rows][8 ch] (pb_weight_prep).
//   * warp roles: warps 0-3 epilogue (TMEM -> registers -> (+bias) -> bf16 stores, split into two outputs for the data gradient
//     of a two-source conv, + InstanceNorm partial sums), warp 4 = single-thread tcgen05.mma issue + TMEM allocation, warp 5 =
//     single-thread TMA producer.  K is consumed in blocks of up to 16 chunks (128 channels) per ring slot; the accumulators
//     are a ring of 128 / NT TMEM stages, so the epilogue of tile i overlaps the MMAs of the following tiles.
#include <cstdlib>
#include <cstring>
#include "tc_common.cuh"

namespace {

constexpr int kC1Threads = 192;
constexpr int kC1Tile = 128;
constexpr int kC1KB = 16;               // chunks (of 8 channels) per ring slot
constexpr int kC1MaxSlots = 8;
constexpr int kC1TmemCols = 128;

struct C1P {
    int N, C0, C1, CO0, CO1, groups, npg;
    long long V;                        // voxels per sample
    int nchr;                           // (C0 + C1) / 8 real chunks
    int nch;                            // chunk planes of the weight image: nchr rounded up to a multiple of 2 (K = 16 per MMA)
    int kblocks;                        // ring slots per tile = ceil(nch / kC1KB)
    int slots;                          // ring depth
    int sub;                            // 128-voxel sub-tiles per ring slot ("super-tile"): amortises the per-slot barrier round trips
    int tiles_per_sample, nt_tiles;     // super-tiles per sample
};

template <int NT>
__global__ void __launch_bounds__(kC1Threads, 2) conv1_tc_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                                                               C1P p, const bf16* __restrict__ wimg, const float* __restrict__ bias,
                                                               bf16* __restrict__ y0, bf16* __restrict__ y1, double* __restrict__ stats, int* err) {
    constexpr int STAGES = kC1TmemCols / NT;
    extern __shared__ __align__(128) uint8_t smem[];
    const int w_bytes = p.nch * NT * 16;
    uint8_t* w_s = smem;
    const int kb_chunks = p.nch < kC1KB ? p.nch : kC1KB;
    const int sub_bytes = kb_chunks * kC1Tile * 16;      // one 128-voxel sub-tile: kb_chunks chunk planes
    const int slot_bytes = p.sub * sub_bytes;
    uint8_t* a_s = smem + ((w_bytes + 127) & ~127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_s + (size_t)p.slots * slot_bytes);
    uint64_t* full = bars;                               // [slots]   TMA -> MMA
    uint64_t* empty = bars + kC1MaxSlots;                // [slots]   MMA -> TMA
    uint64_t* tfull = empty + kC1MaxSlots;               // [STAGES]  MMA -> epilogue
    uint64_t* tempty = tfull + STAGES;                   // [STAGES]  epilogue -> MMA
    uint64_t* wbar = tempty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y, nt = blockIdx.z;
    // contiguous share of the group's tiles: few sample changes per CTA (one statistics flush per change)
    const long long tiles_g = (long long)p.npg * p.tiles_per_sample;
    const long long per = (tiles_g + gridDim.x - 1) / gridDim.x;
    const long long t_begin = (long long)blockIdx.x * per;
    const long long t_end = t_begin + per < tiles_g ? t_begin + per : tiles_g;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.nchr < p.nch) {                                // odd chunk count: the pad chunk plane of every slot stays zero
        const int pad_plane = (p.nchr % kC1KB);
        for (int i = threadIdx.x; i < p.slots * p.sub * kC1Tile; i += kC1Threads) {
            const int s = i / kC1Tile, e = i % kC1Tile;               // s = slot * sub + sub-tile
            *reinterpret_cast<uint4*>(a_s + (size_t)s * sub_bytes + ((size_t)pad_plane * kC1Tile + e) * 16) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)kC1TmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 5) {
        // =============================== TMA producer ===============================
        if (lane == 0 && t_begin < t_end) {
            const int c0ch = p.C0 >> 3;
            uint32_t k = 0;
            for (long long t = t_begin; t < t_end; ++t) {
                const int n = g * p.npg + (int)(t / p.tiles_per_sample);
                const int v0 = (int)(t % p.tiles_per_sample) * kC1Tile * p.sub;
                for (int kb = 0; kb < p.kblocks; ++kb, ++k) {
                    const int slot = k % p.slots;
                    mbar_wait(&empty[slot], ((k / p.slots) & 1) ^ 1, err, 11);
                    const int ch0 = kb * kC1KB;
                    const int nreal = min(kC1KB, p.nchr - ch0);
                    mbar_expect_tx(&full[slot], (uint32_t)(p.sub * nreal) * kC1Tile * 16u);
                    const uint32_t sbase = smem_u32(a_s + (size_t)slot * slot_bytes);
                    for (int sb = 0; sb < p.sub; ++sb)
                        for (int c = 0; c < nreal; ++c) {
                            const int ch = ch0 + c;
                            tma_load_4d(sbase + (uint32_t)(sb * sub_bytes) + (uint32_t)c * kC1Tile * 16u, ch < c0ch ? &map0 : &map1, 0,
                                        v0 + sb * kC1Tile, ch < c0ch ? ch : ch - c0ch, n, &full[slot]);
                        }
                }
            }
        }
    } else if (warp == 4) {
        // =============================== MMA issuer ===============================
        if (lane == 0 && t_begin < t_end) {
            const bf16* wsrc = wimg + ((size_t)(g * p.nt_tiles + nt) * w_bytes) / 2;
            mbar_expect_tx(wbar, (uint32_t)w_bytes);
            for (int off = 0; off < w_bytes; off += 16384) {
                const int nb = min(16384, w_bytes - off);
                bulk_g2s(smem_u32(w_s + off), reinterpret_cast<const uint8_t*>(wsrc) + off, (uint32_t)nb, wbar);
            }
            mbar_wait(wbar, 0, err, 12);
            const uint32_t a_addr = smem_u32(a_s), w_addr = smem_u32(w_s);
            uint32_t k = 0, j = 0;                                     // ring-slot / accumulator-stage counters
            for (long long t = t_begin; t < t_end; ++t) {
                if (p.sub == 1) {
                    const uint32_t stage = j % STAGES;
                    mbar_wait(&tempty[stage], ((j / STAGES) & 1) ^ 1, err, 13);
                    for (int kb = 0; kb < p.kblocks; ++kb, ++k) {
                        const int slot = k % p.slots;
                        mbar_wait(&full[slot], (k / p.slots) & 1, err, 14);
                        tc_fence_after();
                        const int ch0 = kb * kC1KB;
                        const int steps = (min(kC1KB, p.nch - ch0)) / 2;           // K = 16 channels per instruction
                        for (int ks = 0; ks < steps; ++ks) {
                            const uint64_t ad = umma_desc(a_addr + slot * slot_bytes + (uint32_t)(2 * ks) * kC1Tile * 16u, kC1Tile * 16u, 128);
                            const uint64_t bd = umma_desc(w_addr + (uint32_t)(ch0 + 2 * ks) * NT * 16u, NT * 16u, 128);
                            umma_f16(tmem_base + stage * NT, ad, bd, umma_idesc(kC1Tile, NT), (kb > 0 || ks > 0) ? 1u : 0u);
                        }
                        umma_commit(&empty[slot]);
                    }
                    umma_commit(&tfull[stage]);
                    ++j;
                } else {                                               // one slot = p.sub sub-tiles, one K block (cin <= 128)
                    const int slot = k % p.slots;
                    mbar_wait(&full[slot], (k / p.slots) & 1, err, 14);
                    tc_fence_after();
                    const int steps = p.nch / 2;
                    for (int sb = 0; sb < p.sub; ++sb, ++j) {
                        const uint32_t stage = j % STAGES;
                        mbar_wait(&tempty[stage], ((j / STAGES) & 1) ^ 1, err, 13);
                        tc_fence_after();
                        for (int ks = 0; ks < steps; ++ks) {
                            const uint64_t ad = umma_desc(a_addr + slot * slot_bytes + sb * sub_bytes + (uint32_t)(2 * ks) * kC1Tile * 16u,
                                                          kC1Tile * 16u, 128);
                            const uint64_t bd = umma_desc(w_addr + (uint32_t)(2 * ks) * NT * 16u, NT * 16u, 128);
                            umma_f16(tmem_base + stage * NT, ad, bd, umma_idesc(kC1Tile, NT), ks > 0 ? 1u : 0u);
                        }
                        umma_commit(&tfull[stage]);
                    }
                    umma_commit(&empty[slot]);
                    ++k;
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue (warps 0-3 = TMEM lane quadrants) ===============================
        const int cout = p.CO0 + p.CO1;
        const int cb0 = nt * NT;
        const int creal = min(NT, cout - cb0);               // real channels of this tile: 2, 4 or a multiple of 8
        float s1[NT], s2[NT];
#pragma unroll
        for (int c = 0; c < NT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
        int cur_n = -1;
        auto flush = [&](int n) {
            if (stats == nullptr || n < 0) return;
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                const float a = warp_sum(s1[c]), b = warp_sum(s2[c]);
                if (lane == 0 && c < creal) {
                    atomicAdd(&stats[((size_t)n * cout + cb0 + c) * 2], (double)a);
                    atomicAdd(&stats[((size_t)n * cout + cb0 + c) * 2 + 1], (double)b);
                }
                s1[c] = 0.f; s2[c] = 0.f;
            }
        };
        uint32_t j = 0;
        for (long long t = t_begin; t < t_end; ++t) {
            const int n = g * p.npg + (int)(t / p.tiles_per_sample);
            if (n != cur_n) { flush(cur_n); cur_n = n; }
            for (int sb = 0; sb < p.sub; ++sb, ++j) {
                const long long v = ((long long)(t % p.tiles_per_sample) * p.sub + sb) * kC1Tile + warp * 32 + lane;
                const uint32_t stage = j % STAGES;
                mbar_wait(&tfull[stage], (j / STAGES) & 1, err, 15);
                tc_fence_after();
                float acc[NT];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + stage * NT;
#pragma unroll
                for (int c = 0; c < NT; c += 16) tmem_ld16(taddr + c, acc + c);
                tc_fence_before();
                mbar_arrive(&tempty[stage]);
                if (bias != nullptr) {
#pragma unroll
                    for (int c = 0; c < NT; ++c)
                        if (c < creal) acc[c] += __ldg(bias + (size_t)g * cout + cb0 + c);
                }
                if (v < p.V) {
                    const size_t vox = (size_t)n * p.V + v;
                    if (creal >= 8) {
#pragma unroll
                        for (int c8 = 0; c8 < NT / 8; ++c8) {
                            if (c8 * 8 < creal) {
                                const int cb = cb0 + c8 * 8;
                                bf16* dst = cb < p.CO0 ? y0 + vox * p.CO0 + cb : y1 + vox * p.CO1 + (cb - p.CO0);
                                VecIO<bf16, 8>::store(dst, acc + c8 * 8);
                            }
                        }
                    } else if (creal == 4) {
                        VecIO<bf16, 4>::store(y0 + vox * 4, acc);
                    } else {
                        VecIO<bf16, 2>::store(y0 + vox * 2, acc);
                    }
#pragma unroll
                    for (int c = 0; c < NT; ++c) { s1[c] += acc[c]; s2[c] += acc[c] * acc[c]; }
                }
            }
        }
        flush(cur_n);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kC1TmemCols) : "memory");
    }
}

// dense [N][V][C] bf16 seen as (8 ch of a chunk, V voxels, C/8 chunks, N samples); box = [128 voxels][8 ch] of one chunk
int make_voxel_map(CUtensorMap* map, const void* base, int C, long long V, int N) {
    EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr) return -1;
    const cuuint64_t dims[4] = {8, (cuuint64_t)V, (cuuint64_t)(C / 8), (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, 16, (cuuint64_t)V * C * 2};
    const cuuint32_t box[4] = {8, (cuuint32_t)kC1Tile, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

template <int NT>
int launch_c1(const CUtensorMap& m0, const CUtensorMap& m1, C1P p, const void* wimg, const float* bias, void* y0, void* y1, double* stats,
              int* err, cudaStream_t st) {
    const size_t w_bytes = (size_t)p.nch * NT * 16;
    const int kb_chunks = p.nch < kC1KB ? p.nch : kC1KB;
    // sub-tiles per slot: up to half of the accumulator stages, slot <= 16 KB, only for single-K-block classes
    int sub = 1;
    if (p.kblocks == 1) {
        sub = (kC1TmemCols / NT) / 2;
        while (sub > 1 && (size_t)sub * kb_chunks * kC1Tile * 16 > 16384) sub /= 2;
        while (sub > 1 && (long long)kC1Tile * sub > p.V) sub /= 2;
    }
    p.sub = sub;
    p.tiles_per_sample = (int)((p.V + (long long)kC1Tile * sub - 1) / ((long long)kC1Tile * sub));
    const size_t slot_bytes = (size_t)sub * kb_chunks * kC1Tile * 16;
    const size_t fixed = ((w_bytes + 127) & ~(size_t)127) + (2 * kC1MaxSlots + 2 * (kC1TmemCols / NT) + 1) * 8 + 16;
    // ring depth: as many slots as fit next to a second CTA on the SM (113 KB each), at most 8, at least 2 tiles' worth
    int slots = (int)((113 * 1024 - fixed) / slot_bytes);
    if (slots > kC1MaxSlots) slots = kC1MaxSlots;
    if (slots < 2 * p.kblocks) slots = (int)((226 * 1024 - fixed) / slot_bytes) < kC1MaxSlots ? (int)((226 * 1024 - fixed) / slot_bytes) : kC1MaxSlots;
    if (slots < p.kblocks || slots < 2) { pb_set_error("conv1_tc: operands do not fit shared memory"); return PB_EUNSUPPORTED; }
    p.slots = slots;
    const size_t smem = fixed + (size_t)slots * slot_bytes;
    auto kern = conv1_tc_kernel<NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("conv1_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    const long long tiles_g = (long long)p.npg * p.tiles_per_sample;
    int ctas = 148 / (p.groups * p.nt_tiles);
    if (smem <= 113 * 1024) ctas *= 2;
    if (ctas < 1) ctas = 1;
    if (ctas > tiles_g) ctas = (int)tiles_g;
    kern<<<dim3(ctas, p.groups, p.nt_tiles), kC1Threads, smem, st>>>(m0, m1, p, (const bf16*)wimg, bias, (bf16*)y0, (bf16*)y1, stats, err);
    return 0;
}

}  // namespace

// Cout tile of the 1x1x1 tensor-core kernel for (cin, cout) — 16 or 32 (wider outputs run as several tiles, blockIdx.z: the
// epilogue keeps 3 x NT floats per thread) — or 0 when the class is not covered: cin a multiple of 8 in [8, 512]; cout 2, 4 or a
// multiple of 8 up to 128.  The weight image pb_weight_prep writes for ksize = 1 is
// [groups][cout tiles][chunk planes = cin/8 rounded up to even][NT rows][8 ch] bf16, zero padded.
extern "C" int pb_conv1_tc_ntile(int cin, int cout) {
    static const int mode = [] { const char* e = getenv("PB_C1_TC"); return e != nullptr && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 1; }();
    if (mode == 0 || cin % 8 || cin < 8 || cin > 512) return 0;
    if (cout != 2 && cout != 4 && (cout % 8 || cout < 8 || cout > 128)) return 0;
    // Routed by measurement (profiles/r02_conv1_tc.txt): the GEMM wins where a voxel carries enough work per 128-row MMA tile —
    // wide inputs (cin >= 128: the coarse levels, where the FFMA kernel ran 40-120 us launches over a few MB) and the 16..32 ->
    // 16..32 classes; with fewer channels a tile is 2-4 KB and the per-tile TMEM / barrier round trips dominate, and
    // pw_conv2_kernel (3.5-4 TB/s there) stays ahead.  PB_C1_TC=2 routes every covered class (A/B runs), =0 none.
    if (mode == 1 && !(cin >= 128 || (cin >= 16 && cin <= 32 && cout >= 16))) return 0;
    return cout <= 16 ? 16 : 32;
}

// y0|y1 [n][voxels][co0|co1] = (x0|x1 [n][voxels][c0|c1]) * W (+ bias), bf16; optional InstanceNorm sums of the output.
// The data gradient of a 1x1x1 conv is the same call on dy with the transposed weight image (split output = the two sources).
extern "C" int pb_conv1_tc(const pb_conv_desc* d, const void* x0, const void* x1, const void* wimg, const float* bias, void* y0, void* y1,
                           int co0, int co1, double* stats, int* err_flag, pb_stream_t stream) {
    PB_CHECK_ARG(d && x0 && wimg && y0 && err_flag, "null pointer");
    PB_CHECK_ARG(d->dtype == PB_BF16 && d->ksize == 1 && d->stride == 1, "bf16, 1x1x1, stride 1 only");
    const int cin = d->c0 + d->c1, cout = co0 + co1;
    PB_CHECK_ARG(cout == d->cout && (co1 == 0 || (y1 && co0 % 8 == 0 && co1 % 8 == 0)), "bad output split");
    PB_CHECK_ARG(d->c0 % 8 == 0 && d->c1 % 8 == 0 && d->c0 >= 8 && (d->c1 == 0 || x1), "input channels must be multiples of 8");
    PB_CHECK_ARG(d->groups >= 1 && d->n % d->groups == 0, "bad groups");
    const int NT = pb_conv1_tc_ntile(cin, cout);
    if (NT == 0) { pb_set_error("conv1_tc: no kernel for cin %d cout %d", cin, cout); return PB_EUNSUPPORTED; }
    C1P p;
    p.N = d->n; p.C0 = d->c0; p.C1 = d->c1; p.CO0 = co0; p.CO1 = co1; p.groups = d->groups; p.npg = d->n / d->groups;
    p.V = (long long)d->di * d->hi * d->wi;
    if (p.V < kC1Tile || p.V > 0x7fffffffLL) { pb_set_error("conv1_tc: %lld voxels per sample", p.V); return PB_EUNSUPPORTED; }
    p.nchr = cin / 8; p.nch = (p.nchr + 1) & ~1;
    p.kblocks = (p.nch + kC1KB - 1) / kC1KB;
    p.tiles_per_sample = (int)((p.V + kC1Tile - 1) / kC1Tile);
    p.nt_tiles = (cout + NT - 1) / NT;
    p.slots = 0;
    CUtensorMap m0, m1;
    memset(&m0, 0, sizeof(m0));
    memset(&m1, 0, sizeof(m1));
    if (make_voxel_map(&m0, x0, d->c0, p.V, d->n) != 0 || (d->c1 > 0 && make_voxel_map(&m1, x1, d->c1, p.V, d->n) != 0)) {
        pb_set_error("conv1_tc: cuTensorMapEncodeTiled failed");
        return PB_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (NT == 16) rc = launch_c1<16>(m0, m1, p, wimg, bias, y0, y1, stats, err_flag, st);
    else rc = launch_c1<32>(m0, m1, p, wimg, bias, y0, y1, stats, err_flag, st);
    if (rc) return rc;
    PB_CHECK_LAUNCH();
    return PB_OK;
}
