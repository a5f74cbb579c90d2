// 3x3x3 stride-1 convolution as a tcgen05 implicit GEMM (sm_100a), bf16 operands, fp32 accumulation in TMEM.
//
// Replaces nn.Conv3d(k=3, padding_mode='reflect') of general_conv3d (reference models/blocks.py:357) for the
// channel classes that carry 95 % of the FLOPs, forward and (with flipped/transposed
// weights and zero padding) the data gradient.
//
// GEMM view per output voxel tile:  D[128 voxels x NT] += A[128 x 16] * B[16 x NT]  for each of the 27 taps and each
// 16-channel K step.  Tile = 128 CONSECUTIVE positions q = h*PW + w of the padded-pitch plane (PW = W + 2), so that for
// every tap (kd,kh,kw) the A operand is simply the same 128-row window of the staged input plane d+kd shifted by
// kh*PW + kw rows: no im2col expansion in shared memory, one staged copy serves all 27 taps.
//   * smem A layout: channel-chunk planes [chunk of 8 ch][row] with 16 B per row -> the canonical no-swizzle K-major
//     UMMA layout (8 rows x 16 B core matrices, SBO = 128 B between row groups, LBO = plane pitch between the two
//     8-channel halves of a K=16 step); the UMMA descriptor's start address selects tap and plane.
//   * a CTA sweeps along d: a ring of input planes lives in smem, each plane is fetched once per sweep
//     (cp.async 16 B copies; reflect / zero padding and the two-source channel concat are resolved in the address
//     computation), the weights of the CTA's (group, Cout tile) are fetched once with a TMA bulk copy.
//   * warp roles: warps 0-3 epilogue (TMEM -> registers -> bf16 NDHWC stores + InstanceNorm partial sums),
//     warp 4 single-thread tcgen05.mma issue + TMEM allocation, warps 5-7 producers.  Two TMEM accumulator stages
//     overlap the epilogue of plane d with the MMAs of plane d+1.
#include <cstdlib>
#include <cstring>
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kProducerThreads = 96;
constexpr int kTcThreads = 288;         // conv3_tc_kernel: 4 epilogue warps, 1 MMA-issuer warp, 4 producer warps
constexpr int kTcProducers = 128;
constexpr int kMaxCopies = 16;          // 16-B copies per producer thread and plane ...
constexpr int kMaxCopiesWide = 20;      // ... and for the 64-channel variants (NCHR = 8, which have registers to spare)
constexpr int kMaxCopies128 = 26;       // 128 input channels (NCHR = 16): planes up to W = 37 (mmFormer's 128 -> 64 decoder conv runs at 16^3;
                                        // with 20 the 16^3 launch fell back to the FFMA kernel: 2 x 1.9 ms per step at 128^3)
__host__ __device__ constexpr int max_copies(int nchr) { return nchr >= 16 ? kMaxCopies128 : nchr >= 8 ? kMaxCopiesWide : kMaxCopies; }
constexpr int kWgCopies = 5;            // weight-gradient kernels: slab rows per producer thread (planes up to W = 174)
constexpr int kTileM = 128;

struct TcP {
    int N, D, H, W, C0, C1, CO0, CO1;   // inputs x0|x1 (concat), outputs y0|y1 (split)
    int reflect;
    int inset;                          // 1: the output domain is the input grown by one voxel per side ("full" correlation,
                                        // zero padding 2) — the data-gradient of a reflect-padded conv before folding
    int PW, QT, DCH, ND, npg, groups;
    int slab_need, slab_e;              // rows needed / padded rows per chunk plane of a slab
    int nt_tiles;
    long long w_tile_bytes;             // bytes of one (group, Cout tile) weight image
    int tma, tma_rows;                  // 1: the input planes are fetched by TMA (zero padding, one source): per plane and 8-channel
                                        // chunk one cp.async.bulk.tensor box [tma_rows padded rows][PW][8 ch], out-of-bounds = the zero
                                        // padding; the slab then starts at padded row r0 = q0 / PW and the tile at offset q0 - r0 * PW
    int q_stride;                       // positions a tile advances by (128, or 126 for the kw-stacked kernel)
    int probe;                          // developer probe (PB_TC_PROBE=1): cycle counters of the MMA thread and of epilogue thread 0
};

__device__ unsigned long long tc_dbg[8];


// TMA producer of conv3_tc_kernel / conv3_tc_kws_kernel (one thread): streams the input planes of this CTA's work items into the
// slot ring, one box [tma_rows padded rows][PW][8 ch] per plane and 8-channel chunk (chunks of the second source of a two-source
// conv come from its own tensor map).  Zero padding is the tensor map's out-of-bounds fill; a plane outside the volume (zero
// padding along d) is fetched with an h coordinate beyond the volume, i.e. as an all-zero box.
// (Reflect padding cannot be expressed as an out-of-bounds fill.  Fetching the same boxes and patching the one-voxel in-plane
// halo in shared memory by two extra warps was built and measured in round 2: the extra pipeline stage and the larger
// row-aligned slabs made the forward family 6 % SLOWER than the cp.async producers, so reflect-padded launches keep those.)
template <int NCHR, int kSlots>
__device__ __forceinline__ void tma_producer(const TcP& p, const CUtensorMap* map0, const CUtensorMap* map1, uint8_t* slab_s,
                                             int slot_bytes, uint64_t* full, uint64_t* empty, int g, int items, int* err) {
    const int Di = p.D - 2 * p.inset, Hi = p.H - 2 * p.inset;
    const int c0ch = p.C0 >> 3;
    const uint32_t bytes = (uint32_t)NCHR * 16u * (uint32_t)p.PW * (uint32_t)p.tma_rows;
    uint32_t k = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int dc = it % p.ND, r1 = it / p.ND;
        const int qt = r1 % p.QT, n = g * p.npg + r1 / p.QT;
        const int d0 = dc * p.DCH;
        const int nout = min(p.DCH, p.D - d0);
        const int r0 = (qt * p.q_stride) / p.PW;                   // first padded row of the slab
        for (int pl = 0; pl < nout + 2; ++pl, ++k) {
            const int slot = k % kSlots;
            mbar_wait(&empty[slot], ((k / kSlots) & 1) ^ 1, err, 1);
            const int dp = d0 - 1 + pl - p.inset;
            const bool plane_ok = dp >= 0 && dp < Di;
            mbar_expect_tx(&full[slot], bytes);
            const uint32_t sbase = smem_u32(slab_s + (size_t)slot * slot_bytes);
#pragma unroll
            for (int ch = 0; ch < NCHR; ++ch)
                tma_load_5d(sbase + (uint32_t)(ch * p.slab_e) * 16u, ch < c0ch ? map0 : map1, 0, -1 - p.inset,
                            plane_ok ? r0 - 1 - p.inset : Hi + 8, ch < c0ch ? ch : ch - c0ch, plane_ok ? n * Di + dp : 0, &full[slot]);
        }
    }
}


// The three kd taps are stacked along N: the MMAs of INPUT plane pl (9 x K-steps of them, N = 3 NT) add its contribution
// to the three output planes pl-2, pl-1, pl at once, whose accumulators are adjacent blocks of a ring of TMEM column blocks.
// Why: a tcgen05.mma with N <= 48 holds the tensor pipe for ~40-45 cycles however small it is
// (scripts/microbench/umma_rate.cu), so the pipe time of this kernel is (number of MMA instructions) x 40 cycles; stacking
// cuts the instruction count per output plane from 27 to ~10 without any shifted summation in the epilogue (the shifts are
// between planes, i.e. between accumulator blocks).  All MMAs accumulate; the epilogue zeroes a block after draining it.
// NCHR = real input chunks of 8 channels (1,2,4,8,16); NT = Cout tile (16 or 32); the 128-channel variant (NCHR = 16) has
// room for a 2-plane ring only (110 KB of weights + 2 x 40 KB planes, one CTA per SM) — its layers are the small deep ones
template <int NCHR, int NT>
__global__ void __launch_bounds__(kTcThreads, 2) conv3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap1, TcP p,
                                                               const bf16* __restrict__ x0, const bf16* __restrict__ x1,
                                                               const bf16* __restrict__ wimg, const float* __restrict__ bias,
                                                               bf16* __restrict__ y0, bf16* __restrict__ y1, bf16* __restrict__ yext,
                                                               double* __restrict__ stats, int* err) {
    constexpr int NCH = NCHR < 2 ? 2 : NCHR;             // a K=16 MMA step needs two 8-channel chunks (zero chunk if Cin = 8)
    constexpr int KS = NCH / 2;
    constexpr int kSlots = NCHR >= 16 ? 2 : 6;           // input-plane ring (every plane is consumed exactly once)
    constexpr int TMEM_COLS = 256;                        // two CTAs per SM share the 512 columns
    constexpr int R = TMEM_COLS / NT;                     // accumulator ring: one block of NT columns per output plane in flight
    // input extent (= output extent unless p.inset)
    const int Di = p.D - 2 * p.inset, Hi = p.H - 2 * p.inset, Wi = p.W - 2 * p.inset;
    extern __shared__ __align__(128) uint8_t smem[];
    const int w_bytes = 27 * NCH * NT * 16;              // [9 (kh,kw)][NCH chunks][3 NT rows: kd = 2,1,0][8 channels]
    uint8_t* w_s = smem;
    const int slot_bytes = NCH * p.slab_e * 16;
    uint8_t* slab_s = smem + ((w_bytes + 127) & ~127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(slab_s + (size_t)kSlots * slot_bytes);
    uint64_t* full = bars;                   // [kSlots]  producers -> MMA
    uint64_t* empty = bars + kSlots;         // [kSlots]  MMA -> producers
    uint64_t* blk_full = bars + 2 * kSlots;  // [R]       MMA -> epilogue: output plane complete
    uint64_t* blk_empty = blk_full + R;      // [R]       epilogue -> MMA: block drained and zeroed
    uint64_t* wbar = blk_empty + R;          // weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y, nt = blockIdx.z;
    const int items = p.npg * p.QT * p.ND;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], p.tma ? 1 : kTcProducers); mbar_init(&empty[i], 1); }
        for (int i = 0; i < R; ++i) { mbar_init(&blk_full[i], 1); mbar_init(&blk_empty[i], 128); }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (NCHR < 2) {                          // zero the dummy chunk plane of every slot once
        for (int i = threadIdx.x; i < kSlots * p.slab_e; i += kTcThreads) {
            const int s = i / p.slab_e, e = i % p.slab_e;
            *reinterpret_cast<uint4*>(slab_s + (size_t)s * slot_bytes + ((size_t)p.slab_e + e) * 16) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < 4) {                          // every accumulator block starts from zero
#pragma unroll 1
        for (int c = 0; c < TMEM_COLS; c += 16) tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= 5 && p.tma) {
        // =============================== producer: one thread, TMA ===============================
        if (warp == 5 && lane == 0) tma_producer<NCHR, kSlots>(p, &tmap, &tmap1, slab_s, slot_bytes, full, empty, g, items, err);
    } else if (warp >= 5) {
        // =============================== producers ===============================
        // The (h, w) geometry of a slab row depends only on the q-tile, so each thread resolves its <= kMaxCopies
        // copies (source offset inside a plane, padding, destination) once per work item and then streams planes.
        const int pt = threadIdx.x - 5 * 32;
        const int c0ch = p.C0 >> 3;
        const int copies = p.slab_need * NCHR;
        uint32_t k = 0;                                        // running plane counter (whole kernel)
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int dc = it % p.ND, r1 = it / p.ND;
            const int qt = r1 % p.QT, n = g * p.npg + r1 / p.QT;
            const int d0 = dc * p.DCH, q0 = qt * kTileM;
            const int nout = min(p.DCH, p.D - d0);
            int soff[max_copies(NCHR)];                              // element offset inside the source plane, -1 = zero fill
            uint32_t doff[max_copies(NCHR)];                         // byte offset inside the slot
            uint32_t from1 = 0;
#pragma unroll
            for (int i = 0; i < max_copies(NCHR); ++i) {
                const int idx = pt + i * kTcProducers;
                soff[i] = -1; doff[i] = 0;
                if (idx < copies) {
                    const int ch = idx % NCHR, e = idx / NCHR;
                    const int f = q0 + e;
                    const int hp = f / p.PW, wp = f - hp * p.PW;
                    int h = hp - 1 - p.inset, w = wp - 1 - p.inset;
                    bool ok = hp < p.H + 2;
                    if (p.reflect) { h = reflect_idx(h, Hi); w = reflect_idx(w, Wi); ok = ok && h >= 0 && h < Hi; }
                    else ok = ok && h >= 0 && h < Hi && w >= 0 && w < Wi;
                    doff[i] = (uint32_t)(ch * p.slab_e + e) * 16;
                    if (ok) {
                        if (ch < c0ch) soff[i] = (h * Wi + w) * p.C0 + ch * 8;
                        else { soff[i] = (h * Wi + w) * p.C1 + (ch - c0ch) * 8; from1 |= 1u << i; }
                    }
                }
            }
            for (int pl = 0; pl < nout + 2; ++pl, ++k) {
                const int slot = k % kSlots;
                mbar_wait(&empty[slot], ((k / kSlots) & 1) ^ 1, err, 1);
                int dp = d0 - 1 + pl - p.inset;
                bool plane_ok = true;
                if (p.reflect) dp = reflect_idx(dp, Di); else plane_ok = dp >= 0 && dp < Di;
                if (!plane_ok) dp = 0;
                const size_t plane = ((size_t)n * Di + dp) * Hi * Wi;
                const bf16* p0 = x0 + plane * p.C0;
                const bf16* p1 = x1 + plane * p.C1;
                const uint32_t sbase = smem_u32(slab_s + (size_t)slot * slot_bytes);
#pragma unroll
                for (int i = 0; i < max_copies(NCHR); ++i) {
                    if (pt + i * kTcProducers < copies) {
                        const bool ok = plane_ok && soff[i] >= 0;
                        const bf16* src = ((from1 >> i) & 1u) ? p1 : p0;
                        cp_async16(sbase + doff[i], ok ? src + soff[i] : x0, ok ? 16u : 0u);
                    }
                }
                cp_async_arrive_noinc(&full[slot]);
            }
        }
        cp_async_wait_all();
    } else if (warp == 4) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const bf16* wsrc = wimg + ((size_t)(g * p.nt_tiles + nt) * w_bytes) / 2;
            mbar_expect_tx(wbar, (uint32_t)w_bytes);
            for (int off = 0; off < w_bytes; off += 16384) {
                const int nb = min(16384, w_bytes - off);
                bulk_g2s(smem_u32(w_s + off), reinterpret_cast<const uint8_t*>(wsrc) + off, (uint32_t)nb, wbar);
            }
            mbar_wait(wbar, 0, err, 2);
            const uint32_t slab_addr = smem_u32(slab_s);
            const uint64_t b0 = umma_desc(smem_u32(w_s), 3 * NT * 16, 128);      // LBO = one chunk plane of 3 NT rows
            uint32_t toff[9];                                  // (kh, kw) start-address shifts, in 16 B units
#pragma unroll
            for (int r = 0; r < 9; ++r) toff[r] = (uint32_t)((r / 3) * p.PW + (r % 3));
            uint32_t k = 0, j0 = 0;                            // running input-plane / output-plane counters (whole kernel)
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int dc = it % p.ND;
                const int nout = min(p.DCH, p.D - dc * p.DCH);
                // TMA slabs start at a padded-row boundary: the tile begins (q0 mod PW) rows into the slab
                const uint32_t tile_off = p.tma ? (uint32_t)((((it / p.ND) % p.QT) * p.q_stride) % p.PW) * 16u : 0u;
                for (int pl = 0; pl < nout + 2; ++pl, ++k) {
                    mbar_wait(&full[k % kSlots], (k / kSlots) & 1, err, 3);
                    if (pl < nout) {                           // output plane pl gets its first contribution: its block must be free
                        const uint32_t jn = j0 + pl;
                        mbar_wait(&blk_empty[jn % R], ((jn / R) & 1) ^ 1, err, 4);
                    }
                    fence_proxy_async();                       // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
                    tc_fence_after();
                    const int od_lo = pl >= 2 ? pl - 2 : 0, od_hi = pl < nout ? pl : nout - 1;
                    const int nb = od_hi - od_lo + 1;                          // output planes this input plane feeds (1..3)
                    const int row0 = (2 - (pl - od_lo)) * NT;                  // weight rows are ordered kd = 2, 1, 0
                    const int blk0 = (int)((j0 + od_lo) % R);
                    const int n1 = blk0 + nb > R ? R - blk0 : nb;              // blocks before the ring wraps
                    const uint64_t a0 = umma_desc(slab_addr + (k % kSlots) * slot_bytes + tile_off, (uint32_t)p.slab_e * 16, 128);
#pragma unroll
                    for (int t9 = 0; t9 < 9; ++t9) {
                        const uint64_t a1 = a0 + (uint64_t)toff[t9];           // address field is in 16 B units
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t ad = a1 + (uint64_t)(uint32_t)(2 * ks * p.slab_e);
                            const uint64_t bd = b0 + (uint64_t)(uint32_t)((t9 * NCH + 2 * ks) * 3 * NT + row0);
                            umma_f16(tmem_base + blk0 * NT, ad, bd, umma_idesc(kTileM, n1 * NT), 1u);
                            if (n1 < nb) umma_f16(tmem_base, ad, bd + (uint64_t)(uint32_t)(n1 * NT), umma_idesc(kTileM, (nb - n1) * NT), 1u);
                        }
                    }
                    umma_commit(&empty[k % kSlots]);           // every input plane is consumed exactly once
                    if (pl >= 2) umma_commit(&blk_full[(j0 + pl - 2) % R]);    // output plane pl-2 is complete
                }
                j0 += nout;
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue (warps 0-3 = TMEM lane quadrants) ===============================
        const int cout = p.CO0 + p.CO1;
        const int cb0 = nt * NT;
        const int creal = min(NT, cout - cb0);                 // real channels in this tile (multiple of 8)
        const float4* bias4 = bias != nullptr ? reinterpret_cast<const float4*>(bias + (size_t)g * cout + cb0) : nullptr;
        uint32_t j = 0, prev_blk = 0;
        bool have_prev = false;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int dc = it % p.ND, r1 = it / p.ND;
            const int qt = r1 % p.QT, n = g * p.npg + r1 / p.QT;
            const int d0 = dc * p.DCH;
            const int nout = min(p.DCH, p.D - d0);
            const int f = qt * kTileM + warp * 32 + lane;
            const int h = f / p.PW, w = f - h * p.PW;
            const bool valid = h < p.H && w < p.W;
            // inset mode: voxels that need no folding go straight to their final place, the rest to the extended buffer
            const int hi = h - 1, wi = w - 1;
            const bool hw_inside = hi >= 0 && hi < Hi && wi >= 0 && wi < Wi;
            const bool hw_shell = hi == 1 || hi == Hi - 2 || wi == 1 || wi == Wi - 2;
            float s1[NT], s2[NT];
#pragma unroll
            for (int c = 0; c < NT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
            for (int od = 0; od < nout; ++od, ++j) {
                const uint32_t blk = j % R;
                mbar_wait(&blk_full[blk], (j / R) & 1, err, 5);
                tc_fence_after();
                float v[NT];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + blk * NT;
#pragma unroll
                for (int c = 0; c < NT; c += 16) tmem_ld16(taddr + c, v + c);
                // hand the PREVIOUS block back now: its zero-fill (issued one plane ago) has had a whole plane of epilogue work
                // to land, so the wait below is free; the ring is R = 256 / NT blocks deep, one plane of delay costs nothing
                if (have_prev) {
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive(&blk_empty[prev_blk]);
                }
#pragma unroll
                for (int c = 0; c < NT; c += 16) tmem_zero16(taddr + c);       // ready for the output plane R planes later
                prev_blk = blk; have_prev = true;
                if (bias4 != nullptr) {                       // L1-resident broadcast loads; not kept in registers
#pragma unroll
                    for (int c4 = 0; c4 < NT / 4; ++c4) {
                        if (c4 * 4 < creal) {
                            const float4 b = __ldg(bias4 + c4);
                            v[c4 * 4] += b.x; v[c4 * 4 + 1] += b.y; v[c4 * 4 + 2] += b.z; v[c4 * 4 + 3] += b.w;
                        }
                    }
                }
                if (valid) {
                    size_t vox = (((size_t)n * p.D + d0 + od) * p.H + h) * p.W + w;
                    bool to_ext = false;
                    if (p.inset) {
                        const int di = d0 + od - 1;
                        if (hw_inside && di >= 0 && di < Di && !(hw_shell || di == 1 || di == Di - 2))
                            vox = (((size_t)n * Di + di) * Hi + hi) * Wi + wi;
                        else
                            to_ext = true;
                    }
#pragma unroll
                    for (int c8 = 0; c8 < NT / 8; ++c8) {
                        if (c8 * 8 < creal) {
                            const int cb = cb0 + c8 * 8;
                            bf16* dst = to_ext ? yext + vox * cout + cb
                                               : (cb < p.CO0 ? y0 + vox * p.CO0 + cb : y1 + vox * p.CO1 + (cb - p.CO0));
                            VecIO<bf16, 8>::store(dst, v + c8 * 8);
#pragma unroll
                            for (int c = 0; c < 8; ++c) { s1[c8 * 8 + c] += v[c8 * 8 + c]; s2[c8 * 8 + c] += v[c8 * 8 + c] * v[c8 * 8 + c]; }
                        }
                    }
                }
            }
            if (stats != nullptr) {
#pragma unroll
                for (int c = 0; c < NT; ++c) {
                    const float a = warp_sum(s1[c]), b = warp_sum(s2[c]);
                    if (lane == 0 && c < creal) {
                        atomicAdd(&stats[((size_t)n * cout + cb0 + c) * 2], (double)a);
                        atomicAdd(&stats[((size_t)n * cout + cb0 + c) * 2 + 1], (double)b);
                    }
                }
            }
        }
        if (have_prev) tmem_wait_st();                         // the last zero-fill must land before the columns are released
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

template <int NCHR, int NT>
int launch_tc(const CUtensorMap& tmap, const CUtensorMap& tmap1, const TcP& p, const void* x0, const void* x1, const void* wimg, const float* bias, void* y0,
              void* y1, void* yext, double* stats, int* err, cudaStream_t st) {
    constexpr int NCH = NCHR < 2 ? 2 : NCHR;
    constexpr int kSlots = NCHR >= 16 ? 2 : 6;
    const size_t w_bytes = (size_t)27 * NCH * NT * 16;
    const size_t smem = ((w_bytes + 127) & ~(size_t)127) + (size_t)kSlots * NCH * p.slab_e * 16 + (2 * kSlots + 2 * 16 + 1) * 8 + 16;
    auto kern = conv3_tc_kernel<NCHR, NT>;
    if (smem > 227 * 1024) { pb_set_error("conv3d_tc: needs %zu B of shared memory", smem); return PB_EUNSUPPORTED; }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("conv3d_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    const int items = p.npg * p.QT * p.ND;
    int ctas = 148 / (p.groups * p.nt_tiles);
    if (smem <= 110 * 1024) ctas *= 2;
    if (ctas < 1) ctas = 1;
    if (ctas > items) ctas = items;
    kern<<<dim3(ctas, p.groups, p.nt_tiles), kTcThreads, smem, st>>>(tmap, tmap1, p, (const bf16*)x0, (const bf16*)x1, (const bf16*)wimg, bias,
                                                                    (bf16*)y0, (bf16*)y1, (bf16*)yext, stats, err);
    return 0;
}


// ------------------------------------------------------------------------------------ kw-stacked variant (Cout <= 16)
// Same sweep as conv3_tc_kernel, but the three kw taps are stacked along N as well: the MMAs of one input plane are
// 3 (kh) x K-steps instructions with N = 3 (kd) x 3 (kw) x 16 = 144 instead of 9 x K-steps with N = 48 — an MMA with N <= 48
// costs the tensor pipe the same ~40-75 cycles as one with N = 144 (scripts/microbench/umma_rate.cu), so the pipe time per
// plane drops ~3x.  The price is a shifted sum in the epilogue: accumulator row l holds, per kw, the contribution of INPUT
// position l as tap kw, which belongs to output position l - kw, i.e.
//     y[l] = D[l][kw=0] + D[l+1][kw=1] + D[l+2][kw=2].
// A one/two-row shift is a warp shuffle; the two rows a warp needs from its right neighbour (TMEM lane quadrants are private to
// a warp) travel through 1.5 KB of shared memory, double-buffered, one named barrier per plane among the four epilogue warps.
// Rows 126/127 of a tile only serve as right neighbours: tiles advance by 126 positions.
// Ring block of one output plane = 48 columns [kw][16 co]; weight image [3 kh][chunk][rows: kd = 2,1,0 | kw | co][8].
constexpr int kKwsStride = 126;


template <int NCHR, int CR>       // CR = real output channels of the tile / 8 (1 or 2)
__global__ void __launch_bounds__(kTcThreads, 2) conv3_tc_kws_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap1, TcP p,
                                                                   const bf16* __restrict__ x0, const bf16* __restrict__ x1,
                                                                   const bf16* __restrict__ wimg, const float* __restrict__ bias,
                                                                   bf16* __restrict__ y0, bf16* __restrict__ y1, bf16* __restrict__ yext,
                                                                   double* __restrict__ stats, int* err) {
    constexpr int NT = 16, NB = 3 * NT, CRE = 8 * CR;
    constexpr int NCH = NCHR < 2 ? 2 : NCHR;
    constexpr int KS = NCH / 2;
    constexpr int kSlots = NCHR >= 16 ? 2 : 6;
    constexpr int TMEM_COLS = 256;
    constexpr int R = TMEM_COLS / NB;                     // 5 accumulator blocks: 3 being written, 2 draining
    const int Di = p.D - 2 * p.inset, Hi = p.H - 2 * p.inset, Wi = p.W - 2 * p.inset;
    extern __shared__ __align__(128) uint8_t smem[];
    const int w_bytes = 27 * NCH * NT * 16;
    uint8_t* w_s = smem;
    const int slot_bytes = NCH * p.slab_e * 16;
    uint8_t* slab_s = smem + ((w_bytes + 127) & ~127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(slab_s + (size_t)kSlots * slot_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kSlots;
    uint64_t* blk_full = bars + 2 * kSlots;
    uint64_t* blk_empty = blk_full + R;
    uint64_t* wbar = blk_empty + R;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* xch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // [2][4 warps][3][NT], 16 B aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y, nt = blockIdx.z;
    const int items = p.npg * p.QT * p.ND;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], p.tma ? 1 : kTcProducers); mbar_init(&empty[i], 1); }
        for (int i = 0; i < R; ++i) { mbar_init(&blk_full[i], 1); mbar_init(&blk_empty[i], 128); }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (NCHR < 2) {
        for (int i = threadIdx.x; i < kSlots * p.slab_e; i += kTcThreads) {
            const int s = i / p.slab_e, e = i % p.slab_e;
            *reinterpret_cast<uint4*>(slab_s + (size_t)s * slot_bytes + ((size_t)p.slab_e + e) * 16) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;          // no zero fill: the first MMA into a block overwrites it (accumulate = 0)

    if (warp >= 5 && p.tma) {
        // =============================== producer: one thread, TMA ===============================
        if (warp == 5 && lane == 0) tma_producer<NCHR, kSlots>(p, &tmap, &tmap1, slab_s, slot_bytes, full, empty, g, items, err);
    } else if (warp >= 5) {
        // =============================== producers (as conv3_tc_kernel; tiles advance by kKwsStride) ===============================
        const int pt = threadIdx.x - 5 * 32;
        const int c0ch = p.C0 >> 3;
        const int copies = p.slab_need * NCHR;
        uint32_t k = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int dc = it % p.ND, r1 = it / p.ND;
            const int qt = r1 % p.QT, n = g * p.npg + r1 / p.QT;
            const int d0 = dc * p.DCH, q0 = qt * kKwsStride;
            const int nout = min(p.DCH, p.D - d0);
            int soff[max_copies(NCHR)];
            uint32_t doff[max_copies(NCHR)];
            uint32_t from1 = 0;
#pragma unroll
            for (int i = 0; i < max_copies(NCHR); ++i) {
                const int idx = pt + i * kTcProducers;
                soff[i] = -1; doff[i] = 0;
                if (idx < copies) {
                    const int ch = idx % NCHR, e = idx / NCHR;
                    const int f = q0 + e;
                    const int hp = f / p.PW, wp = f - hp * p.PW;
                    int h = hp - 1 - p.inset, w = wp - 1 - p.inset;
                    bool ok = hp < p.H + 2;
                    if (p.reflect) { h = reflect_idx(h, Hi); w = reflect_idx(w, Wi); ok = ok && h >= 0 && h < Hi; }
                    else ok = ok && h >= 0 && h < Hi && w >= 0 && w < Wi;
                    doff[i] = (uint32_t)(ch * p.slab_e + e) * 16;
                    if (ok) {
                        if (ch < c0ch) soff[i] = (h * Wi + w) * p.C0 + ch * 8;
                        else { soff[i] = (h * Wi + w) * p.C1 + (ch - c0ch) * 8; from1 |= 1u << i; }
                    }
                }
            }
            for (int pl = 0; pl < nout + 2; ++pl, ++k) {
                const int slot = k % kSlots;
                mbar_wait(&empty[slot], ((k / kSlots) & 1) ^ 1, err, 1);
                int dp = d0 - 1 + pl - p.inset;
                bool plane_ok = true;
                if (p.reflect) dp = reflect_idx(dp, Di); else plane_ok = dp >= 0 && dp < Di;
                if (!plane_ok) dp = 0;
                const size_t plane = ((size_t)n * Di + dp) * Hi * Wi;
                const bf16* p0 = x0 + plane * p.C0;
                const bf16* p1 = x1 + plane * p.C1;
                const uint32_t sbase = smem_u32(slab_s + (size_t)slot * slot_bytes);
#pragma unroll
                for (int i = 0; i < max_copies(NCHR); ++i) {
                    if (pt + i * kTcProducers < copies) {
                        const bool ok = plane_ok && soff[i] >= 0;
                        const bf16* src = ((from1 >> i) & 1u) ? p1 : p0;
                        cp_async16(sbase + doff[i], ok ? src + soff[i] : x0, ok ? 16u : 0u);
                    }
                }
                cp_async_arrive_noinc(&full[slot]);
            }
        }
        cp_async_wait_all();
    } else if (warp == 4) {
        // =============================== MMA issuer: 3 (kh) x KS instructions per input plane ===============================
        if (lane == 0) {
            const bf16* wsrc = wimg + ((size_t)(g * p.nt_tiles + nt) * w_bytes) / 2;
            mbar_expect_tx(wbar, (uint32_t)w_bytes);
            for (int off = 0; off < w_bytes; off += 16384) {
                const int nb = min(16384, w_bytes - off);
                bulk_g2s(smem_u32(w_s + off), reinterpret_cast<const uint8_t*>(wsrc) + off, (uint32_t)nb, wbar);
            }
            mbar_wait(wbar, 0, err, 2);
            const uint32_t slab_addr = smem_u32(slab_s);
            const uint64_t b0 = umma_desc(smem_u32(w_s), 9 * NT * 16, 128);      // LBO = one chunk plane of 9 NT rows
            uint32_t k = 0, j0 = 0;
            long long pw_full = 0, pw_blk = 0, p_issue = 0, p_n = 0;
            const long long p_t0 = clock64();
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int dc = it % p.ND;
                const int nout = min(p.DCH, p.D - dc * p.DCH);
                // TMA slabs start at a padded-row boundary: the tile begins (q0 mod PW) rows into the slab
                const uint32_t tile_off = p.tma ? (uint32_t)((((it / p.ND) % p.QT) * p.q_stride) % p.PW) * 16u : 0u;
                for (int pl = 0; pl < nout + 2; ++pl, ++k) {
                    const long long c0 = clock64();
                    mbar_wait(&full[k % kSlots], (k / kSlots) & 1, err, 3);
                    const long long c1 = clock64();
                    if (pl < nout) {
                        const uint32_t jn = j0 + pl;
                        mbar_wait(&blk_empty[jn % R], ((jn / R) & 1) ^ 1, err, 4);
                    }
                    const long long c2 = clock64();
                    pw_full += c1 - c0; pw_blk += c2 - c1; ++p_n;
                    fence_proxy_async();
                    tc_fence_after();
                    const int od_lo = pl >= 2 ? pl - 2 : 0, od_hi = pl < nout ? pl : nout - 1;
                    const int nb = od_hi - od_lo + 1;
                    const int row0 = (2 - (pl - od_lo)) * NB;                  // weight rows: kd = 2, 1, 0 blocks of [kw][co]
                    const int blk0 = (int)((j0 + od_lo) % R);
                    const uint64_t a0 = umma_desc(slab_addr + (k % kSlots) * slot_bytes + tile_off, (uint32_t)p.slab_e * 16, 128);
                    // output plane pl (if any) gets its FIRST contribution from this input plane: that block is written with
                    // accumulate = 0 by the first instruction, so the epilogue never has to zero a block (and hands it back as soon
                    // as it has been read); every other (block, instruction) accumulates
                    const bool has_new = pl < nout;
                    const int nb_old = nb - (has_new ? 1 : 0);
                    auto issue = [&](uint64_t ad, uint64_t bd, int first_blk, int nblk, uint32_t acc) {     // nblk consecutive ring blocks
                        const int b0i = (blk0 + first_blk) % R;
                        const int m1 = b0i + nblk > R ? R - b0i : nblk;
                        const uint64_t bdd = bd + (uint64_t)(uint32_t)(first_blk * NB);
                        umma_f16(tmem_base + b0i * NB, ad, bdd, umma_idesc(kTileM, m1 * NB), acc);
                        if (m1 < nblk) umma_f16(tmem_base, ad, bdd + (uint64_t)(uint32_t)(m1 * NB), umma_idesc(kTileM, (nblk - m1) * NB), acc);
                    };
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const uint64_t a1 = a0 + (uint64_t)(uint32_t)(kh * p.PW);
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t ad = a1 + (uint64_t)(uint32_t)(2 * ks * p.slab_e);
                            const uint64_t bd = b0 + (uint64_t)(uint32_t)((kh * NCH + 2 * ks) * 9 * NT + row0);
                            if (kh == 0 && ks == 0 && has_new) {
                                if (nb_old > 0) issue(ad, bd, 0, nb_old, 1u);
                                issue(ad, bd, nb_old, 1, 0u);
                            } else {
                                issue(ad, bd, 0, nb, 1u);
                            }
                        }
                    }
                    p_issue += clock64() - c2;
                    umma_commit(&empty[k % kSlots]);
                    if (pl >= 2) umma_commit(&blk_full[(j0 + pl - 2) % R]);
                }
                j0 += nout;
            }
            if (p.probe) {
                atomicAdd(&tc_dbg[0], (unsigned long long)pw_full); atomicAdd(&tc_dbg[1], (unsigned long long)pw_blk);
                atomicAdd(&tc_dbg[2], (unsigned long long)p_issue); atomicAdd(&tc_dbg[3], (unsigned long long)p_n);
                atomicAdd(&tc_dbg[4], (unsigned long long)(clock64() - p_t0)); atomicAdd(&tc_dbg[5], 1ULL);
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue ===============================
        const int cout = p.CO0 + p.CO1;
        const int cb0 = nt * NT;
        const float4* bias4 = bias != nullptr ? reinterpret_cast<const float4*>(bias + (size_t)g * cout + cb0) : nullptr;
        uint32_t j = 0;
        const int tl = warp * 32 + lane;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int dc = it % p.ND, r1 = it / p.ND;
            const int qt = r1 % p.QT, n = g * p.npg + r1 / p.QT;
            const int d0 = dc * p.DCH;
            const int nout = min(p.DCH, p.D - d0);
            const int f = qt * kKwsStride + tl;
            const int h = f / p.PW, w = f - h * p.PW;
            const bool valid = tl < kKwsStride && h < p.H && w < p.W;
            const int hi = h - 1, wi = w - 1;
            const bool hw_inside = hi >= 0 && hi < Hi && wi >= 0 && wi < Wi;
            const bool hw_shell = hi == 1 || hi == Hi - 2 || wi == 1 || wi == Wi - 2;
            float s1[CRE], s2[CRE];
#pragma unroll
            for (int c = 0; c < CRE; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
            for (int od = 0; od < nout; ++od, ++j) {
                const uint32_t blk = j % R;
                const long long e0 = clock64();
                mbar_wait(&blk_full[blk], (j / R) & 1, err, 5);
                if (p.probe && threadIdx.x == 0) atomicAdd(&tc_dbg[6], (unsigned long long)(clock64() - e0));
                tc_fence_after();
                float v0[CRE], v1[CRE], v2[CRE];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + blk * NB;
#pragma unroll
                for (int c = 0; c < CRE; c += 8) {
                    tmem_ld8(taddr + c, v0 + c); tmem_ld8(taddr + NT + c, v1 + c); tmem_ld8(taddr + 2 * NT + c, v2 + c);
                }
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(&blk_empty[blk]);                  // the block is in registers: the MMA thread may overwrite it
                // rows 0 / 1 of this warp are rows 32 / 33 of its left neighbour
                if (lane < 2) {
                    float4* xw = reinterpret_cast<float4*>(xch + ((j & 1) * 4 + warp) * 3 * NT);
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < CRE; c += 4) {
                            xw[c / 4] = make_float4(v1[c], v1[c + 1], v1[c + 2], v1[c + 3]);
                            xw[NT / 4 + c / 4] = make_float4(v2[c], v2[c + 1], v2[c + 2], v2[c + 3]);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < CRE; c += 4) xw[2 * NT / 4 + c / 4] = make_float4(v2[c], v2[c + 1], v2[c + 2], v2[c + 3]);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const float4* xn = reinterpret_cast<const float4*>(xch + ((j & 1) * 4 + ((warp + 1) & 3)) * 3 * NT);
                float y[CRE];
#pragma unroll
                for (int c = 0; c < CRE; c += 4) {              // broadcast loads, selects instead of branches
                    const float4 n0 = xn[c / 4], n1 = xn[NT / 4 + c / 4], n2 = xn[2 * NT / 4 + c / 4];
                    const float e0[4] = {n0.x, n0.y, n0.z, n0.w}, e1[4] = {n1.x, n1.y, n1.z, n1.w}, e2[4] = {n2.x, n2.y, n2.z, n2.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float a1 = __shfl_down_sync(0xffffffffu, v1[c + i], 1);
                        float a2 = __shfl_down_sync(0xffffffffu, v2[c + i], 2);
                        a1 = lane == 31 ? e0[i] : a1;
                        a2 = lane == 31 ? e2[i] : (lane == 30 ? e1[i] : a2);
                        y[c + i] = v0[c + i] + a1 + a2;
                    }
                }
                if (bias4 != nullptr) {
#pragma unroll
                    for (int c4 = 0; c4 < CRE / 4; ++c4) {
                        const float4 b = __ldg(bias4 + c4);
                        y[c4 * 4] += b.x; y[c4 * 4 + 1] += b.y; y[c4 * 4 + 2] += b.z; y[c4 * 4 + 3] += b.w;
                    }
                }
                if (valid) {
                    size_t vox = (((size_t)n * p.D + d0 + od) * p.H + h) * p.W + w;
                    bool to_ext = false;
                    if (p.inset) {
                        const int di = d0 + od - 1;
                        if (hw_inside && di >= 0 && di < Di && !(hw_shell || di == 1 || di == Di - 2))
                            vox = (((size_t)n * Di + di) * Hi + hi) * Wi + wi;
                        else
                            to_ext = true;
                    }
#pragma unroll
                    for (int c8 = 0; c8 < CR; ++c8) {
                        const int cb = cb0 + c8 * 8;
                        bf16* dst = to_ext ? yext + vox * cout + cb
                                           : (cb < p.CO0 ? y0 + vox * p.CO0 + cb : y1 + vox * p.CO1 + (cb - p.CO0));
                        VecIO<bf16, 8>::store(dst, y + c8 * 8);
#pragma unroll
                        for (int c = 0; c < 8; ++c) { s1[c8 * 8 + c] += y[c8 * 8 + c]; s2[c8 * 8 + c] += y[c8 * 8 + c] * y[c8 * 8 + c]; }
                    }
                }
            }
            if (stats != nullptr) {
#pragma unroll
                for (int c = 0; c < CRE; ++c) {
                    const float a = warp_sum(s1[c]), b = warp_sum(s2[c]);
                    if (lane == 0) {
                        atomicAdd(&stats[((size_t)n * cout + cb0 + c) * 2], (double)a);
                        atomicAdd(&stats[((size_t)n * cout + cb0 + c) * 2 + 1], (double)b);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

template <int NCHR, int CR>
int launch_tc_kws(const CUtensorMap& tmap, const CUtensorMap& tmap1, const TcP& p, const void* x0, const void* x1, const void* wimg, const float* bias, void* y0,
                  void* y1, void* yext, double* stats, int* err, cudaStream_t st) {
    constexpr int NCH = NCHR < 2 ? 2 : NCHR;
    constexpr int kSlots = NCHR >= 16 ? 2 : 6;
    const size_t w_bytes = (size_t)27 * NCH * 16 * 16;
    const size_t smem = ((w_bytes + 127) & ~(size_t)127) + (size_t)kSlots * NCH * p.slab_e * 16 + (2 * kSlots + 2 * 16 + 1) * 8 + 16
                        + 2 * 4 * 3 * 16 * sizeof(float) + 16;
    auto kern = conv3_tc_kws_kernel<NCHR, CR>;
    if (smem > 227 * 1024) { pb_set_error("conv3d_tc_kws: needs %zu B of shared memory", smem); return PB_EUNSUPPORTED; }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("conv3d_tc_kws: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    const int items = p.npg * p.QT * p.ND;
    int ctas = 148 / (p.groups * p.nt_tiles);
    if (smem <= 113 * 1024) ctas *= 2;           // two CTAs per SM: 2 x (smem + 1 KB reserved per CTA) <= 228 KB
    if (ctas < 1) ctas = 1;
    if (ctas > items) ctas = items;
    kern<<<dim3(ctas, p.groups, p.nt_tiles), kTcThreads, smem, st>>>(tmap, tmap1, p, (const bf16*)x0, (const bf16*)x1, (const bf16*)wimg, bias,
                                                                    (bf16*)y0, (bf16*)y1, (bf16*)yext, stats, err);
    return 0;
}


// ------------------------------------------------------------------------------------ weight gradient on tcgen05
// dw[tap][ci][co] = sum_{n,q} x[n, q + off(tap)][ci] * dy[n, q][co]  (reflect / zero padding resolved when staging x).
// GEMM view with the VOXEL index as K:  D[co, (kw, ci)] += A[co, q] * B[(kw, ci), q]  for every (kd, kh), where
//   A = dy tile  [128 rows q][co]  read as an MN-major operand (co contiguous, q strided by 16 B),
//   B = x slab   rows q + kh*PW + kw + ...  read as an MN-major operand whose N-groups are the kw = 0..3 shifts of the
//       same 8-channel plane (SBO = 16 B: a one-row shift) — again no im2col copy in shared memory.
// One CTA owns one 8-channel chunk of the input and keeps all nine (kd,kh) accumulators (128 lanes x 32 columns each)
// resident in TMEM (M = 64: rows 16w..16w+15 in lanes 32w..32w+15) while it sweeps its share of the volume; they are read out once at the end and added to dw with fp32
// atomics.  M is 64 although only `co` rows are meaningful: rows beyond co read don't-care shared memory and are never
// stored (each D row depends on its own A row only).
constexpr int kWgSlotsX = 6, kWgSlotsY = 3;
constexpr int kWgM = 64;                              // UMMA M: Cout <= 64 rows are meaningful, M = 64 halves the A-operand fetch
constexpr int kWgYPlane = (kTileM + 1) * 16;          // co-chunk plane pitch: 129 rows, so the 8 M-groups of one K row hit 8 different banks
constexpr int kWgYSlotBytes = (kWgM / 8) * kWgYPlane;  // 8 M-groups (co chunks)

struct WgP {
    int N, D, H, W, C0, C1, Cout, reflect;
    int PW, QT, DCH, ND, npg, groups, nchunks;
    int slab_need, slab_e;
};

__host__ __device__ constexpr uint32_t umma_idesc_mn(int M, int N) {       // both operands MN-major
    return umma_idesc(M, N) | (1u << 15) | (1u << 16);
}

template <int NCO>   // NCO = Cout / 8
__global__ void __launch_bounds__(kThreads, 1) conv3_wgrad_tc_kernel(WgP p, const bf16* __restrict__ x0, const bf16* __restrict__ x1,
                                                                     const bf16* __restrict__ dy, float* __restrict__ dw, int* err) {
    constexpr uint32_t IDESC = umma_idesc_mn(kWgM, 32);
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* y_s = smem;                                              // [kWgSlotsY][8 planes][128 rows][16 B]
    uint8_t* x_s = smem + kWgSlotsY * kWgYSlotBytes;                  // [kWgSlotsX][slab_e rows][16 B]
    const int xslot_bytes = p.slab_e * 16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(x_s + (size_t)kWgSlotsX * xslot_bytes);
    uint64_t* fullx = bars;
    uint64_t* emptyx = bars + kWgSlotsX;
    uint64_t* fully = bars + 2 * kWgSlotsX;
    uint64_t* emptyy = fully + kWgSlotsY;
    uint64_t* done = emptyy + kWgSlotsY;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y / p.nchunks, chunk = blockIdx.y % p.nchunks;
    const int items = p.npg * p.QT * p.ND;

    if (threadIdx.x == 0) {
        // three MMA issuers (one per kd, as in conv3_wgrad_tc8_kernel): one arrival from each on the consumer-side barriers
        for (int i = 0; i < kWgSlotsX; ++i) { mbar_init(&fullx[i], kProducerThreads); mbar_init(&emptyx[i], 3); }
        for (int i = 0; i < kWgSlotsY; ++i) { mbar_init(&fully[i], kProducerThreads); mbar_init(&emptyy[i], 3); }
        mbar_init(done, 3);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 5) {
        // =============================== producers ===============================
        const int pt = threadIdx.x - 5 * 32;
        const int c0ch = p.C0 >> 3;
        const bool from1 = chunk >= c0ch;
        const bf16* xsrc = from1 ? x1 : x0;
        const int cs = from1 ? p.C1 : p.C0, coff = (from1 ? chunk - c0ch : chunk) * 8;
        uint32_t kx = 0, ky = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int dc = it % p.ND, r1 = it / p.ND;
            const int qt = r1 % p.QT, n = g * p.npg + r1 / p.QT;
            const int d0 = dc * p.DCH, q0 = qt * kTileM;
            const int nout = min(p.DCH, p.D - d0);
            int soff[kWgCopies];                           // x slab rows of this thread (slab_need <= kWgCopies * 96)
            int yoff[2 * NCO];                             // dy copies of this thread (128 * NCO <= 2 * NCO * 96)
#pragma unroll
            for (int i = 0; i < kWgCopies; ++i) {
                const int e = pt + i * kProducerThreads;
                soff[i] = -1;
                if (e < p.slab_need) {
                    const int f = q0 + e;
                    const int hp = f / p.PW, wp = f - hp * p.PW;
                    int h = hp - 1, w = wp - 1;
                    bool ok = hp < p.H + 2;
                    if (p.reflect) { h = reflect_idx(h, p.H); w = reflect_idx(w, p.W); ok = ok && h >= 0 && h < p.H && w >= 0 && w < p.W; }
                    else ok = ok && h >= 0 && h < p.H && w >= 0 && w < p.W;
                    if (ok) soff[i] = (h * p.W + w) * cs + coff;
                }
            }
#pragma unroll
            for (int i = 0; i < 2 * NCO; ++i) {
                const int idx = pt + i * kProducerThreads;
                yoff[i] = -1;
                if (idx < kTileM * NCO) {
                    const int c = idx % NCO, r = idx / NCO;
                    const int f = q0 + r;
                    const int h = f / p.PW, w = f - h * p.PW;
                    if (h < p.H && w < p.W) yoff[i] = (h * p.W + w) * p.Cout + c * 8;
                }
            }
            for (int pl = 0; pl < nout + 2; ++pl, ++kx) {
                const int slot = kx % kWgSlotsX;
                mbar_wait(&emptyx[slot], ((kx / kWgSlotsX) & 1) ^ 1, err, 11);
                int dp = d0 - 1 + pl;
                bool plane_ok = true;
                if (p.reflect) dp = reflect_idx(dp, p.D); else plane_ok = dp >= 0 && dp < p.D;
                if (!plane_ok) dp = 0;
                const bf16* pp = xsrc + ((size_t)n * p.D + dp) * p.H * p.W * cs;
                const uint32_t sbase = smem_u32(x_s + (size_t)slot * xslot_bytes);
#pragma unroll
                for (int i = 0; i < kWgCopies; ++i) {
                    const int e = pt + i * kProducerThreads;
                    if (e < p.slab_need) {
                        const bool ok = plane_ok && soff[i] >= 0;
                        cp_async16(sbase + (uint32_t)e * 16, ok ? pp + soff[i] : x0, ok ? 16u : 0u);
                    }
                }
                cp_async_arrive_noinc(&fullx[slot]);
                // the dy plane that pairs with x planes pl-1, pl, pl+1 is output plane d0 + pl - 1: stage it one step late
                if (pl >= 1 && pl <= nout) {
                    const int ys = ky % kWgSlotsY;
                    mbar_wait(&emptyy[ys], ((ky / kWgSlotsY) & 1) ^ 1, err, 12);
                    const bf16* py = dy + ((size_t)n * p.D + d0 + pl - 1) * p.H * p.W * p.Cout;
                    const uint32_t ybase = smem_u32(y_s + (size_t)ys * kWgYSlotBytes);
#pragma unroll
                    for (int i = 0; i < 2 * NCO; ++i) {
                        const int idx = pt + i * kProducerThreads;
                        if (idx < kTileM * NCO) {
                            const int c = idx % NCO, r = idx / NCO;
                            const bool ok = yoff[i] >= 0;
                            cp_async16(ybase + (uint32_t)(c * kWgYPlane + r * 16), ok ? py + yoff[i] : dy, ok ? 16u : 0u);
                        }
                    }
                    cp_async_arrive_noinc(&fully[ys]);
                    ++ky;
                }
            }
        }
        cp_async_wait_all();
    }
    // =============================== MMA issuers: lane 0 of warp 4 (kd = 0) and of the idle epilogue warps 0, 1 (kd = 1, 2) ===
    // (see conv3_wgrad_tc8_kernel for the protocol: every issuer sees every x plane's `full` phase and releases it itself)
    const int my_kd = warp == 4 ? 0 : (warp == 0 ? 1 : (warp == 1 ? 2 : -1));
    if (my_kd >= 0) {
        if (lane == 0) {
            const uint32_t x_addr = smem_u32(x_s), y_addr = smem_u32(y_s);
            uint32_t kx = 0, ky = 0;
            bool first = true;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int dc = it % p.ND;
                const int nout = min(p.DCH, p.D - dc * p.DCH);
                for (int pl = 0; pl < nout + 2; ++pl) {
                    const uint32_t kk = kx + pl;
                    mbar_wait(&fullx[kk % kWgSlotsX], (kk / kWgSlotsX) & 1, err, 13);
                    const int od = pl - my_kd;
                    if (od >= 0 && od < nout) {
                        const uint32_t ko = ky + od;
                        mbar_wait(&fully[ko % kWgSlotsY], (ko / kWgSlotsY) & 1, err, 14);
                        fence_proxy_async();
                        tc_fence_after();
                        const uint64_t a0 = umma_desc(y_addr + (ko % kWgSlotsY) * kWgYSlotBytes, 128, kWgYPlane);   // LBO = 8 rows, SBO = co-chunk plane
                        const uint64_t b0 = umma_desc(x_addr + (kk % kWgSlotsX) * xslot_bytes, 128, 16);            // LBO = 8 rows, SBO = one-row (kw) shift
                        // consecutive MMAs go to DIFFERENT accumulators (the three kh tiles of this kd)
#pragma unroll 1
                        for (int ks = 0; ks < kTileM / 16; ++ks) {
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
                                const uint32_t d_tmem = tmem_base + (my_kd * 3 + kh) * 32;
                                umma_f16(d_tmem, a0 + (uint64_t)(16 * ks), b0 + (uint64_t)(uint32_t)(kh * p.PW + 16 * ks), IDESC,
                                         (first && ks == 0) ? 0u : 1u);
                            }
                        }
                        first = false;
                        umma_commit(&emptyy[ko % kWgSlotsY]);
                    }
                    umma_commit(&emptyx[kk % kWgSlotsX]);
                }
                kx += nout + 2;
                ky += nout;
            }
            umma_commit(done);
        }
        __syncwarp();
    }
    if (warp < 4) {
        // =============================== epilogue: TMEM -> fp32 atomics into dw ===============================
        if (blockIdx.x < items) {
            mbar_wait(done, 0, err, 15);
            tc_fence_after();
            // M = 64 accumulator layout (cute tmem_frg, "half subpartitions"): row m lives in lane (m % 16) + 32 * (m / 16)
            const int co = lane < 16 ? warp * 16 + lane : p.Cout;
            const int cin = p.C0 + p.C1;
            for (int a = 0; a < 9; ++a) {
                float v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + a * 32;
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
                if (co < p.Cout) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int tap = a * 3 + kw;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            atomicAdd(dw + (((size_t)g * 27 + tap) * cin + chunk * 8 + j) * p.Cout + co, v[kw * 8 + j]);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- small-Cout specialisation (Cout = 8 * NCO, NCO in {1, 2, 4}): the three kh taps are stacked along M ---------
// Per 8-channel output chunk the M-groups of the A operand carry the kh shifts of dy:
//     D_kd[(kh, co), (kw, ci)] = sum_u dy[u - kh*PW][co] * x[plane d+kd-1, u + kw][ci]      (u = padded-plane position)
// i.e. A = the dy slab read with SBO = PW*16 B (one padded row per M-group), B = the x slab with SBO = 16 B (kw).
// 24 tcgen05.mma per plane and output chunk instead of 72, three 64x32 accumulators (96 TMEM columns) per chunk
// instead of nine; the x slab (B operand) is shared by the NCO chunks, the dy slab holds one plane per chunk.
constexpr int kWg8SlotsX = 6, kWg8SlotsY = 3;
constexpr bool kWgStackCout16 = true, kWgStackCout32 = false;     // which Cout classes use the kh-stacked kernel (measured:
                                                                  // Cout 16: 2.4x faster than the general kernel; Cout 32: no gain)

template <int NCO>
__global__ void __launch_bounds__(kThreads, 2) conv3_wgrad_tc8_kernel(WgP p, const bf16* __restrict__ x0, const bf16* __restrict__ x1,
                                                                      const bf16* __restrict__ dy, float* __restrict__ dw, int* err) {
    constexpr uint32_t IDESC = umma_idesc_mn(kWgM, 32);
    extern __shared__ __align__(128) uint8_t smem[];
    const int yrows = kTileM + 2 * p.PW;                              // dy slab: positions u0 - 2*PW .. u0 + 127
    const int yplane_bytes = ((yrows + 7) & ~7) * 16;                 // one output chunk
    const int yslot_bytes = NCO * yplane_bytes;
    constexpr uint32_t TMEM_COLS = NCO == 1 ? 128u : (NCO == 2 ? 256u : 512u);
    const int xslot_bytes = p.slab_e * 16;                            // x slab: positions u0 .. u0 + 130
    uint8_t* y_s = smem;
    uint8_t* x_s = smem + kWg8SlotsY * yslot_bytes;
    uint8_t* pad_s = x_s + (size_t)kWg8SlotsX * xslot_bytes;          // head-room for the don't-care M-groups 3..7
    uint64_t* bars = reinterpret_cast<uint64_t*>(pad_s + 8 * p.PW * 16);
    uint64_t* fullx = bars;
    uint64_t* emptyx = bars + kWg8SlotsX;
    uint64_t* fully = bars + 2 * kWg8SlotsX;
    uint64_t* emptyy = fully + kWg8SlotsY;
    uint64_t* done = emptyy + kWg8SlotsY;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y / p.nchunks, chunk = blockIdx.y % p.nchunks;
    const int items = p.npg * p.QT * p.ND;

    if (threadIdx.x == 0) {
        // three MMA issuers (one per kd, see below): every consumer-side barrier expects one arrival from each of them
        for (int i = 0; i < kWg8SlotsX; ++i) { mbar_init(&fullx[i], kProducerThreads); mbar_init(&emptyx[i], 3); }
        for (int i = 0; i < kWg8SlotsY; ++i) { mbar_init(&fully[i], kProducerThreads); mbar_init(&emptyy[i], 3); }
        mbar_init(done, 3);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 5) {
        // =============================== producers ===============================
        const int pt = threadIdx.x - 5 * 32;
        const int c0ch = p.C0 >> 3;
        const bool from1 = chunk >= c0ch;
        const bf16* xsrc = from1 ? x1 : x0;
        const int cs = from1 ? p.C1 : p.C0, coff = (from1 ? chunk - c0ch : chunk) * 8;
        uint32_t kx = 0, ky = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int dc = it % p.ND, r1 = it / p.ND;
            const int qt = r1 % p.QT, n = g * p.npg + r1 / p.QT;
            const int d0 = dc * p.DCH, u0 = qt * kTileM;
            const int nout = min(p.DCH, p.D - d0);
            int soff[2];                                   // x slab rows of this thread (131 <= 2 * 96)
            int yoff[kWgCopies];                           // dy slab rows of this thread (128 + 2*PW <= kWgCopies * 96)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int e = pt + i * kProducerThreads;
                soff[i] = -1;
                if (e < p.slab_need) {
                    const int f = u0 + e;
                    const int hp = f / p.PW, wp = f - hp * p.PW;
                    int h = hp - 1, w = wp - 1;
                    bool ok = hp < p.H + 2;
                    if (p.reflect) { h = reflect_idx(h, p.H); w = reflect_idx(w, p.W); ok = ok && h >= 0 && h < p.H && w >= 0 && w < p.W; }
                    else ok = ok && h >= 0 && h < p.H && w >= 0 && w < p.W;
                    if (ok) soff[i] = (h * p.W + w) * cs + coff;
                }
            }
#pragma unroll
            for (int i = 0; i < kWgCopies; ++i) {
                const int e = pt + i * kProducerThreads;
                yoff[i] = -1;
                if (e < yrows) {
                    const int q = u0 - 2 * p.PW + e;            // output-plane position q = h*PW + w
                    if (q >= 0) {
                        const int h = q / p.PW, w = q - h * p.PW;
                        if (h < p.H && w < p.W) yoff[i] = (h * p.W + w) * p.Cout;         // + 8 * co chunk
                    }
                }
            }
            for (int pl = 0; pl < nout + 2; ++pl, ++kx) {
                const int slot = kx % kWg8SlotsX;
                mbar_wait(&emptyx[slot], ((kx / kWg8SlotsX) & 1) ^ 1, err, 31);
                int dp = d0 - 1 + pl;
                bool plane_ok = true;
                if (p.reflect) dp = reflect_idx(dp, p.D); else plane_ok = dp >= 0 && dp < p.D;
                if (!plane_ok) dp = 0;
                const bf16* pp = xsrc + ((size_t)n * p.D + dp) * p.H * p.W * cs;
                const uint32_t sbase = smem_u32(x_s + (size_t)slot * xslot_bytes);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int e = pt + i * kProducerThreads;
                    if (e < p.slab_need) {
                        const bool ok = plane_ok && soff[i] >= 0;
                        cp_async16(sbase + (uint32_t)e * 16, ok ? pp + soff[i] : x0, ok ? 16u : 0u);
                    }
                }
                cp_async_arrive_noinc(&fullx[slot]);
                if (pl >= 1 && pl <= nout) {
                    const int ys = ky % kWg8SlotsY;
                    mbar_wait(&emptyy[ys], ((ky / kWg8SlotsY) & 1) ^ 1, err, 32);
                    const bf16* py = dy + ((size_t)n * p.D + d0 + pl - 1) * p.H * p.W * p.Cout;
                    const uint32_t ybase = smem_u32(y_s + (size_t)ys * yslot_bytes);
#pragma unroll
                    for (int i = 0; i < kWgCopies; ++i) {
                        const int e = pt + i * kProducerThreads;
                        if (e < yrows) {
                            const bool ok = yoff[i] >= 0;
#pragma unroll
                            for (int cc = 0; cc < NCO; ++cc)
                                cp_async16(ybase + (uint32_t)(cc * yplane_bytes + e * 16), ok ? py + yoff[i] + cc * 8 : dy, ok ? 16u : 0u);
                        }
                    }
                    cp_async_arrive_noinc(&fully[ys]);
                    ++ky;
                }
            }
        }
        cp_async_wait_all();
    }
    // =============================== MMA issuers ===============================
    // A tcgen05.mma with N <= 32 occupies the tensor pipe for ~40-45 cycles however small it is, and ONE thread cannot issue
    // them faster than every ~45-60 cycles (scripts/microbench/umma_rate.cu: 61 -> 31 -> 23 cycles per M=64 MMA with 1 -> 2 -> 4
    // issuing warps).  The three kd taps accumulate into different TMEM tiles, so each gets its own issuing thread: lane 0 of
    // warp 4 (kd = 0) and of the otherwise idle epilogue warps 0 and 1 (kd = 1, 2).  Every issuer walks the input planes in
    // order, waits for each plane (also the ones only the other issuers read: an arrival on `empty` is only legal once the
    // plane's `full` phase has been seen) and releases it with its own tcgen05.commit.
    const int my_kd = warp == 4 ? 0 : (warp == 0 ? 1 : (warp == 1 ? 2 : -1));
    if (my_kd >= 0) {
        if (lane == 0) {
            const uint32_t x_addr = smem_u32(x_s), y_addr = smem_u32(y_s);
            uint32_t kx = 0, ky = 0;
            bool first = true;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int dc = it % p.ND;
                const int nout = min(p.DCH, p.D - dc * p.DCH);
                for (int pl = 0; pl < nout + 2; ++pl) {
                    const uint32_t kk = kx + pl;
                    mbar_wait(&fullx[kk % kWg8SlotsX], (kk / kWg8SlotsX) & 1, err, 33);
                    const int od = pl - my_kd;                           // the output plane this issuer pairs with x plane pl
                    if (od >= 0 && od < nout) {
                        const uint32_t ko = ky + od;
                        mbar_wait(&fully[ko % kWg8SlotsY], (ko / kWg8SlotsY) & 1, err, 34);
                        fence_proxy_async();
                        tc_fence_after();
                        const uint64_t b0 = umma_desc(x_addr + (kk % kWg8SlotsX) * xslot_bytes, 128, 16);
#pragma unroll
                        for (int cc = 0; cc < NCO; ++cc) {
                            // A: M-group g' (= 2 - kh) starts g' padded rows further into the dy slab
                            const uint64_t a0 = umma_desc(y_addr + (ko % kWg8SlotsY) * yslot_bytes + cc * yplane_bytes, 128, (uint32_t)p.PW * 16);
                            const uint32_t d_tmem = tmem_base + (cc * 3 + my_kd) * 32;
#pragma unroll
                            for (int ks = 0; ks < kTileM / 16; ++ks)
                                umma_f16(d_tmem, a0 + (uint64_t)(16 * ks), b0 + (uint64_t)(16 * ks), IDESC, (first && ks == 0) ? 0u : 1u);
                        }
                        first = false;
                        umma_commit(&emptyy[ko % kWg8SlotsY]);
                    }
                    umma_commit(&emptyx[kk % kWg8SlotsX]);
                }
                kx += nout + 2;
                ky += nout;
            }
            umma_commit(done);
        }
        __syncwarp();
    }
    if (warp < 4) {
        // =============================== epilogue ===============================
        if (blockIdx.x < items) {
            mbar_wait(done, 0, err, 35);
            tc_fence_after();
            // M = 64 layout: row m in lane (m % 16) + 32 * (m / 16); m = g'*8 + co with kh = 2 - g'
            const int m = lane < 16 ? warp * 16 + lane : 64;
            const int kh = 2 - (m >> 3), co = m & 7;
            const int cin = p.C0 + p.C1;
            for (int a = 0; a < 3 * NCO; ++a) {
                const int cc = a / 3, kd = a - cc * 3;
                float v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + a * 32;
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
                if (m < 24) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            atomicAdd(dw + (((size_t)g * 27 + tap) * cin + chunk * 8 + j) * p.Cout + cc * 8 + co, v[kw * 8 + j]);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int NCO>
int launch_wgrad_tc8(WgP p, const void* x0, const void* x1, const void* dy, float* dw, int* err, cudaStream_t st) {
    p.QT = ((p.H + 2) * p.PW + kTileM - 1) / kTileM;                  // tiles run over the PADDED plane positions u
    p.slab_need = kTileM + 3;
    p.slab_e = (p.slab_need + 7) & ~7;
    const int yrows = kTileM + 2 * p.PW;
    if (yrows > kWgCopies * kProducerThreads) { pb_set_error("conv3d_wgrad_tc8: dy slab of %d rows exceeds the producer budget", yrows); return PB_EUNSUPPORTED; }
    const int target = 148 * 4 / (p.groups * p.nchunks) + 1;
    int nd = 1;
    while (p.npg * p.QT * nd < target && (p.D + nd) / (nd + 1) >= 8) ++nd;
    p.DCH = (p.D + nd - 1) / nd;
    p.ND = (p.D + p.DCH - 1) / p.DCH;
    const size_t smem = (size_t)kWg8SlotsY * NCO * (((yrows + 7) & ~7) * 16) + (size_t)kWg8SlotsX * p.slab_e * 16 + (size_t)8 * p.PW * 16 +
                        (2 * kWg8SlotsX + 2 * kWg8SlotsY + 1) * 8 + 16;
    auto kern = conv3_wgrad_tc8_kernel<NCO>;
    if (smem > 227 * 1024) { pb_set_error("conv3d_wgrad_tc8: needs %zu B of shared memory", smem); return PB_EUNSUPPORTED; }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("conv3d_wgrad_tc8: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    const int items = p.npg * p.QT * p.ND;
    int ctas = (NCO == 4 ? 1 : 2) * 148 / (p.groups * p.nchunks);     // NCO = 4 takes all 512 TMEM columns: one CTA per SM
    if (ctas < 1) ctas = 1;
    if (ctas > items) ctas = items;
    kern<<<dim3(ctas, p.groups * p.nchunks), kThreads, smem, st>>>(p, (const bf16*)x0, (const bf16*)x1, (const bf16*)dy, dw, err);
    return 0;
}

// ---- 1x1x1 weight gradient on tcgen05 ----------------------------------------------------------------------------
// dw[ci][co] = sum_v x[v][ci] * dy[v][co]: a skinny GEMM whose K is the voxel count, HBM-bound (the FFMA kernels for it are
// FP32-pipe bound at 3.2 FMA per byte).  A = dy tile (MN-major: M = co, 8-channel chunk planes), B = x tile (MN-major:
// N = ci, chunk planes), 128 voxels per tile = 8 K-steps; the [64 x Cin] fp32 accumulator stays in TMEM for the CTA's
// whole stream of tiles.  Tiles are plain contiguous row ranges of the two tensors (no halo), staged by cp.async into a
// ring of slots; one atomic flush per CTA.
constexpr int kW1Plane = (kTileM + 1) * 16;               // chunk-plane pitch (129 rows: spreads the MN-groups over the banks)
constexpr int kW1Producers = kThreads - 32;               // all warps except the MMA issuer stage tiles

struct W1P {
    int C0, C1, Cout;
    long long VT;                                         // voxels per weight group (samples of a group are contiguous)
    int slots, tiles;
};

__global__ void __launch_bounds__(kThreads, 2) conv1_wgrad_tc_kernel(W1P p, const bf16* __restrict__ x0, const bf16* __restrict__ x1,
                                                                     const bf16* __restrict__ dy, float* __restrict__ dw, int* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int cin = p.C0 + p.C1;
    const int nci = cin >> 3, nco = p.Cout >> 3, nch = nci + nco;
    const int slot_bytes = nch * kW1Plane;                // [co chunk planes][ci chunk planes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.slots * slot_bytes + 8 * kW1Plane);   // tail pad: don't-care M-groups
    uint64_t* full = bars;
    uint64_t* empty = bars + p.slots;
    uint64_t* done = empty + p.slots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    // copy table, one packed word per 16-byte piece of a tile: [0,16) source element offset from the tile's first row in its
    // tensor, [16,29) destination offset / 16 inside the slot, [29,31) source tensor (0 dy, 1 x0, 2 x1).  Built once; the
    // producers then issue a copy with a handful of instructions (a per-copy division / 64-bit address chain made the three
    // producer warps of this 1-CTA-per-SM kernel latency-bound at ~1 TB/s).
    uint32_t* table = tmem_slot + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y;
    {
        const int c0ch = p.C0 >> 3;
        for (int idx = threadIdx.x; idx < kTileM * nch; idx += kThreads) {
            const int r = idx / nch, ch = idx - r * nch;
            uint32_t src, dst, sel;
            if (ch < nco) { sel = 0; src = r * p.Cout + ch * 8; dst = (ch * kW1Plane) / 16 + r; }
            else {
                const int ci = ch - nco;
                if (ci < c0ch) { sel = 1; src = r * p.C0 + ci * 8; } else { sel = 2; src = r * p.C1 + (ci - c0ch) * 8; }
                dst = ((nco + ci) * kW1Plane) / 16 + r;
            }
            table[idx] = src | (dst << 16) | (sel << 29);
        }
    }
    // NACC independent accumulators (K-step ks of every tile goes to accumulator ks % NACC): a chain of dependent
    // tcgen05.mma on ONE accumulator runs at the MMA latency (~0.2 us each, 1.6 us per tile measured), not at its throughput
    const int acc_stride = ((cin + 15) / 16) * 16;
    const int nacc = 2;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < nacc * acc_stride) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.slots; ++i) { mbar_init(&full[i], kW1Producers); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bf16* xg0 = x0 + (size_t)g * p.VT * p.C0;
    const bf16* xg1 = p.C1 ? x1 + (size_t)g * p.VT * p.C1 : nullptr;
    const bf16* dyg = dy + (size_t)g * p.VT * p.Cout;

    if (warp != 4) {
        // =============================== producers: every warp but the MMA issuer ===============================
        const int pt = warp < 4 ? threadIdx.x : threadIdx.x - 32;
        const int copies = kTileM * nch;
        uint32_t k = 0;
        for (int it = blockIdx.x; it < p.tiles; it += gridDim.x, ++k) {
            const int slot = k % p.slots;
            mbar_wait(&empty[slot], ((k / p.slots) & 1) ^ 1, err, 41);
            const long long v0 = (long long)it * kTileM;
            const uint32_t sbase = smem_u32(smem + (size_t)slot * slot_bytes);
            const bf16* base[3] = {dyg + v0 * p.Cout, xg0 + v0 * p.C0, p.C1 ? xg1 + v0 * p.C1 : xg0};
            const int rows_ok = (int)min((long long)kTileM, p.VT - v0);       // < 128 only in the last tile
            // consecutive threads take consecutive 16-byte pieces of a row: first the dy row, then the x row(s)
            for (int idx = pt; idx < copies; idx += kW1Producers) {
                const uint32_t e = table[idx];
                const uint32_t dst16 = (e >> 16) & 0x1FFFu;
                const bool ok = rows_ok == kTileM || (int)(idx / nch) < rows_ok;
                cp_async16(sbase + dst16 * 16, ok ? base[e >> 29] + (e & 0xFFFFu) : dy, ok ? 16u : 0u);
            }
            cp_async_arrive_noinc(&full[slot]);
        }
        cp_async_wait_all();
    }
    if (warp == 4) {
        // =============================== MMA issuer: 8 K-steps per tile ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_mn(kWgM, cin);
            uint32_t k = 0;
            bool first = true;
            for (int it = blockIdx.x; it < p.tiles; it += gridDim.x, ++k) {
                const int slot = k % p.slots;
                mbar_wait(&full[slot], (k / p.slots) & 1, err, 42);
                fence_proxy_async();
                tc_fence_after();
                const uint32_t sbase = smem_u32(smem + (size_t)slot * slot_bytes);
                const uint64_t a0 = umma_desc(sbase, 128, kW1Plane);                      // M-groups = co chunk planes
                const uint64_t b0 = umma_desc(sbase + nco * kW1Plane, 128, kW1Plane);     // N-groups = ci chunk planes
#pragma unroll
                for (int ks = 0; ks < kTileM / 16; ++ks)
                    umma_f16(tmem_base + (uint32_t)((ks % nacc) * acc_stride), a0 + (uint64_t)(16 * ks), b0 + (uint64_t)(16 * ks), idesc,
                             (first && ks < nacc) ? 0u : 1u);
                first = false;
                umma_commit(&empty[slot]);
            }
            umma_commit(done);
        }
        __syncwarp();
    } else if (warp < 4) {
        // =============================== epilogue: TMEM -> fp32 atomics into dw [g][cin][cout] ===============================
        if ((int)blockIdx.x < p.tiles) {
            mbar_wait(done, 0, err, 43);
            tc_fence_after();
            const int co = lane < 16 ? warp * 16 + lane : p.Cout;            // M = 64 layout: row m in lane (m % 16) + 32 * (m / 16)
            for (int c0 = 0; c0 < cin; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
                for (int a = 1; a < nacc; ++a) {
                    float u[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + a * acc_stride + c0, u);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += u[j];
                }
                if (co < p.Cout) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < cin) atomicAdd(dw + ((size_t)g * cin + c0 + j) * p.Cout + co, v[j]);
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

template <int NCO>
int launch_wgrad_tc(const WgP& p, const void* x0, const void* x1, const void* dy, float* dw, int* err, cudaStream_t st) {
    const size_t smem = (size_t)kWgSlotsY * kWgYSlotBytes + (size_t)kWgSlotsX * p.slab_e * 16 + (2 * kWgSlotsX + 2 * kWgSlotsY + 1) * 8 + 16;
    auto kern = conv3_wgrad_tc_kernel<NCO>;
    if (smem > 227 * 1024) { pb_set_error("conv3d_wgrad_tc: needs %zu B of shared memory", smem); return PB_EUNSUPPORTED; }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("conv3d_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    const int items = p.npg * p.QT * p.ND;
    int ctas = 148 / (p.groups * p.nchunks);
    if (ctas < 1) ctas = 1;
    if (ctas > items) ctas = items;
    kern<<<dim3(ctas, p.groups * p.nchunks), kThreads, smem, st>>>(p, (const bf16*)x0, (const bf16*)x1, (const bf16*)dy, dw, err);
    return 0;
}

}  // namespace

// Weight image layout expected by the kernel: [groups][cout tiles][27 taps][NCH chunks][NT rows][8 channels] bf16,
// NCH = max(2, (c0+c1)/8), NT = pb_conv3d_tc_ntile(cout); rows beyond Cout and the padding chunk are zero.
extern "C" int pb_conv3d_tc_ntile(int cin, int cout) {
    if (cin % 8 || cout % 8 || cin < 8 || cout < 8) return 0;
    if (cin > 64 && cin != 128) return 0;        // 8-channel chunk counts the kernel is built for: 1, 2, 4, 8, 16
    if (cin != 8 && cin != 16 && cin != 32 && cin != 64 && cin != 128) return 0;
    if (cout >= 32 && cin <= 32) return 32;      // cin >= 64 keeps NT = 16: weights + plane ring must fit 227 KB
    return 16;
}

// 1 = the (cin, cout) class runs on the kw-stacked kernel (conv3_tc_kws_kernel) and its weight image uses that kernel's
// layout [G][tile][3 kh][chunk][rows: kd = 2,1,0 | kw | 16 co][8].  Measured (profiles/r02_conv_tc_kws.txt): faster than
// conv3_tc_kernel for Cout = 8 (one 8-channel epilogue pass per plane: -16..19 % per launch), slower for Cout = 16 (the
// shifted sum over 16 channels makes the four epilogue warps the bottleneck), so only Cout = 8 is routed here by default.
// PB_TC_KWS=0 switches the variant off, PB_TC_KWS=2 also routes Cout = 16 (A/B measurements).
extern "C" int pb_conv3d_tc_kws(int cin, int cout) {
    static const int mode = [] { const char* e = getenv("PB_TC_KWS"); return e != nullptr && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 1; }();
    if (mode == 0 || pb_conv3d_tc_ntile(cin, cout) != 16) return 0;
    return cout == 8 || (mode == 2 && cout == 16) ? 1 : 0;
}

namespace {
int tc_entry(const pb_conv_desc* d, const void* x0, const void* x1, const void* wimg, const float* bias, void* y0, void* y1,
             void* yext, int inset, int co0, int co1, double* stats, int* err_flag, pb_stream_t stream);
}

namespace {
int tc_tma_mode() {          // 0: cp.async producers everywhere, 1 (default): TMA for the zero-padded launches
    static const int mode = [] { const char* e = getenv("PB_TC_TMA"); return e != nullptr && e[0] == '0' ? 0 : 1; }();
    return mode;
}

// dense NDHWC bf16 volume [planes][H][W][C] seen as (8 ch of a chunk, W, H, C/8 chunks, planes); box = [rows][PW][8 ch] of one chunk
int make_plane_map(CUtensorMap* map, const void* base, int C, int W, int H, long long planes, int box_w, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr) return -1;
    const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)planes};
    const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[5] = {8, (cuuint32_t)box_w, (cuuint32_t)box_rows, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}
}  // namespace

// developer probe: reads and clears the kw-stacked kernel's cycle counters (launches made with PB_TC_PROBE=1)
extern "C" int pb_conv3d_tc_debug(unsigned long long* out8) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out8, tc_dbg, sizeof(z)) != cudaSuccess) return PB_ECUDA;
    if (cudaMemcpyToSymbol(tc_dbg, z, sizeof(z)) != cudaSuccess) return PB_ECUDA;
    return PB_OK;
}

extern "C" int pb_conv3d_tc(const pb_conv_desc* d, const void* x0, const void* x1, const void* wimg, const float* bias,
                            void* y0, void* y1, int co0, int co1, double* stats, int* err_flag, pb_stream_t stream) {
    PB_CHECK_ARG(d && x0 && wimg && y0 && err_flag, "null pointer");
    PB_CHECK_ARG(d->di == d->dout && d->hi == d->ho && d->wi == d->wo, "same-size output only");
    return tc_entry(d, x0, x1, wimg, bias, y0, y1, nullptr, 0, co0, co1, stats, err_flag, stream);
}

extern "C" int pb_conv3d_tc_full(const pb_conv_desc* d, const void* x, const void* wimg, void* y0, void* y1, int co0, int co1,
                                 void* yext, int* err_flag, pb_stream_t stream) {
    PB_CHECK_ARG(d && x && wimg && y0 && yext && err_flag, "null pointer");
    PB_CHECK_ARG(d->dout == d->di + 2 && d->ho == d->hi + 2 && d->wo == d->wi + 2, "output must be the input grown by one voxel per side");
    PB_CHECK_ARG(d->pad_mode == PB_PAD_ZERO && d->c1 == 0, "zero padding, single source");
    PB_CHECK_ARG(d->di >= 4 && d->hi >= 4 && d->wi >= 4, "sizes >= 4");
    return tc_entry(d, x, nullptr, wimg, nullptr, y0, y1, yext, 1, co0, co1, nullptr, err_flag, stream);
}

namespace {

// dx[i] = sum of the extended-domain values that reflect onto i: per axis {i+1} U {0 if i == 1} U {size+1 if i == size-2}
__device__ __forceinline__ int fold_sources(int i, int size, int* src) {
    int n = 0;
    src[n++] = i + 1;
    if (i == 1) src[n++] = 0;
    if (i == size - 2) src[n++] = size + 1;
    return n;
}

// I = index type of the shell enumeration: 32-bit whenever the launch allows it (the enumeration is five divisions per thread; as
// 64-bit divisions they were most of this kernel's time: 71 us for the 600 K shell vectors of a c16 80^3 n8 tensor)
template <typename I>
__global__ void reflect_fold_kernel(const bf16* __restrict__ ext, bf16* __restrict__ y0, bf16* __restrict__ y1, int N, int D, int H,
                                    int W, int CO0, int CO1) {
    const int C = CO0 + CO1, C8 = C >> 3;
    // Threads enumerate exactly the voxels one step inside a face (sizes >= 4): (A) d in {1, D-2}; (B) d elsewhere,
    // h in {1, H-2}; (C) d, h elsewhere, w in {1, W-2}.
    const I HW = (I)H * W;
    const I nA = 2 * HW, nB = (I)(D - 2) * 2 * W, nC = (I)(D - 2) * (H - 2) * 2;
    const I per_n = nA + nB + nC;
    const I total = (I)N * per_n * C8;
    auto inner = [](int j, int S) { return j == 0 ? 0 : (j == S - 3 ? S - 1 : j + 1); };      // j-th index not in {1, S-2}
    for (I t = (I)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (I)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % C8);
        const I v = t / C8;
        const int n = (int)(v / per_n);
        const I iv = v - (I)n * per_n;
        int d, h, w;
        if (iv < nA) {
            d = iv < HW ? 1 : D - 2;
            const int r = (int)(iv < HW ? iv : iv - HW);
            h = r / W; w = r - h * W;
        } else if (iv < nA + nB) {
            const int r = (int)(iv - nA);
            const int q = r / (2 * W);
            d = inner(q, D);
            const int r2 = r - q * 2 * W;
            h = r2 < W ? 1 : H - 2; w = r2 < W ? r2 : r2 - W;
        } else {
            const int r = (int)(iv - nA - nB);
            const int q = r / (2 * (H - 2));
            d = inner(q, D);
            const int r2 = r - q * 2 * (H - 2);
            h = inner(r2 >> 1, H); w = (r2 & 1) ? W - 2 : 1;
        }
        int sd[3], sh[3], sw[3];
        const int nd = fold_sources(d, D, sd), nh = fold_sources(h, H, sh), nw = fold_sources(w, W, sw);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int a = 0; a < nd; ++a)
            for (int b = 0; b < nh; ++b)
                for (int c = 0; c < nw; ++c) {
                    float x[8];
                    VecIO<bf16, 8>::load(ext + ((((size_t)n * (D + 2) + sd[a]) * (H + 2) + sh[b]) * (W + 2) + sw[c]) * C + c8 * 8, x);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] += x[i];
                }
        const size_t vox = (((size_t)n * D + d) * H + h) * W + w;
        const int cb = c8 * 8;
        bf16* dst = cb < CO0 ? y0 + vox * CO0 + cb : y1 + vox * CO1 + (cb - CO0);
        VecIO<bf16, 8>::store(dst, acc);
    }
}

}  // namespace

extern "C" int pb_reflect_fold(const void* yext, void* y0, void* y1, int n, int d, int h, int w, int co0, int co1,
                               pb_stream_t stream) {
    PB_CHECK_ARG(yext && y0 && (co1 == 0 || y1), "null pointer");
    PB_CHECK_ARG(co0 % 8 == 0 && co1 % 8 == 0 && co0 > 0 && d >= 4 && h >= 4 && w >= 4, "bad shape");
    const long long shell = 2LL * h * w + (long long)(d - 2) * 2 * w + (long long)(d - 2) * (h - 2) * 2;
    const long long total = (long long)n * shell * ((co0 + co1) / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    if (total + 148LL * 32 * 256 < 0x7fffffffLL)
        reflect_fold_kernel<int><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)yext, (bf16*)y0, (bf16*)y1, n, d, h, w, co0, co1);
    else
        reflect_fold_kernel<long long><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)yext, (bf16*)y0, (bf16*)y1, n, d, h, w, co0, co1);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

namespace {
int tc_entry(const pb_conv_desc* d, const void* x0, const void* x1, const void* wimg, const float* bias, void* y0, void* y1,
             void* yext, int inset, int co0, int co1, double* stats, int* err_flag, pb_stream_t stream) {
    PB_CHECK_ARG(d->dtype == PB_BF16 && d->ksize == 3 && d->stride == 1, "bf16, 3x3x3, stride 1 only");
    const int cin = d->c0 + d->c1, cout = co0 + co1;
    PB_CHECK_ARG(cout == d->cout && co0 % 8 == 0 && co1 % 8 == 0 && (co1 == 0 || y1), "bad output split");
    PB_CHECK_ARG(d->c0 % 8 == 0 && d->c1 % 8 == 0 && (d->c1 == 0 || x1), "channels must be multiples of 8");
    const int NT = pb_conv3d_tc_ntile(cin, cout);
    PB_CHECK_ARG(NT != 0, "unsupported channel class");
    PB_CHECK_ARG(d->groups >= 1 && d->n % d->groups == 0, "bad groups");
    TcP p;
    p.N = d->n; p.D = d->dout; p.H = d->ho; p.W = d->wo; p.C0 = d->c0; p.C1 = d->c1; p.CO0 = co0; p.CO1 = co1;
    p.reflect = d->pad_mode == PB_PAD_REFLECT;
    p.inset = inset;
    PB_CHECK_ARG(!p.reflect || (p.D >= 2 && p.H >= 2 && p.W >= 2), "reflect padding needs size >= 2");
    p.PW = p.W + 2;
    const bool kws = pb_conv3d_tc_kws(cin, cout) != 0;
    p.QT = kws ? (p.H * p.PW + kKwsStride - 1) / kKwsStride : (p.H * p.PW + kTileM - 1) / kTileM;
    p.npg = d->n / d->groups; p.groups = d->groups;
    p.nt_tiles = (cout + NT - 1) / NT;
    // depth chunking: enough work items to balance ~2 waves of CTAs, chunks of at least 8 planes
    const int target = 148 * 4 / (p.groups * p.nt_tiles);
    int nd = 1;
    while (p.npg * p.QT * nd < target && (p.D + nd) / (nd + 1) >= 8) ++nd;
    p.DCH = (p.D + nd - 1) / nd;
    p.ND = (p.D + p.DCH - 1) / p.DCH;
    p.slab_need = kws ? kTileM + 2 * p.PW : kTileM + 2 * p.PW + 2;
    p.q_stride = kws ? kKwsStride : kTileM;
    const int nchr = cin / 8, nch = nchr < 2 ? 2 : nchr;
    // pad the plane pitch so that the nch chunk planes start in different shared-memory banks
    const int want = nch >= 8 ? 1 : 8 / nch;
    int se = p.slab_need;
    while (se % 8 != want % 8) ++se;
    p.slab_e = se;
    // TMA-fed input planes for zero-padded launches — every data gradient (PB_TC_TMA=0: cp.async producers everywhere).
    // Zero padding is the tensor map's out-of-bounds fill; reflect-padded launches keep the cp.async producers (see tma_producer).
    CUtensorMap tmap, tmap1;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap1, 0, sizeof(tmap1));
    p.tma = 0; p.tma_rows = 0;
    if (tc_tma_mode() >= 1 && !p.reflect && d->c0 % 8 == 0 && d->c1 % 8 == 0) {
        const int rows = (p.slab_need + p.PW - 2) / p.PW + 1;               // padded rows a slab of slab_need positions can touch
        const int se_t = ((rows * p.PW + 7) / 8) * 8;                       // chunk-plane pitch: TMA destinations are 128 B aligned
        const int Di = p.D - 2 * inset, Hi = p.H - 2 * inset, Wi = p.W - 2 * inset;
        if (p.PW <= 256 && rows <= 256 && make_plane_map(&tmap, x0, d->c0, Wi, Hi, (long long)p.N * Di, p.PW, rows) == 0 &&
            (d->c1 == 0 || make_plane_map(&tmap1, x1, d->c1, Wi, Hi, (long long)p.N * Di, p.PW, rows) == 0)) {
            p.tma = 1; p.tma_rows = rows; p.slab_e = se_t;
        }
    }
    p.w_tile_bytes = (long long)27 * nch * NT * 16;
    { const char* e = getenv("PB_TC_PROBE"); p.probe = e ? atoi(e) : 0; }
    if (p.slab_need * nchr > max_copies(nchr) * kTcProducers) {
        pb_set_error("conv3d_tc: plane slab of %d x %d copies exceeds the producer budget", p.slab_need, nchr);
        return PB_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = PB_EUNSUPPORTED;
#define KWS_CASE(NCHR_, CR_) if (kws && nchr == NCHR_ && cout == 8 * CR_) rc = launch_tc_kws<NCHR_, CR_>(tmap, tmap1, p, x0, x1, wimg, bias, y0, y1, yext, stats, err_flag, st)
    KWS_CASE(1, 1); KWS_CASE(2, 1); KWS_CASE(4, 1); KWS_CASE(8, 1); KWS_CASE(16, 1);
    KWS_CASE(1, 2); KWS_CASE(2, 2); KWS_CASE(4, 2); KWS_CASE(8, 2); KWS_CASE(16, 2);
#undef KWS_CASE
#define TC_CASE(NCHR_, NT_) if (!kws && nchr == NCHR_ && NT == NT_) rc = launch_tc<NCHR_, NT_>(tmap, tmap1, p, x0, x1, wimg, bias, y0, y1, yext, stats, err_flag, st)
    TC_CASE(1, 16); TC_CASE(2, 16); TC_CASE(4, 16); TC_CASE(8, 16); TC_CASE(16, 16);
    TC_CASE(1, 32); TC_CASE(2, 32); TC_CASE(4, 32); TC_CASE(8, 32);
#undef TC_CASE
    if (rc) { if (rc == PB_EUNSUPPORTED) pb_set_error("conv3d_tc: no kernel for cin %d cout %d", cin, cout); return rc; }
    PB_CHECK_LAUNCH();
    return PB_OK;
}
}  // namespace

extern "C" int pb_conv3d_wgrad_tc(const pb_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw, int* err_flag,
                                  pb_stream_t stream) {
    PB_CHECK_ARG(d && x0 && dy && dw && err_flag, "null pointer");
    PB_CHECK_ARG(d->dtype == PB_BF16 && d->ksize == 3 && d->stride == 1, "bf16, 3x3x3, stride 1 only");
    PB_CHECK_ARG(d->di == d->dout && d->hi == d->ho && d->wi == d->wo, "same-size output only");
    PB_CHECK_ARG(d->c0 % 8 == 0 && d->c1 % 8 == 0 && d->c0 >= 8 && (d->c1 == 0 || x1), "input channels must be multiples of 8");
    PB_CHECK_ARG(d->cout % 8 == 0 && d->cout >= 8 && d->cout <= 64, "cout must be a multiple of 8 in [8, 64]");
    PB_CHECK_ARG(d->groups >= 1 && d->n % d->groups == 0, "bad groups");
    WgP p;
    p.N = d->n; p.D = d->di; p.H = d->hi; p.W = d->wi; p.C0 = d->c0; p.C1 = d->c1; p.Cout = d->cout;
    p.reflect = d->pad_mode == PB_PAD_REFLECT;
    PB_CHECK_ARG(!p.reflect || (p.D >= 2 && p.H >= 2 && p.W >= 2), "reflect padding needs size >= 2");
    {   // row-stacked kernel (conv3d_wgrad_rs.cu): one MMA per 16 voxels and input chunk for all 27 taps
        const int rc = pb_wgrad_rs_launch(d, x0, x1, dy, dw, err_flag, (cudaStream_t)stream);
        if (rc == 0) { PB_CHECK_LAUNCH(); return PB_OK; }
        if (rc != PB_EUNSUPPORTED) return rc;
    }
    p.PW = p.W + 2;
    p.QT = (p.H * p.PW + kTileM - 1) / kTileM;
    p.npg = d->n / d->groups; p.groups = d->groups;
    p.nchunks = (d->c0 + d->c1) / 8;
    const int target = 148 * 2 / (p.groups * p.nchunks) + 1;
    int nd = 1;
    while (p.npg * p.QT * nd < target && (p.D + nd) / (nd + 1) >= 8) ++nd;
    p.DCH = (p.D + nd - 1) / nd;
    p.ND = (p.D + p.DCH - 1) / p.DCH;
    p.slab_need = kTileM + 2 * p.PW + 3;
    p.slab_e = (p.slab_need + 7) & ~7;
    if (p.slab_need > kWgCopies * kProducerThreads) { pb_set_error("conv3d_wgrad_tc: plane slab of %d rows exceeds the producer budget", p.slab_need); return PB_EUNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = PB_EUNSUPPORTED;
    switch (d->cout / 8) {
        case 1: rc = launch_wgrad_tc8<1>(p, x0, x1, dy, dw, err_flag, st); break;
        case 2: rc = kWgStackCout16 ? launch_wgrad_tc8<2>(p, x0, x1, dy, dw, err_flag, st) : launch_wgrad_tc<2>(p, x0, x1, dy, dw, err_flag, st); break;
        case 4: rc = kWgStackCout32 ? launch_wgrad_tc8<4>(p, x0, x1, dy, dw, err_flag, st) : launch_wgrad_tc<4>(p, x0, x1, dy, dw, err_flag, st); break;
        case 8: rc = launch_wgrad_tc<8>(p, x0, x1, dy, dw, err_flag, st); break;
        default: pb_set_error("conv3d_wgrad_tc: cout %d not supported", d->cout); break;
    }
    if (rc) return rc;
    PB_CHECK_LAUNCH();
    return PB_OK;
}

extern "C" int pb_conv1_wgrad_tc(const pb_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw, int* err_flag,
                                 pb_stream_t stream) {
    PB_CHECK_ARG(d && x0 && dy && dw && err_flag, "null pointer");
    PB_CHECK_ARG(d->dtype == PB_BF16 && d->ksize == 1 && d->stride == 1, "bf16, 1x1x1, stride 1 only");
    PB_CHECK_ARG(d->c0 % 8 == 0 && d->c1 % 8 == 0 && d->c0 >= 8 && (d->c1 == 0 || x1), "input channels must be multiples of 8");
    PB_CHECK_ARG(d->cout % 8 == 0 && d->cout >= 8 && d->cout <= 64, "cout must be a multiple of 8 in [8, 64]");
    PB_CHECK_ARG(d->groups >= 1 && d->n % d->groups == 0, "bad groups");
    const int cin = d->c0 + d->c1;
    if (cin > 256) { pb_set_error("conv1_wgrad_tc: cin %d > 256", cin); return PB_EUNSUPPORTED; }
    W1P p;
    p.C0 = d->c0; p.C1 = d->c1; p.Cout = d->cout;
    p.VT = (long long)(d->n / d->groups) * d->dout * d->ho * d->wo;
    const long long tiles = (p.VT + kTileM - 1) / kTileM;
    if (tiles > 0x7fffffffLL) { pb_set_error("conv1_wgrad_tc: too many tiles"); return PB_EUNSUPPORTED; }
    p.tiles = (int)tiles;
    const int slot_bytes = (cin / 8 + d->cout / 8) * kW1Plane;
    const int table_bytes = kTileM * (cin / 8 + d->cout / 8) * 4;
    // two CTAs per SM when the accumulators (2 x cin columns each) and two rings fit, else one
    const bool two = cin <= 128 && 2 * slot_bytes + 8 * kW1Plane + table_bytes <= 100 * 1024;
    int slots = ((two ? 100 : 200) * 1024 - 8 * kW1Plane - table_bytes) / slot_bytes;
    if (slots > 8) slots = 8;
    if (slots < 2) { pb_set_error("conv1_wgrad_tc: a tile of %d B does not leave room for two slots", slot_bytes); return PB_EUNSUPPORTED; }
    p.slots = slots;
    const size_t smem = (size_t)slots * slot_bytes + 8 * kW1Plane + (2 * slots + 1) * 8 + 16 + table_bytes;
    cudaError_t e = cudaFuncSetAttribute(conv1_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("conv1_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    int ctas = (two ? 296 : 148) / d->groups;
    if (ctas < 1) ctas = 1;
    if (ctas > p.tiles) ctas = p.tiles;
    conv1_wgrad_tc_kernel<<<dim3(ctas, d->groups), kThreads, smem, (cudaStream_t)stream>>>(p, (const bf16*)x0, (const bf16*)x1,
                                                                                         (const bf16*)dy, dw, err_flag);
    PB_CHECK_LAUNCH();
    return PB_OK;
}

