// Row-stacked tcgen05 weight gradient of the 3x3x3 stride-1 convs (sm_100a) — round 2 replacement of the kh-stacked kernels of
// conv3d_tc.cu.  Replaces the weight-gradient half of nn.Conv3d's autograd node (reference models/blocks.py:357, general_conv3d).
//
//   dw[kd,kh,kw][ci][co] = sum_{n,d,h,w} xp[n, d+kd-1, h+kh-1, w+kw-1][ci] * dy[n,d,h,w][co]        (xp = reflect / zero padded x)
//
// Why a new shape: the old kernels (M = 64, N = 32: six tcgen05.mma per 16 voxels of a c16 -> 8 layer) were bound by the
// instruction count — an MMA with both operands MN-major costs the pipe max(36, M*N*16 / 2048 [M = 64] or 4096 [M = 128]) cycles
// and ONE thread cannot issue them faster than one per 61.5 cycles (scripts/microbench/umma_rate_mn*.cu, profiles/r02_wgrad_rs.txt).
// Here ONE instruction per 16 voxels, UNIT of input channels and 8-channel output chunk produces all 27 taps:
//   * K = 16 consecutive w positions of ONE padded input row (d, h); a CTA walks whole rows.
//   * A (MN-major) = that row of xp for one unit of 8 / 16 / 32 input channels; its M-groups are the same row shifted by 0, 1, 2, ...
//     positions (one position per group): groups 0..2 are the three kw taps, the others are don't-care rows that are never read back.
//     8 channels = plain 16-byte rows (M = 64); 16 / 32 channels = the tensor's own 32- / 64-byte NDHWC rows under SWIZZLE_32B (M = 64) /
//     SWIZZLE_64B (M = 128) — the producers store each 16-byte chunk at its swizzled position (rs_adesc below).
//   * B (N = 88, MN-major) = the dy rows (d-1..d+1) x (h-1..h+1) that pair with this input row, one 8-channel output chunk.
//     The dy strip lives in shared memory in a DIAGONAL layout: row h' of the q-th plane of the sweep sits at row index
//     L = 4 h' + q of one linear buffer.  The nine rows of an input row are then L0 + {0,1,2, 4,5,6, 8,9,10} with L0 moving by one
//     per plane: ONE descriptor with SBO = row pitch and 11 N-groups addresses them, always in the same order (no rotation of
//     the accumulator columns); groups 3 and 7 belong to plane q + 3, which is being prefetched while plane q's step runs —
//     their accumulator columns are garbage and are never read back.  Plane q + 4 overwrites plane q one row further up.
// A CTA (one per SM) owns (weight group, up to four input chunks, one output chunk) and a stream of work items (sample, strip of
// <= 16 input rows, depth chunk of <= 16 planes) which it takes in interleaved PAIRS (steps A.0, B.0, A.1, B.1, ...; each item of a pair has
// its own dy buffer region and plane barriers), so that three steps of other work lie between the retirement of a dy plane and the step
// that needs the plane loaded in its place.  The accumulators (96 TMEM columns per unit) stay in TMEM for the whole stream and are
// flushed once with fp32 atomics.
// Warp roles (13 warps): warps 0-7 stage the input rows (16-byte cp.async; reflect / zero padding and the two-source concat are
// resolved in the address computation, the K tail is zero-filled), running up to three steps ahead; warp 8 streams the dy rows, one row
// per lane: plain bulk copies where a dy row is contiguous (Cout = 8, W a multiple of 16; rows outside the volume zero-filled by the
// warp), else one TMA box [KW positions][8 ch] per row (rows, planes and the K tail outside the volume = the tensor map's zero fill);
// lanes 0 of warps 9-12 issue the MMAs, all accumulating into the same zero-initialised accumulators, with a division-free issue loop
// whose K loop is unrolled (one thread needs ~65-75 cycles per instruction; 119 with a rolled loop, 185 with a division per row).
// Warps 0-3 flush at the end.  PB_WG_RS=0 falls back to the kh-stacked kernels, =2 / 3 / 5 are developer probes (no MMAs / no input
// copies / issuer cycle counters, scripts/probe_wgrad_issuer.py); PB_WG_RS_RU > 1 puts several input rows side by side in one instruction
// (built, measured slower, off).
#include <cstdlib>
#include "tc_common.cuh"

namespace {

constexpr int kRsThreads = 416;            // warps 0-7 input rows, 8 dy rows, 9-12 MMA issuers
constexpr int kRsXProducers = 256;         // warps 0-7
constexpr int kRsXSlots = 3;
constexpr int kRsXCopies = 8;              // 16-byte copies per x-producer thread and step
constexpr int kRsMaxRows = 16;
constexpr int kRsMaxSteps = 16;            // planes per work item (bounds the diagonal dy buffer)
constexpr int kRsMaxChunks = 4;            // input chunks per CTA
constexpr int kRsIssuers = 4;

struct RsP {
    int N, D, H, W, C0, C1, Cout, reflect;
    int KW;                                // K extent of a row: W rounded up to 16
    int nr;                                // input rows per strip
    int nstrips, ND, npg, groups, nT, nP, TG, ntg;   // TG input chunks per CTA, ntg chunk groups
    int h_lo, rows_t, d_lo, planes_t;      // padded input rows / planes that contribute: [-1, H] for reflect, [0, H) for zero padding
    int yrows, nregions;                   // rows of one diagonal dy buffer region; regions (2: one per item of a pair)
    int dbg;                               // developer probes: 2 = no MMAs, 3 = no input-row copies (results are garbage)
    uint32_t tmem_cols;
    int UC;                                // 16-byte sub-units per position of a unit = CU chunks x RU rows: 1 plain rows, 2 SWIZZLE_32B, 4 SWIZZLE_64B
    int CU, RU;                            // input chunks and input ROWS that share one MMA (the RU rows of a group sit side by side like channels)
    int NN, setw;                          // UMMA N = 8 * (11 + 4 * (RU - 1)); TMEM columns per unit
    int prefetch;                          // dy planes ahead to prefetch into L2 (0 = off)
    int ybulk;                             // 1: dy rows are contiguous (Cout = 8, W % 16 == 0) and fetched by plain bulk copies
};

__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// Producer-side wait: ONE lane polls, with a back-off, and the warp follows.  The MMAs of this kernel read ~108 B of shared memory
// per clock (4.75 KB per 44-cycle instruction) out of the SM's 128: nine warps spinning on mbarrier.try_wait took a third of it.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane, volatile int* err, int code) {
    if (lane == 0 && !mbar_try_wait(bar, parity)) {
        const long long t0 = clock64();
        uint32_t spins = 0;
        while (!mbar_try_wait(bar, parity)) {
            __nanosleep(40);
            if ((++spins & 255u) == 0) {
                if (*err != 0) break;
                if (clock64() - t0 > 4000000000LL) { atomicCAS((int*)err, 0, code); break; }
            }
        }
    }
    __syncwarp();
}

__device__ unsigned long long rs_dbg[8];            // developer probe (PB_WG_RS=5): cycles of issuer 0 spent waiting / issuing, summed over CTAs

struct RsItem { int n, hs, nr, dx0, nsteps; };       // sample, first padded input row, rows, first padded plane, planes (0 = no such item)

__device__ __forceinline__ RsItem rs_item(const RsP& p, int it, int items, int g) {
    RsItem r;
    if (it >= items) { r.n = 0; r.hs = 0; r.nr = 0; r.dx0 = 0; r.nsteps = 0; return r; }
    const int dc = it % p.ND, r1 = it / p.ND;
    const int st = r1 % p.nstrips;
    r.n = g * p.npg + r1 / p.nstrips;
    r.hs = p.h_lo + (st * p.rows_t) / p.nstrips;
    r.nr = p.h_lo + ((st + 1) * p.rows_t) / p.nstrips - r.hs;
    r.dx0 = p.d_lo + (dc * p.planes_t) / p.ND;
    r.nsteps = p.d_lo + ((dc + 1) * p.planes_t) / p.ND - r.dx0;
    return r;
}

// A operand of one input row for a unit of UC chunks, MN-major (cute canonical layouts, T = 8 elements = 16 B):
//   UC = 1: no swizzle   ((T,1,m),(8,k)) : ((1,T,SBO),(1T,LBO))   m = shift groups at SBO = 16 B (one position), 8 k at LBO = 128 B
//   UC = 2: SWIZZLE_32B  ((T,2,m),(8,k)) : ((1,T,LBO),(2T,SBO))   16-channel rows of 32 B, m = shift groups at LBO = 32 B, 8 k at SBO = 256 B
//   UC = 4: SWIZZLE_64B  ((T,4,m),(8,k)) : ((1,T,LBO),(4T,SBO))   32-channel rows of 64 B, LBO = 64 B, SBO = 512 B
// The swizzle XORs the 16-byte unit index within a row with address bits 7.. (one bit / two bits): the producers store a position's
// chunks permuted accordingly; row starts are multiples of 256 / 512 B so the pattern is a function of the position alone and one
// staged row serves all shifts.
template <int UC>
__device__ __forceinline__ uint64_t rs_adesc(uint32_t addr) {
    if constexpr (UC == 1) return umma_desc(addr, 128, 16);
    else if constexpr (UC == 2) return umma_desc(addr, 32, 256) | (6ULL << 61);
    else return umma_desc(addr, 64, 512) | (4ULL << 61);
}

template <int UC, int NN, int KS>
__device__ __forceinline__ void rs_issue(uint32_t d_tmem, uint32_t xrow, uint32_t yrow, int ypitch_b) {
    constexpr uint32_t IDESC = umma_idesc(UC == 4 ? 128 : 64, NN) | (1u << 15) | (1u << 16);      // both operands MN-major
    const uint64_t a0 = rs_adesc<UC>(xrow);
    const uint64_t b0 = umma_desc(yrow, 128, (uint32_t)ypitch_b);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) umma_f16(d_tmem, a0 + (uint64_t)(16 * UC * ks), b0 + (uint64_t)(16 * ks), IDESC, 1u);
}

// All MMAs one issuer contributes to a step: every kRsIssuers-th (unit, row) pair, KS instructions each.  Templated on the unit
// width and the K-step count and free of divisions: the issuing threads are the bottleneck of this kernel (a thread needs ~65-75
// cycles per instruction even with the descriptors in registers; with a division, a switch and the descriptor arithmetic per pair
// ncu / clock64 probes showed 185).
template <int UC, int NN, int KS>
__device__ __forceinline__ void rs_issue_step(int issuer, int nunits, int ngr, int ngr_pitch, int ru, int setw, uint32_t tmem_base, uint32_t xs0,
                                              uint32_t upitch, uint32_t yrow0, int ypitch_b) {
    int u = 0, i = issuer;                                  // i = row GROUP (ru input rows per instruction)
    while (i >= ngr) { i -= ngr; ++u; }
    while (u < nunits) {
        rs_issue<UC, NN, KS>(tmem_base + u * setw, xs0 + (u * ngr_pitch + i) * upitch, yrow0 + 4 * i * ru * ypitch_b, ypitch_b);
        i += kRsIssuers;
        while (i >= ngr) { i -= ngr; ++u; }
    }
}

template <int UC, int NN>
__device__ __forceinline__ void rs_issue_step_k(int ksteps, int issuer, int nunits, int ngr, int ngr_pitch, int ru, int setw, uint32_t tmem_base,
                                                uint32_t xs0, uint32_t upitch, uint32_t yrow0, int ypitch_b) {
    switch (ksteps) {
        case 1: rs_issue_step<UC, NN, 1>(issuer, nunits, ngr, ngr_pitch, ru, setw, tmem_base, xs0, upitch, yrow0, ypitch_b); break;
        case 2: rs_issue_step<UC, NN, 2>(issuer, nunits, ngr, ngr_pitch, ru, setw, tmem_base, xs0, upitch, yrow0, ypitch_b); break;
        case 3: rs_issue_step<UC, NN, 3>(issuer, nunits, ngr, ngr_pitch, ru, setw, tmem_base, xs0, upitch, yrow0, ypitch_b); break;
        case 4: rs_issue_step<UC, NN, 4>(issuer, nunits, ngr, ngr_pitch, ru, setw, tmem_base, xs0, upitch, yrow0, ypitch_b); break;
        case 5: rs_issue_step<UC, NN, 5>(issuer, nunits, ngr, ngr_pitch, ru, setw, tmem_base, xs0, upitch, yrow0, ypitch_b); break;
        case 6: rs_issue_step<UC, NN, 6>(issuer, nunits, ngr, ngr_pitch, ru, setw, tmem_base, xs0, upitch, yrow0, ypitch_b); break;
        case 7: rs_issue_step<UC, NN, 7>(issuer, nunits, ngr, ngr_pitch, ru, setw, tmem_base, xs0, upitch, yrow0, ypitch_b); break;
        default: rs_issue_step<UC, NN, 8>(issuer, nunits, ngr, ngr_pitch, ru, setw, tmem_base, xs0, upitch, yrow0, ypitch_b); break;
    }
}

__global__ void __launch_bounds__(kRsThreads, 1)
conv3_wgrad_rs_kernel(const __grid_constant__ CUtensorMap ymap, RsP p, const bf16* __restrict__ x0, const bf16* __restrict__ x1,
                      const bf16* __restrict__ dy, float* __restrict__ dw, int* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ypitch_b = p.KW * 16;
    const int xpitch_b = (p.KW + 8) * 16;
    const int yregion_bytes = p.yrows * ypitch_b;
    const int ybuf_bytes = p.nregions * yregion_bytes;
    const int xslot_bytes = p.TG * p.nr * xpitch_b;
    // input-row slots first, on a 1024-byte boundary of the shared address space (the swizzle patterns are functions of address
    // bits 4..8); slot, unit and row strides are multiples of 512 B
    uint8_t* x_s = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    uint8_t* y_s = x_s + (size_t)kRsXSlots * xslot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(y_s + ybuf_bytes);
    uint64_t* fullx = bars;
    uint64_t* emptyx = bars + kRsXSlots;
    uint64_t* fully = bars + 2 * kRsXSlots;
    uint64_t* emptyy = fully + 8;                       // fully / emptyy: [2 items of a pair][4 plane slots]
    uint64_t* item_done = emptyy + 8;                  // [2]: one per dy buffer region
    uint64_t* done = item_done + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int by = blockIdx.y;
    const int cochunk = by % p.nP, tg = (by / p.nP) % p.ntg, g = by / (p.nP * p.ntg);
    const int ch0 = tg * p.TG;
    const int nch = min(p.TG, p.nT - ch0);                    // input chunks of this CTA
    const int items = p.npg * p.nstrips * p.ND;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRsXSlots; ++i) { mbar_init(&fullx[i], kRsXProducers); mbar_init(&emptyx[i], kRsIssuers); }
        for (int i = 0; i < 8; ++i) { mbar_init(&fully[i], 1); mbar_init(&emptyy[i], kRsIssuers); }
        mbar_init(&item_done[0], kRsIssuers);
        mbar_init(&item_done[1], kRsIssuers);
        mbar_init(done, kRsIssuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < 4) {
        // every MMA accumulates (several threads issue into the same accumulators, in no particular order): start from zero
        for (int c = 0; c < (int)p.tmem_cols; c += 16) tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // Work items are taken in PAIRS (items it and it + gridDim.x of this CTA's stream) whose steps are interleaved: A.0, B.0, A.1, B.1, ...
    // Each of the two has its own dy buffer region and plane barriers.  The copy of a plane can only start when the MMAs of the
    // step that last read the plane it overwrites have retired; interleaving puts three steps of other work between that moment
    // and the step that needs the new plane (one step was not enough: ncu showed the issuers spinning on the dy `full` barrier).
    const int pair_stride = 2 * gridDim.x;
    if (warp < 8) {
        // =============================== input-row producers (run ahead of the MMAs by up to kRsXSlots steps) ===============
        const int pt = threadIdx.x;
        const int c0ch = p.C0 >> 3;
        const int xrow_e = p.KW + 8;
        const int per_chunk = p.nr * xrow_e;
        uint32_t kx = 0;
        for (int it0 = blockIdx.x; it0 < items; it0 += pair_stride) {
            int xoff[2][kRsXCopies];                                            // -2: nothing to copy, -1: zero fill, else offset in the plane
            uint32_t xdst[kRsXCopies];                                          // (same for both items: they differ in rows, not in layout)
            uint32_t src1 = 0;                                                  // bit i: copy i reads the second source
            int nsteps[2], dxs[2], ns[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const RsItem im = rs_item(p, it0 + k * (int)gridDim.x, items, g);
                nsteps[k] = im.nsteps; dxs[k] = im.dx0; ns[k] = im.n;
#pragma unroll
                for (int i = 0; i < kRsXCopies; ++i) {
                    const int e = pt + i * kRsXProducers;
                    xoff[k][i] = -2;
                    if (k == 0) xdst[i] = 0;
                    if (e < nch * per_chunk) {
                        const int c = e / per_chunk, rem = e - c * per_chunk;
                        const int r = rem / xrow_e, t = rem - r * xrow_e;
                        const int ch = ch0 + c;
                        const bool from1 = ch >= c0ch;
                        // unit u = c / CU; RU consecutive input rows form a row group whose rows sit side by side like channels:
                        // [unit][row group][position][(row in group, chunk in unit) x 16 B], the 16-byte sub-units swizzled
                        const int u = c / p.CU, cu = c - u * p.CU;
                        const int rg = r / p.RU, rho = r - rg * p.RU;
                        const int sw = p.UC == 1 ? 0 : (p.UC == 2 ? ((t >> 2) & 1) : ((t >> 1) & 3));
                        xdst[i] = (uint32_t)(((u * (p.nr / p.RU) + rg) * xrow_e + t) * 16 * p.UC + (((rho * p.CU + cu) ^ sw) * 16));
                        if (from1) src1 |= 1u << i;
                        if (r >= im.nr && r < ((im.nr + p.RU - 1) / p.RU) * p.RU) xoff[k][i] = -1;      // pad of the last row group: zeros
                        if (r < im.nr) {
                            const int cs = from1 ? p.C1 : p.C0, coff = (from1 ? ch - c0ch : ch) * 8;
                            int h = im.hs + r, w = t - 1;
                            bool ok;
                            if (p.reflect) { ok = w >= -1 && w <= p.W; h = reflect_idx(h, p.H); w = reflect_idx(w, p.W); }
                            else ok = w >= 0 && w < p.W;
                            xoff[k][i] = ok ? (h * p.W + w) * cs + coff : -1;
                        }
                    }
                }
            }
            const int smax = nsteps[0] > nsteps[1] ? nsteps[0] : nsteps[1];
            for (int s = 0; s < smax; ++s) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (s >= nsteps[k]) continue;
                    const int slot = kx % kRsXSlots;
                    mbar_wait_warp(&emptyx[slot], ((kx / kRsXSlots) & 1) ^ 1, lane, err, 42);
                    const int dx = dxs[k] + s;
                    const int dpx = p.reflect ? reflect_idx(dx, p.D) : dx;
                    const size_t pl = ((size_t)ns[k] * p.D + dpx) * p.H * p.W;
                    const bf16* pp0 = x0 + pl * p.C0;
                    const bf16* pp1 = x1 + pl * p.C1;
                    const uint32_t sbase = smem_u32(x_s) + (uint32_t)(slot * xslot_bytes);
#pragma unroll
                    for (int i = 0; i < kRsXCopies; ++i) {
                        if (xoff[k][i] != -2 && p.dbg != 3) {
                            const bool ok = xoff[k][i] >= 0;
                            const bf16* src = ((src1 >> i) & 1u) ? pp1 : pp0;
                            cp_async16(sbase + xdst[i], ok ? src + xoff[k][i] : x0, ok ? 16u : 0u);
                        }
                    }
                    cp_async_arrive_noinc(&fullx[slot]);
                    ++kx;
                }
            }
        }
        cp_async_wait_all();
    } else if (warp == 8) {
        // =============================== dy-row producer (one warp, one row per lane) ===============================
        const uint32_t plane_bytes = (uint32_t)((p.nr + 2) * ypitch_b);
        uint32_t ky[2] = {0, 0}, kpair = 0;
        for (int it0 = blockIdx.x; it0 < items; it0 += pair_stride, ++kpair) {
            RsItem im[2];
            im[0] = rs_item(p, it0, items, g);
            im[1] = rs_item(p, it0 + (int)gridDim.x, items, g);
            // the diagonal buffer of a region restarts at q = 0 with every item: the MMAs of the item that used it last must have drained
            const int smax = im[0].nsteps > im[1].nsteps ? im[0].nsteps : im[1].nsteps;
            // planes dx0-1 .. dx0+nsteps (q = 0 .. nsteps+1) of each item: q = 0, 1, 2 with its step 0, then one per step; plane q
            // may land once plane q-4 is dead.  Always nr + 2 rows (a shorter strip leaves the last rows unused).
            for (int s = 0; s < smax; ++s) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (s >= im[k].nsteps) continue;
                    if (s == 0 && kpair > 0) mbar_wait_warp(&item_done[k], (kpair - 1) & 1, lane, err, 40);
                    for (int q = (s == 0 ? 0 : s + 2); q <= s + 2; ++q, ++ky[k]) {
                        const int dp = im[k].dx0 - 1 + q;
                        const int slot = ky[k] & 3;
                        uint64_t* full_b = &fully[4 * k + slot];
                        mbar_wait_warp(&emptyy[4 * k + slot], ((ky[k] >> 2) & 1) ^ 1, lane, err, 41);
                        const bool plane_ok = dp >= 0 && dp < p.D;
                        const uint32_t ybase = smem_u32(y_s) + (uint32_t)k * (uint32_t)yregion_bytes + (uint32_t)(q * ypitch_b);
                        const int n = im[k].n, hs = im[k].hs;
                        if (p.ybulk) {
                            // Cout = 8 and W a multiple of 16: a dy row is W * 16 contiguous bytes — one bulk copy per row and lane
                            // instead of a tensor box of W 16-byte rows.  Rows outside the volume are zero-filled by the warp itself.
                            const int h = hs - 1 + lane;
                            const bool row_ok = plane_ok && lane < p.nr + 2 && h >= 0 && h < p.H;
                            const unsigned okmask = __ballot_sync(0xffffffffu, row_ok);
                            unsigned zmask = ~okmask & ((p.nr + 2 >= 32) ? 0xffffffffu : ((1u << (p.nr + 2)) - 1u));
                            while (zmask) {
                                const int r = __ffs(zmask) - 1;
                                zmask &= zmask - 1;
                                uint4* dst = reinterpret_cast<uint4*>(y_s + k * (size_t)yregion_bytes + (size_t)(q + 4 * r) * ypitch_b);
                                for (int i = lane; i < p.KW; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
                            }
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) mbar_expect_tx(full_b, (uint32_t)__popc(okmask) * (uint32_t)ypitch_b);
                            __syncwarp();
                            if (row_ok)
                                bulk_g2s(ybase + (uint32_t)(lane * 4 * ypitch_b), dy + (((size_t)n * p.D + dp) * p.H + h) * p.W * 8,
                                         (uint32_t)ypitch_b, full_b);
                        } else {
                            // one TMA box [KW positions][8 ch] per row; rows, planes and the K tail outside the volume are zero fill
                            if (lane == 0) mbar_expect_tx(full_b, plane_bytes);
                            __syncwarp();
                            if (lane < p.nr + 2)
                                tma_load_5d(ybase + (uint32_t)(lane * 4 * ypitch_b), &ymap, 0, 0, plane_ok ? hs - 1 + lane : p.H + 8, cochunk,
                                            plane_ok ? n * p.D + dp : 0, full_b);
                        }
                    }
                }
            }
        }
    } else if (lane == 0) {
        // =============================== MMA issuers (warps 9-12) ===============================
        // One thread cannot issue faster than one MMA per ~62 cycles and the pipe retires one in 44: every issuer walks all steps
        // (it must see every `full` phase before it may arrive on the matching `empty`) and issues every fourth (unit, row) pair.
        const int issuer = warp - 9;
        const uint32_t x_addr = smem_u32(x_s), y_addr = smem_u32(y_s);
        const int ksteps = p.KW >> 4;
        uint32_t kx = 0, ky[2] = {0, 0};
        long long dbg_y0 = 0, dbg_y = 0, dbg_x = 0, dbg_issue = 0, dbg_steps = 0;
        const long long dbg_t0 = clock64();
        for (int it0 = blockIdx.x; it0 < items; it0 += pair_stride) {
            RsItem im[2];
            im[0] = rs_item(p, it0, items, g);
            im[1] = rs_item(p, it0 + (int)gridDim.x, items, g);
            const int smax = im[0].nsteps > im[1].nsteps ? im[0].nsteps : im[1].nsteps;
            for (int s = 0; s < smax; ++s) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (s >= im[k].nsteps) continue;
                    // step s reads planes q = s, s+1, s+2: three new ones at the start of an item, one per step afterwards
                    const int nl = s == 0 ? 3 : 1;
                    const long long t0 = clock64();
                    for (int l = 0; l < nl; ++l, ++ky[k]) mbar_wait(&fully[4 * k + (ky[k] & 3)], (ky[k] >> 2) & 1, err, 43);
                    const long long t1 = clock64();
                    mbar_wait(&fullx[kx % kRsXSlots], (kx / kRsXSlots) & 1, err, 44);
                    const long long t2 = clock64();
                    fence_proxy_async();
                    tc_fence_after();
                    if (s == 0) dbg_y0 += t1 - t0; else dbg_y += t1 - t0;
                    dbg_x += t2 - t1;
                    const uint32_t xs0 = x_addr + (kx % kRsXSlots) * xslot_bytes;
                    const uint32_t yreg = y_addr + (uint32_t)k * (uint32_t)yregion_bytes;
                    const int nr = im[k].nr;
                    const int nunits = nch / p.CU;
                    const int ngr = (nr + p.RU - 1) / p.RU, ngr_pitch = p.nr / p.RU;
                    const uint32_t upitch = (uint32_t)xpitch_b * p.UC;         // bytes of one row group of a unit
                    const uint32_t yrow0 = yreg + s * ypitch_b;
                    if (p.dbg != 2) {
#define RS_GO(UC_, NN_) rs_issue_step_k<UC_, NN_>(ksteps, issuer, nunits, ngr, ngr_pitch, p.RU, p.setw, tmem_base, xs0, upitch, yrow0, ypitch_b)
                        if (p.UC == 1) RS_GO(1, 88);
                        else if (p.UC == 2 && p.RU == 1) RS_GO(2, 88);
                        else if (p.UC == 2) RS_GO(2, 120);
                        else if (p.RU == 1) RS_GO(4, 88);
                        else if (p.RU == 2) RS_GO(4, 120);
                        else RS_GO(4, 184);
#undef RS_GO
                    }
                    dbg_issue += clock64() - t2;
                    ++dbg_steps;
                    umma_commit(&emptyx[kx % kRsXSlots]);
                    ++kx;
                    // plane q = s is dead after this step; its barrier slot is the one of load (ky - 3)
                    umma_commit(&emptyy[4 * k + ((ky[k] - 3) & 3)]);
                    if (s == im[k].nsteps - 1) {
                        umma_commit(&emptyy[4 * k + ((ky[k] - 2) & 3)]);
                        umma_commit(&emptyy[4 * k + ((ky[k] - 1) & 3)]);
                        umma_commit(&item_done[k]);
                    }
                }
            }
        }
        umma_commit(done);
        if (p.dbg == 5 && issuer == 0) {
            atomicAdd(&rs_dbg[0], (unsigned long long)dbg_y0); atomicAdd(&rs_dbg[1], (unsigned long long)dbg_y);
            atomicAdd(&rs_dbg[2], (unsigned long long)dbg_x); atomicAdd(&rs_dbg[3], (unsigned long long)dbg_issue);
            atomicAdd(&rs_dbg[4], (unsigned long long)dbg_steps); atomicAdd(&rs_dbg[5], (unsigned long long)(clock64() - dbg_t0));
            atomicAdd(&rs_dbg[6], 1ULL);
        }
    }
    __syncwarp();
    if (warp < 4 && blockIdx.x < items) {
        // =============================== epilogue: TMEM -> fp32 atomics into dw ===============================
        mbar_wait(done, 0, err, 45);
        tc_fence_after();
        // accumulator row m = kw * (8 UC) + (row in group * CU + chunk in unit) * 8 + channel.  M = 64 (UC = 1, 2): row m lives in lane
        // (m % 16) + 32 * (m / 16); M = 128 (UC = 4): row m lives in lane m.  N-group 4 * rho + 4a + dq of the row with index rho in its
        // group: dy row h-1+a of plane d-1+dq.
        const int m = p.UC == 4 ? warp * 32 + lane : (lane < 16 ? warp * 16 + lane : 64);
        const int uw = 8 * p.UC;
        const int kw = m / uw, sub = (m - kw * uw) >> 3;
        const int rho = sub / p.CU, cu = sub - rho * p.CU;
        const int cin = p.C0 + p.C1;
        const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int nunits = nch / p.CU;
        for (int u = 0; u < nunits; ++u) {
            const int ch = ch0 + u * p.CU + cu;
            const bool valid = kw < 3;
            // (tcgen05.ld takes one column address per warp: walk the row-in-group index uniformly, the lanes it belongs to store)
            for (int rr = 0; rr < p.RU; ++rr) {
                for (int a = 0; a < 3; ++a) {
                    for (int dq = 0; dq < 3; ++dq) {
                        float v[8];
                        tmem_ld8(tlane + u * p.setw + (4 * rr + 4 * a + dq) * 8, v);
                        tmem_wait_ld();
                        if (valid && rho == rr) {
                            const int tap = ((2 - dq) * 3 + (2 - a)) * 3 + kw;
                            float* dst = dw + (((size_t)g * 27 + tap) * cin + ch * 8 + (m & 7)) * p.Cout + cochunk * 8;
#pragma unroll
                            for (int q = 0; q < 8; ++q) atomicAdd(dst + q, v[q]);
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

int env_int(const char* name, int dflt) { const char* s = getenv(name); return s ? atoi(s) : dflt; }

}  // namespace

// developer probe: reads and clears the issuer-0 cycle counters gathered by launches with PB_WG_RS=5
extern "C" int pb_wgrad_rs_debug(unsigned long long* out8) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out8, rs_dbg, sizeof(z)) != cudaSuccess) return PB_ECUDA;
    if (cudaMemcpyToSymbol(rs_dbg, z, sizeof(z)) != cudaSuccess) return PB_ECUDA;
    return PB_OK;
}

// Internal entry (called by pb_conv3d_wgrad_tc, which has validated the descriptor): PB_EUNSUPPORTED = class not covered, the
// caller falls back to the kh-stacked kernels.
int pb_wgrad_rs_launch(const pb_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw, int* err_flag,
                       cudaStream_t st) {
    // read per call (a getenv is noise beside a launch) so that a developer script can compare both paths in one process
    const int mode = env_int("PB_WG_RS", 1);
    if (mode == 0) return PB_EUNSUPPORTED;
    RsP p;
    p.dbg = mode;
    p.N = d->n; p.D = d->di; p.H = d->hi; p.W = d->wi; p.C0 = d->c0; p.C1 = d->c1; p.Cout = d->cout;
    p.reflect = d->pad_mode == PB_PAD_REFLECT;
    p.KW = (p.W + 15) & ~15;
    if (p.KW > 128) return PB_EUNSUPPORTED;
    // measured (scripts/bench_wgrad.py, profiles/r02_wgrad_rs_classes.txt): 1.2-2.1x faster than the kh-stacked kernels from 20^3 up; at
    // 10^3 a launch is a handful of items and the per-item prologue (three dy planes) dominates: those stay on the old kernels
    if ((long long)p.D * p.H * p.W < env_int("PB_WG_RS_MIN_VOX", 4000)) return PB_EUNSUPPORTED;
    p.npg = d->n / d->groups; p.groups = d->groups;
    p.nT = (d->c0 + d->c1) / 8;
    p.nP = d->cout / 8;
    p.TG = p.nT < kRsMaxChunks ? p.nT : kRsMaxChunks;
    p.ntg = (p.nT + p.TG - 1) / p.TG;
    // chunks per MMA: 32- or 16-channel rows when the chunk count allows (PB_WG_RS_UC caps it); narrower units are filled up with
    // RU consecutive input ROWS side by side (PB_WG_RS_RU caps it), so that an instruction always carries 32 "channels" (M = 128)
    p.CU = p.nT % 4 == 0 ? 4 : (p.nT % 2 == 0 ? 2 : 1);
    { const int cap = env_int("PB_WG_RS_UC", 4); while (p.CU > cap) p.CU /= 2; }
    // (measured: RU > 1 is SLOWER — c16->8 80^3 127 -> 147 us with two rows per M = 128 / N = 120 instruction, c8->8 105 -> 151 us with four rows
    // per N = 184 instruction: those shapes fetch 7.75 / 9.75 KB of operands per instruction, more than the 128 B per clock the shared memory
    // delivers in their MAC time, and in the kernel they retire in 94 / 172 cycles.  Kept behind PB_WG_RS_RU, default one row.)
    p.RU = 4 / p.CU;
    { const int cap = env_int("PB_WG_RS_RU", 1); while (p.RU > cap) p.RU /= 2; }
    p.UC = p.CU * p.RU;
    p.NN = 8 * (11 + 4 * (p.RU - 1));
    p.setw = (p.NN + 31) & ~31;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < (p.TG / p.CU) * p.setw) p.tmem_cols *= 2;
    if (p.tmem_cols > 512) return PB_EUNSUPPORTED;
    p.h_lo = p.reflect ? -1 : 0; p.rows_t = p.reflect ? p.H + 2 : p.H;
    p.d_lo = p.reflect ? -1 : 0; p.planes_t = p.reflect ? p.D + 2 : p.D;
    if ((long long)p.H * p.W * (p.C0 > p.C1 ? p.C0 : p.C1) >= (1LL << 30)) return PB_EUNSUPPORTED;          // int offsets within a plane
    p.prefetch = env_int("PB_WG_RS_PREFETCH", 0);
    p.ybulk = (p.Cout == 8 && p.KW == p.W && env_int("PB_WG_RS_YBULK", 1)) ? 1 : 0;
    p.nregions = 2;                                               // one per item of a pair
    const int gy = p.groups * p.ntg * p.nP;
    const int ctas_max = 148 / gy < 1 ? 1 : 148 / gy;
    // rows per strip: bounded by the copy budget of the input-row producers and by shared memory
    int nr = env_int("PB_WG_RS_ROWS", kRsMaxRows);
    if (nr > kRsMaxRows) nr = kRsMaxRows;
    if (nr > p.rows_t) nr = p.rows_t;
    nr = ((nr + p.RU - 1) / p.RU) * p.RU;                          // whole row groups (pad rows are zero-filled)
    auto smem_of = [&](int rows, int steps) {
        return (size_t)p.nregions * (4 * (rows + 2) + steps + 2) * p.KW * 16 + (size_t)kRsXSlots * p.TG * rows * (p.KW + 8) * 16 + 256 + 1024;
    };
    auto too_big = [&](int rows) { return p.TG * rows * (p.KW + 8) > kRsXCopies * kRsXProducers || smem_of(rows, kRsMaxSteps) > 227 * 1024; };
    while (nr > p.RU && too_big(nr)) nr -= p.RU;
    if (too_big(nr)) return PB_EUNSUPPORTED;
    p.nstrips = (p.rows_t + nr - 1) / nr;
    p.nr = (p.rows_t + p.nstrips - 1) / p.nstrips;
    p.nr = ((p.nr + p.RU - 1) / p.RU) * p.RU;                     // <= nr
    // depth chunks: at most kRsMaxSteps planes per item, at least 6 (the two halo planes of dy are loaded per item); among those the
    // count that fills whole rounds of CTAs best (an item is ~25 us of work: a ragged last round is the largest loss)
    int best_nd = 0;
    double best_eff = -1.0;
    const int nd_force = env_int("PB_WG_RS_ND", 0);
    for (int nd = (p.planes_t + kRsMaxSteps - 1) / kRsMaxSteps; nd == (p.planes_t + kRsMaxSteps - 1) / kRsMaxSteps || p.planes_t / nd >= 6; ++nd) {
        const int it = p.npg * p.nstrips * nd;
        const int ct = it < ctas_max ? it : ctas_max;
        const int rounds = (it + ct - 1) / ct;
        // whole-round efficiency, discounted by the per-item overhead of the two extra dy planes
        const double steps = (double)p.planes_t / nd;
        const double eff = (double)it / ((double)rounds * ct) * steps / (steps + 1.0);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_nd = nd; }
        if (nd > p.planes_t) break;
    }
    p.ND = nd_force > 0 ? nd_force : best_nd;
    const int max_steps = (p.planes_t + p.ND - 1) / p.ND;
    if (max_steps > kRsMaxSteps) return PB_EUNSUPPORTED;
    p.yrows = 4 * (p.nr + 2) + max_steps + 2;
    const int items = p.npg * p.nstrips * p.ND;
    const int ctas = items < ctas_max ? items : ctas_max;
    const size_t smem = (size_t)p.nregions * p.yrows * p.KW * 16 + (size_t)kRsXSlots * p.TG * p.nr * (p.KW + 8) * 16 + (2 * kRsXSlots + 19) * 8 + 16 + 1024;
    if (smem > 227 * 1024) return PB_EUNSUPPORTED;
    // dense NDHWC dy seen as (8 ch of a chunk, W, H, Cout/8 chunks, N*D planes); box = one row of KW positions of one chunk
    CUtensorMap ymap;
    {
        EncodeTiledFn enc = encode_tiled();
        if (enc == nullptr) return PB_EUNSUPPORTED;
        const cuuint64_t dims[5] = {8, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)(p.Cout / 8), (cuuint64_t)p.N * p.D};
        const cuuint64_t strides[4] = {(cuuint64_t)p.Cout * 2, (cuuint64_t)p.W * p.Cout * 2, 16, (cuuint64_t)p.H * p.W * p.Cout * 2};
        const cuuint32_t box[5] = {8, (cuuint32_t)p.KW, 1, 1, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (enc(&ymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return PB_EUNSUPPORTED;
    }
    cudaError_t e = cudaFuncSetAttribute(conv3_wgrad_rs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pb_set_error("conv3d_wgrad_rs: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PB_ECUDA; }
    conv3_wgrad_rs_kernel<<<dim3(ctas, gy), kRsThreads, smem, st>>>(ymap, p, (const bf16*)x0, (const bf16*)x1, (const bf16*)dy, dw, err_flag);
    return 0;
}
