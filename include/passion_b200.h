/* passion_b200 — C ABI of the B200-native PASSION training hot path.
 *
 * The reference (Jun-Jie-Shi/PASSION) is pure Python/PyTorch and has no FFI layer; its hot
 * path reaches the GPU through torch.nn library modules.  Each entry point below replaces one
 * of those library call sites (cited as reference file:line) with a hand-written sm_100a
 * kernel.  The Python host side (passion_b200/ops.py) binds these with ctypes and exposes them
 * behind the reference's own module interface (Model.forward, criterions.*_bs).
 *
 * Conventions
 *   - every function returns 0 on success, a negative PB_E* code otherwise;
 *     pb_last_error() returns a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch); no hidden allocation,
 *     no hidden synchronisation; every kernel is enqueued on `stream` and is CUDA-graph
 *     capture safe.  Scratch comes from caller-provided workspaces.
 *   - activations are dense NDHWC ("channels last"): [N][D][H][W][C], C fastest.
 *     dtype 0 = float32 (check mode), 1 = bfloat16 (storage only; all arithmetic accumulates in fp32).
 *   - statistics buffers are float64 so that the atomically accumulated sums are order-insensitive
 *     to ~1e-16 relative.
 */
#ifndef PASSION_B200_H
#define PASSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pb_stream_t;            /* cudaStream_t */

enum { PB_OK = 0, PB_EINVAL = -1, PB_ECUDA = -2, PB_EUNSUPPORTED = -3 };
enum { PB_F32 = 0, PB_BF16 = 1 };
enum { PB_PAD_ZERO = 0, PB_PAD_REFLECT = 1 };

int         pb_version(void);
const char* pb_last_error(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
long long   pb_launch_count(void);

/* ---- 3-D convolution -------------------------------------------------------------------
 * Replaces nn.Conv3d inside general_conv3d (models/blocks.py:354-370, conv at :357) and the plain
 * heads (models/rfnet.py:69,107; models/blocks.py:407,455).
 * Input = channel concatenation of x0 (C0 channels) and x1 (C1 channels, may be NULL/0):
 * the torch.cat((a, b), dim=1) feeding a conv (rfnet.py:75,79,83,133,139,145; blocks.py:463) is
 * never materialised.  Samples are split evenly into `groups` weight groups (the four
 * modality encoders, rfnet.py:234-237, run as one launch with groups = 4).
 * Weights are fp32 in the kernel layout [groups][k^3 taps][Cin][Cout] (fwd, wgrad output)
 * or [groups][k^3 taps][Cout][Cin] (dgrad).  ksize in {1,3}; stride in {1,2}; padding = ksize/2.
 */
typedef struct pb_conv_desc {
    int32_t dtype;                 /* PB_F32 | PB_BF16 : storage type of x, y, dy, dx        */
    int32_t n;                     /* batch                                                   */
    int32_t di, hi, wi;            /* input spatial size                                      */
    int32_t dout, ho, wo;          /* output spatial size                                     */
    int32_t c0, c1;                /* channels of the two concatenated inputs (c1 may be 0)   */
    int32_t cout;
    int32_t ksize, stride;
    int32_t pad_mode;              /* PB_PAD_ZERO | PB_PAD_REFLECT                            */
    int32_t groups;
} pb_conv_desc;

/* y = conv(cat(x0,x1)) (+bias).  If stats != NULL it must be zero-filled [n][cout][2] float64 and
 * receives per-(sample,channel) sum and sum of squares of the fp32 accumulators (the
 * InstanceNorm statistics, blocks.py:18) — fused into the conv epilogue. */
int pb_conv3d_fwd(const pb_conv_desc* d, const void* x0, const void* x1, const float* w,
                  const float* bias, void* y, double* stats, pb_stream_t stream);
/* dx0/dx1 = gradient w.r.t. the two inputs (exact adjoint incl. reflect padding). wt is the
 * [groups][taps][Cout][Cin] layout. */
int pb_conv3d_dgrad(const pb_conv_desc* d, const void* dy, const float* wt, void* dx0, void* dx1,
                    pb_stream_t stream);
/* dw[groups][taps][Cin][Cout] += sum over samples/voxels; dw must be zero-filled by the caller. */
int pb_conv3d_wgrad(const pb_conv_desc* d, const void* x0, const void* x1, const void* dy,
                    float* dw, pb_stream_t stream);

/* tcgen05 / TMEM implicit-GEMM path for the dominant class: bf16, 3x3x3, stride 1, same-size output, c0/c1/cout
 * multiples of 8, c0 + c1 <= 64 (csrc/conv3d_tc.cu).  `wimg` is the bf16 weight image
 * [groups][cout tiles][27][max(2,cin/8)][NT][8] with NT = pb_conv3d_tc_ntile(cin, cout) (0 = class unsupported).
 * The output may be split channel-wise into y0 (co0 channels) | y1 (co1 channels) — used by the data gradient
 * of a two-source conv, which is this same kernel run on dy with flipped/transposed weights and zero padding
 * (pb_conv3d_dgrad_reflect_fix then adds the reflected-halo terms).  `err_flag` is a zero-initialised device
 * int that receives a non-zero code if the kernel's internal pipeline ever times out. */
int pb_conv3d_tc_ntile(int cin, int cout);
/* 1 = the class runs on the kw-stacked variant (Cout <= 16: 3 MMAs of N = 144 per input plane instead of 9 of N = 48) and its
 * weight image is laid out [groups][cout tiles][3 kh][max(2,cin/8)][rows: kd = 2,1,0 | kw | 16 co][8]; else the layout above
 * ([9 (kh,kw)][chunk][rows: kd = 2,1,0 | NT co][8]).  pb_weight_prep writes whichever applies. */
int pb_conv3d_tc_kws(int cin, int cout);
int pb_conv3d_tc(const pb_conv_desc* d, const void* x0, const void* x1, const void* wimg, const float* bias /* [groups][cout] or NULL */,
                 void* y0, void* y1, int co0, int co1, double* stats, int* err_flag, pb_stream_t stream);
/* Weight gradient of the same class on tcgen05: voxels are the GEMM K dimension, the nine (kd,kh) accumulators of one
 * 8-channel input chunk stay resident in TMEM for the CTA's whole sweep.  dw is the fp32 [groups][27][cin][cout]
 * layout of pb_conv3d_wgrad and must be zero-filled; cout in {8,16,32,64}. */
int pb_conv3d_wgrad_tc(const pb_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw,
                       int* err_flag, pb_stream_t stream);
/* 1x1x1 weight gradient on the tensor cores: dw [groups][cin][cout] += sum_v x[v][ci] dy[v][co] (bf16, c0, c1, cout
 * multiples of 8, cin <= 256, cout <= 64; PB_EUNSUPPORTED otherwise).  Same autograd node as pb_conv3d_wgrad. */
int pb_conv1_wgrad_tc(const pb_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw,
                      int* err_flag, pb_stream_t stream);
/* dx0/dx1 += the contributions that reach an input voxel through the reflect padding (voxels one step inside a
 * face); completes a zero-padding data gradient into the exact adjoint of the reflect-padded forward conv. */
int pb_conv3d_dgrad_reflect_fix(const pb_conv_desc* d, const void* dy, const float* wt, void* dx0, void* dx1,
                                pb_stream_t stream);
/* Data gradient of a REFLECT-padded 3x3x3 conv on the tensor cores (replaces the same autograd node as pb_conv3d_dgrad):
 *   pb_conv3d_tc_full : "full" correlation of dy [n,di,hi,wi,c0] with the flipped/transposed weight image on the domain
 *                       grown by one voxel per side (d->dout = di+2 ...).  Voxels whose value is final (not within two
 *                       voxels' reach of a reflection: interior coordinate not in {1, size-2} on any axis) are written
 *                       straight into y0|y1 [n,di,hi,wi,co0|co1]; all others into yext [n,di+2,hi+2,wi+2,co0+co1].
 *   pb_reflect_fold   : y0|y1[v] = sum of the yext voxels that the reflect padding maps onto v, for the remaining voxels
 *                       (per axis: {i+1} U {0 if i == 1} U {size+1 if i == size-2}).  bf16, sizes >= 4. */
/* 1x1x1 conv (and, on dy with the transposed image, its data gradient) as a TMA-fed tcgen05 GEMM (csrc/conv1_tc.cu): bf16, channels
 * multiples of 8 (cout also 2 or 4), voxels per sample >= 128.  `wimg` = [groups][cout tiles][cin/8 rounded up to even][NT][8] bf16
 * with NT = pb_conv1_tc_ntile(cin, cout) (0 = class not covered; pb_weight_prep writes the image for ksize = 1 when nt is set).
 * Output split y0 (co0 channels) | y1 (co1 channels) as in pb_conv3d_tc; optional bias and InstanceNorm sums. */
int pb_conv1_tc_ntile(int cin, int cout);
int pb_conv1_tc(const pb_conv_desc* d, const void* x0, const void* x1, const void* wimg, const float* bias, void* y0, void* y1,
                int co0, int co1, double* stats, int* err_flag, pb_stream_t stream);
/* Token-path GEMM of the mmFormer transformer blocks on tcgen05 (csrc/gemm_tc.cu) — replaces torch.nn.Linear of SelfAttention
 * (qkv, proj) and FeedForward, reference models/mmformer.py:192-280, forward and both gradients:
 *     D[m][n] = sum_k A(m,k) * B(n,k) (+ bias[n]),   bf16 operands, fp32 accumulation, D bf16 (d_fp32 = 0) or fp32 (1), ldd elements per row.
 * a_kmajor = 1: A is stored [M][K] (lda elements between rows); 0: A is stored [K][M] (A(m,k) at a[k*lda + m]); b likewise with N.
 *     forward  Y = X W^T + b : A = X (1), B = W [N][K] (1);   dX = dY W : A = dY (1), B = W read as [K = N'][K'] (0);
 *     dW = dY^T X : A = dY read as [K = M][N'] (0), B = X read as [K = M][K'] (0), D fp32.
 * lda, ldb multiples of 8, operands 16-byte aligned.  err_flag as in pb_conv3d_tc.  Outputs with few tiles and a long K run split-K:
 * pb_gemm_tc_workspace_floats(M, N, K) > 0 is the size of the ZERO-FILLED fp32 workspace the call then needs (partials are added
 * with atomics, a second small launch adds the bias and converts); otherwise workspace may be NULL. */
long long pb_gemm_tc_workspace_floats(int M, int N, int K);
/* developer probe of csrc/conv3d_wgrad_rs.cu (launches made with PB_WG_RS=5): issuer-thread cycle counters, read and cleared */
int pb_wgrad_rs_debug(unsigned long long* out8);
/* same for the kw-stacked forward / data-gradient kernel of csrc/conv3d_tc.cu (launches made with PB_TC_PROBE=1) */
int pb_conv3d_tc_debug(unsigned long long* out8);
int pb_gemm_tc(const void* a, const void* b, const float* bias, void* d, float* workspace, int M, int N, int K, int lda, int ldb,
               int ldd, int a_kmajor, int b_kmajor, int d_fp32, int* err_flag, pb_stream_t stream);
/* The same GEMM once per (b0, b1) — the attention products of SelfAttention.forward (reference models/mmformer.py:203-213), one
 * problem per (sample, head): S = Q K^T, O = P V and the four gradient products, with Q / K / V read in place as column blocks of the
 * [tokens][3 * dim] qkv rows.  sa0 / sa1, sb0 / sb1, sd0 / sd1 = element strides of A, B, D over the two batch indices (multiples of
 * 8 for the operands); rows beyond M / N / K inside one batch entry read as zero, never as the next entry's data.  No bias, no split-K. */
int pb_gemm_tc_batched(const void* a, const void* b, void* d, int M, int N, int K, int lda, int ldb, int ldd, int a_kmajor, int b_kmajor,
                       int d_fp32, int nb0, int nb1, long long sa0, long long sa1, long long sb0, long long sb1, long long sd0,
                       long long sd1, int* err_flag, pb_stream_t stream);
/* Row kernels of the token path (csrc/attn.cu).
 * pb_attn_softmax_fwd — reference mmformer.py:206-208: p = softmax(scale * s) over the T keys of each of `rows` fp32 score rows
 *   (row stride T), written as bf16 with row stride ldp >= T (a multiple of 8 so that p can be a TMA operand; pad columns are
 *   written as zero); with drop_p > 0 also p_drop = p * keep / (1 - drop_p), keep drawn from a counter hash of the device-side
 *   int64 *seed (give a fresh one per call), else p_drop = NULL.
 * pb_attn_softmax_bwd — ds = scale * (p_drop .* dp - p * sum_j(p_drop_j dp_j)) (bf16, row stride ldp) from the fp32 gradient dp
 *   (row stride T) with respect to p_drop; without dropout pass p for p_drop.
 * pb_layernorm_fwd / _bwd — nn.LayerNorm over the last dimension (mmformer.py:233-250), C in {256, 512, 1024}, x / y / dy / dx in
 *   `dtype` (PB_F32 | PB_BF16), affine parameters and statistics fp32; mean / rstd [rows] are written by the forward and read by
 *   the backward; dw / db [C] fp32 are ACCUMULATED into (zero them first). */
int pb_attn_softmax_fwd(const float* s, void* p, void* p_drop, long long rows, int T, int ldp, float scale, float drop_p,
                        const long long* seed, pb_stream_t stream);
int pb_attn_softmax_bwd(const float* dp, const void* p, const void* p_drop, void* ds, long long rows, int T, int ldp, float scale,
                        pb_stream_t stream);
/* The same two softmax steps FUSED with the score products (csrc/gemm_tc.cu attn_rows_kernel; head width d <= 64): no [N][H][T][T] fp32
 * tensor is written.  q / k / v / d_o are [N][T][H][d] views: ld = elements between tokens, s0 between samples, s1 between heads.
 *   pb_attn_scores_softmax: p (and p_drop) = softmax(scale * q k^T) per (sample, head), as pb_attn_softmax_fwd would write them;
 *   pb_attn_delta:          delta[n][h][t] = sum_c d_o[n][t][h][c] * o[n][t][h][c]  (contiguous d_o, o);
 *   pb_attn_dsoftmax:       ds = scale * (p_drop .* (d_o v^T) - p * delta), as pb_attn_softmax_bwd would write it. */
int pb_attn_scores_softmax(const void* q, const void* k, void* p, void* p_drop, int N, int H, int T, int d, int ld, long long s0,
                           long long s1, int ldp, float scale, float drop_p, const long long* seed, int* err_flag, pb_stream_t stream);
int pb_attn_delta(const void* d_o, const void* o, float* delta, int N, int T, int H, int d, pb_stream_t stream);
int pb_attn_dsoftmax(const void* d_o, const void* v, const void* p, const void* p_drop, const float* delta, void* ds, int N, int H,
                     int T, int d, int ld_o, long long so0, long long so1, int ld_v, long long sv0, long long sv1, int ldp,
                     float scale, int* err_flag, pb_stream_t stream);
int pb_layernorm_fwd(int dtype, const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, long long rows, int C,
                     float eps, pb_stream_t stream);
int pb_layernorm_bwd(int dtype, const void* dy, const void* x, const float* mean, const float* rstd, const float* w, void* dx, float* dw,
                     float* db, long long rows, int C, pb_stream_t stream);
int pb_conv3d_tc_full(const pb_conv_desc* d, const void* x, const void* wimg, void* y0, void* y1, int co0, int co1,
                      void* yext, int* err_flag, pb_stream_t stream);
int pb_reflect_fold(const void* yext, void* y0, void* y1, int n, int d, int h, int w, int co0, int co1, pb_stream_t stream);

/* ---- 3x3x3 stride-1 convs with very few channels (csrc/conv3d_small.cu): shared-memory tiled FFMA kernels for the
 * classes pb_conv3d_small_supported(cin, cout) reports (1->8: first encoder conv, rfnet.py:24 / mmformer.py:28;
 * 2->2 and 4->4: PRM embedding layers, blocks.py:399-401).  Single source (c1 = 0), fp32 or bf16 storage.
 *   fwd   : w = [G][27][cin][cout] f32, optional bias [G][cout], optional stats as pb_conv3d_fwd
 *   dgrad : wt = [G][27][cout][cin] f32; reflect padding needs `ext`, a scratch of n*(di+2)*(hi+2)*(wi+2)*cin elements
 *   wgrad : dw [G][27][cin][cout] f32, accumulated with atomics (zero-filled by the caller) */
int pb_conv3d_small_supported(int cin, int cout);
int pb_conv3d_small_fwd(const pb_conv_desc* d, const void* x, const float* w, const float* bias, void* y, double* stats,
                        pb_stream_t stream);
int pb_conv3d_small_dgrad(const pb_conv_desc* d, const void* dy, const float* wt, void* dx, void* ext, pb_stream_t stream);
int pb_conv3d_small_wgrad(const pb_conv_desc* d, const void* x, const void* dy, float* dw, pb_stream_t stream);

/* ---- weight layout conversion ------------------------------------------------------------
 * The parameters stay in nn.Conv3d's layout [cout][cin][k][k][k] fp32 (models/blocks.py:357, state_dict compatible);
 * pb_weight_prep gathers, for up to 4 weight groups (the four modality encoders, rfnet.py:234-237), every layout the
 * conv kernels of one layer need in ONE launch (NULL output = skip):
 *   wk  [G][taps][cin][cout] f32 (pb_conv3d_fwd / _wgrad)      wt   [G][taps][cout][cin] f32 (pb_conv3d_dgrad)
 *   img  = pb_conv3d_tc weight image, nt  = pb_conv3d_tc_ntile(cin, cout)
 *   imgT = weight image of the data gradient (taps mirrored, channels transposed), ntT = pb_conv3d_tc_ntile(cout, cin)
 *   bias [G][cout] f32 gathered from b[g]
 * pb_weight_grad_unpack scatters a kernel-layout weight gradient dw [G][taps][cin][cout] (and the bias gradient, given
 * either as db [G][cout] or as the channel sums of dy) back into per-group gradients in the parameter layout. */
typedef struct {
    const float* w[4];
    const float* b[4];
    int groups, cin, cout, ksize;
    float* wk;
    float* wt;
    void* img;
    int nt;
    void* imgT;
    int ntT;
    float* bias;
    int w_cin_stride;       /* 0 = cin.  > cin: each w[g] points at a cin-SLICE of a wider parameter [cout][w_cin_stride][k][k][k]
                             * (the single-modality decoder passes read one modality's quarter of a 4C-input 1x1x1 conv) */
} pb_weight_prep_desc;
typedef struct {
    const float* dw;
    const float* db;        /* bias gradient [G][cout], or NULL ... */
    const double* dy_stats; /* ... or the pb_channel_stats of dy [n][cout][2]: db[g][c] = sum over the group's npg samples */
    int npg;
    float* gw[4];
    float* gb[4];
    int groups, cin, cout, ksize;
    int accumulate;         /* 0: gw / gb are overwritten; 1: added to (autograd's accumulation into an existing .grad) */
    int w_cin_stride;       /* as in pb_weight_prep_desc: gw[g] points at a cin-slice of a wider gradient tensor */
} pb_weight_unpack_desc;
int pb_weight_prep(const pb_weight_prep_desc* d, pb_stream_t stream);
int pb_weight_grad_unpack(const pb_weight_unpack_desc* d, pb_stream_t stream);
/* Batched forms: ALL conv layers of a model in one launch each (the weights change once per step, after the optimizer; the
 * weight gradients are complete once backward has finished) instead of one launch per layer and direction.
 * `table` = device scratch of pb_weight_batch_table_bytes(n) bytes.  upload != 0: the n descriptors are packed on the host
 * and copied into `table` first (a pageable-memory copy: NOT capturable — do it in a warm-up step); upload == 0: the table
 * written by the previous call is reused as is (same descriptors; capturable into a CUDA graph). */
size_t pb_weight_batch_table_bytes(int n);
int pb_weight_prep_batch(const pb_weight_prep_desc* descs, int n, void* table, int upload, pb_stream_t stream);
int pb_weight_grad_unpack_batch(const pb_weight_unpack_desc* descs, int n, void* table, int upload, pb_stream_t stream);

/* ---- InstanceNorm3d(affine=False, eps) + LeakyReLU(slope) (+ residual) -----------------
 * Replaces norm + activation of general_conv3d (blocks.py:18, :363, :367-369) and the encoder
 * residual add (rfnet.py:37,40,43,46).
 *   pb_inorm_finalize : stats (sum, sumsq) -> mr[n][c] = (mean, rstd) float32
 *   pb_inorm_lrelu_fwd: out = lrelu((y-mean)*rstd) (+ res)
 *   pb_inorm_lrelu_bwd: given dout, y, mr -> dy ; sums is a zero-filled [n][c][2] float64 scratch.
 */
/* sum / sum of squares per (sample, channel) of an arbitrary tensor: the statistics of a PRE-norm block
 * (general_conv3d_prenorm, blocks.py:300-316: InstanceNorm -> LeakyReLU -> Conv).  stats zero-filled [n][c][2] float64. */
int pb_channel_stats(int dtype, const void* x, double* stats, int n, long long voxels, int c, pb_stream_t stream);
int pb_inorm_finalize(const double* stats, float* mr, int n, int c, long long voxels, float eps,
                      pb_stream_t stream);
int pb_inorm_lrelu_fwd(int dtype, const void* y, const float* mr, const void* res, void* out,
                       int n, long long voxels, int c, float slope, pb_stream_t stream);
/* finalize + forward apply in one launch (mr is an OUTPUT here: written for the backward pass) */
int pb_inorm_lrelu_fwd_stats(int dtype, const void* y, const double* stats, float* mr, const void* res, void* out,
                             int n, long long voxels, int c, float eps, float slope, pb_stream_t stream);
int pb_inorm_lrelu_bwd(int dtype, const void* dout, const void* y, const float* mr, double* sums,
                       void* dy, int n, long long voxels, int c, float slope, pb_stream_t stream);

/* ---- trilinear up-sampling, align_corners=True, integer scale ---------------------------
 * Replaces nn.Upsample (rfnet.py:54,59,64,110-112,208-210). */
int pb_upsample_fwd(int dtype, const void* x, void* y, int n, int d, int h, int w, int c, int scale,
                    pb_stream_t stream);
int pb_upsample_bwd(int dtype, const void* dy, void* dx, int n, int d, int h, int w, int c, int scale,
                    pb_stream_t stream);
/* one axis of the separable adjoint: in [outer][n_big][inner] -> out [outer][n_small][inner] (inner % c == 0, c = the
 * channel count that fixes the vector width); the host composes W, H, D with caller-allocated intermediates. */
int pb_upsample_bwd_axis(int dtype, const void* in, void* out, long long outer, int n_big, int n_small,
                         long long inner, int c, pb_stream_t stream);

/* ---- region-aware modal fusion (models/blocks.py:495-517, 597-616) ----------------------
 * y  [n][v][K*C]  masked modality features (channel = k*C + c), storage dtype
 * p  [n][v][4]    float32 class probabilities (softmax of the PRM logits, detached)
 *   pool : S[n][i][k*C+c] = sum_v y*p_i   (float64, zero-filled), Psum[n][i] = sum_v p_i
 *   mix  : R[n][v][i*C+c] = p_i * sum_k gate[n][i][k] * y[n][v][k*C+c]
 *   mix_bwd_gate : dgate[n][i][k] = sum_{v,c} p_i * y[k*C+c] * dR[i*C+c]   (float64, zero-filled)
 *   bwd_y : dy[n][v][k*C+c] = sum_i p_i * (gate[n][i][k]*dR[n][v][i*C+c] + dS[n][i][k*C+c])
 */
/* Gate MLP of modal_fusion (blocks.py:507-513) fused: pooled float64 sums of pb_rfm_pool -> gate [n][4][K] (K = 4: all modality
 * slots; K = 1: single-modality samples, sample n = modality n / b, gate [n][4]) plus the hidden pre-activations z1 [n][4][128]
 * for the backward pass.  w / dw: 16 device pointers = 4 classes x {w0 [128][4c+1], b0 [128], w2 [4][128], b2 [4]} in that
 * order (w0 of all classes first).  pb_rfm_gate_bwd ACCUMULATES the parameter gradients into zero-filled dw buffers. */
int pb_rfm_gate_fwd(const float* const* w, const double* S, const double* Psum, float* z1, float* gate, int n, int b,
                    long long voxels, int k, int c, pb_stream_t stream);
int pb_rfm_gate_bwd(const float* const* w, float* const* dw, const double* S, const double* Psum, const float* z1,
                    const double* dgate, float* dS, int n, int b, long long voxels, int k, int c, pb_stream_t stream);
/* MaskModal (rfnet.py:154-163, 239-242; mmformer.py:316-326) for `passes` decoder passes in one launch:
 *   out[p*b + i][v][m*c + ch] = enc[m*b + i][v][ch] * ms[p][i][m]     (enc: modality-major output of the grouped encoders)
 * and its adjoint denc[m*b + i][v][ch] = sum_p ms[p][i][m] * dout[p*b + i][v][m*c + ch]. */
int pb_masked_stack_fwd(int dtype, const void* enc, const float* ms, void* out, int passes, int b, long long voxels, int c,
                        pb_stream_t stream);
int pb_masked_stack_bwd(int dtype, const void* dout, const float* ms, void* denc, int passes, int b, long long voxels, int c,
                        pb_stream_t stream);
int pb_rfm_pool(int dtype, const void* y, const float* p, double* S, double* Psum,
                int n, long long voxels, int kc, pb_stream_t stream);
int pb_rfm_mix(int dtype, const void* y, const float* p, const float* gate, void* r,
               int n, long long voxels, int k, int c, pb_stream_t stream);
int pb_rfm_mix_bwd_gate(int dtype, const void* y, const float* p, const void* dr, double* dgate,
                        int n, long long voxels, int k, int c, pb_stream_t stream);
int pb_rfm_bwd_y(int dtype, const float* p, const float* gate, const void* dr, const float* dS,
                 void* dy, int n, long long voxels, int k, int c, pb_stream_t stream);

/* ---- PASSION objective (utils/criterions.py:25-38, 59-76, 92-103, 144-180), csrc/loss.cu --------------------
 * The loss kernels read class LABELS (uint8 [b][V]) instead of the float64 one-hot target; sample n of a prediction
 * batch uses the labels / teacher of sample n % b (the 5B / 4B batched decoder passes).  Probabilities are float32
 * [n][V][4] at label resolution (low-resolution PRM heads are first brought there by pb_upsample_fwd, the
 * reference's up_op).  All sums are float64 and must be zero-filled by the caller.
 *   pb_softmax4      probs = softmax(logits * inv_temp) over the 4 classes (storage dtype in, f32 out)   (:93-94)
 *   pb_cedice_fwd    sums[n][12]: A_c = sum p_c t_c | L_c = sum p_c | E_c = sum t_c log(clamp(p_c,.005,1)) (:30-32, :69)
 *   pb_cedice_bwd    dprobs[n][V][4] = coef[n][4+c] + [t==c](coef[n][c] + coef[n][8+c] * d log clamp / dp)
 *   pb_kl_fwd/bwd    sums[n] = sum_{v,c} pt (log pt - log ps), both clamped to [.005, 1]; dps = coef[n] * d/dps (:98-101)
 *   pb_proto_sums    P[n][4][8] = sum_v f[v] [t_v == i]                                                  (:158-159)
 *   pb_proto_fwd     out[n][2] = sum over present classes and voxels of (d^2, |d|), d = cos(fs,Ps_i) - cos(ft,Pt_i) (:161-178)
 *   pb_proto_bwd1    dfs (direct term) and dPs[n][4][8] given coef[n] = dL/d(sum d^2)
 *   pb_proto_bwd2    dfs[v] += dproto[n][t_v]  (gradient through the masked class means)
 */
int pb_softmax4(int dtype, const void* logits, float* probs, long long rows, float inv_temp, pb_stream_t stream);
int pb_softmax4_bwd(int dtype, const float* probs, const float* dprobs, void* dlogits, long long rows,
                    float inv_temp, pb_stream_t stream);
/* Fused logit-level loss pass (the north_star's "one pass over the logits"): logits [passes*b][voxels][4] (sample = pass*b + sample),
 * labels uint8 [b][voxels]; no probability tensor is materialised.
 *   mode 0: pass 0 -> softmax (T = 1) -> ce_sums[b][12] (A, L, E as pb_cedice_fwd) and, if probs0 != NULL, the probabilities
 *           [b][voxels][4] fp32 (Model.forward's first output); passes 1.. -> kl_sums[(pass-1)*b + sample] =
 *           sum_{v,c} pt (log pt - log ps), pt / ps = clamp(softmax(logit / T), .005, 1) of pass 0 / pass p (rfnet.py:284-377).
 *   mode 1: every pass -> ce_sums[pass*b + sample][12] (the four decoder_sep predictions).
 * pb_logit_loss_bwd writes d/d logits [passes*b][voxels][4] (dtype of the logits) from ce_coef (same shape as ce_sums, fp32),
 * kl_coef and, optionally, an incoming gradient dprobs0 of the returned probabilities.  (passes, mode) in {(5,0), (1,0), (4,1)}. */
int pb_logit_loss_fwd(int dtype, const void* logits, const uint8_t* labels, float* probs0, double* ce_sums, double* kl_sums,
                      int passes, int b, long long voxels, int mode, float inv_temp, pb_stream_t stream);
int pb_logit_loss_bwd(int dtype, const void* logits, const uint8_t* labels, const float* ce_coef, const float* kl_coef,
                      const float* dprobs0, void* dlogits, int passes, int b, long long voxels, int mode, float inv_temp,
                      pb_stream_t stream);
int pb_cedice_fwd(const float* probs, const uint8_t* labels, double* sums, int n, int b, long long voxels,
                  pb_stream_t stream);
int pb_cedice_bwd(const float* probs, const uint8_t* labels, const float* coef, float* dprobs, int n, int b,
                  long long voxels, pb_stream_t stream);
int pb_kl_fwd(const float* ps, const float* pt, double* sums, int n, int b, long long voxels, pb_stream_t stream);
int pb_kl_bwd(const float* ps, const float* pt, const float* coef, float* dps, int n, int b, long long voxels,
              pb_stream_t stream);
int pb_proto_sums(int dtype, const void* f, const uint8_t* labels, double* P, int n, int b, long long voxels, int c,
                  pb_stream_t stream);
int pb_proto_fwd(int dtype, const void* fs, const void* ft, const float* protos, const float* protot,
                 const float* present, double* out, int n, int b, long long voxels, int c, float eps,
                 pb_stream_t stream);
int pb_proto_bwd1(int dtype, const void* fs, const void* ft, const float* protos, const float* protot,
                  const float* present, const float* coef, void* dfs, double* dprotos, int n, int b,
                  long long voxels, int c, float eps, pb_stream_t stream);
int pb_proto_bwd2(int dtype, const uint8_t* labels, const float* dproto, void* dfs, int n, int b, long long voxels,
                  int c, pb_stream_t stream);

/* ---- training-sample pipeline on the device, csrc/augment.cu ------------------------------
 * Replaces, for one batch and in ONE launch, what the reference's DataLoader workers do per item on the CPU with
 * numpy / scipy: options.py:50's Compose([RandCrop3D, RandomRotion(10), RandomIntensityChange, RandomFlip, NumpyType])
 * (data/transforms.py:407-418, 86-120, 133-155, 217-240, 378-390) and the transposes / one-hot of
 * data/datasets_nii.py:141-160.  The random DRAWS stay on the host, in the reference's order (passion_b200/data.py);
 * the kernel applies them bit-exactly:
 *   out position (i,j,k) -> undo the flips -> p;  rotation (scipy.ndimage.rotate, order 0, mode 'constant',
 *   cval -1, reshape False): in = M (p[a0], p[a1]) + off in float64 with separate multiplies and adds in scipy's order;
 *   a coordinate outside [0, n-1] yields -1 for the image and 0 for the uint8 label, else the voxel floor(in + 0.5)
 *   of the crop at `start`;  x = float32(float64(v) * scale[p0][c] + shift[p0][c]).
 * vol: [H][W][Z][4] float32 (preprocessing/preprocess_brats.py:71-83), seg: [H][W][Z] uint8; both DEVICE pointers,
 * the struct array and the factor tables are DEVICE memory too.
 * Outputs: x [b][4][s0][s1][s2] float32 (Model.forward's input), labels [b][s0][s1][s2] uint8 (may be NULL),
 * onehot [b][4][s0][s1][s2] float64 (the reference's target format, may be NULL).
 */
typedef struct pb_augment_sample {
    const float*   vol;
    const uint8_t* seg;
    int32_t shape[3];              /* H, W, Z of the volume                                   */
    int32_t start[3];              /* crop origin (RandCrop3D.buffer)                         */
    int32_t flip[3];               /* RandomFlip x / y / z buffers                            */
    int32_t rot_axes[2];           /* sorted rotation plane a0 < a1, axes of the crop         */
    int32_t _pad;
    double  rot_m[4];              /* [[c, s], [-s, c]], c/s = cosdg/sindg(angle)             */
    double  rot_off[2];            /* centre - M centre, centre = (n - 1) / 2                 */
} pb_augment_sample;
int pb_augment_sample_size(void);  /* sizeof(pb_augment_sample), for the host-side packer's layout check */
int pb_augment_batch(const pb_augment_sample* samples, const double* scale, const double* shift, int b, int s0, int s1,
                     int s2, float* x, uint8_t* labels, double* onehot, pb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PASSION_B200_H */
