// Device-side training-sample pipeline: random crop -> nearest-neighbour rotation -> per-row intensity
// change -> flips -> [B,4,S,S,S] float32 image, uint8 label map and (optionally) the reference's float64 one-hot target, one
// launch per batch.  Bit-exact against the reference's numpy / scipy transforms (data/transforms.py:86-120, 133-155,
// 217-240, 407-418; data/datasets_nii.py:141-160): all coordinate and intensity arithmetic is float64 with explicitly
// separate multiplies and adds (__dmul_rn / __dadd_rn — a fused multiply-add would round differently from scipy's C loop).
// HBM-bound gather: 20 B read and 16 B (+ 1 B label, + 32 B one-hot) written per output voxel; the volumes stay resident in
// HBM (a preprocessed BraTS case is 143 MB; the 219 training cases are 31 GB of the 180 GB).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) augment_kernel(const pb_augment_sample* __restrict__ samples,
                                                      const double* __restrict__ scale, const double* __restrict__ shift,
                                                      int B, int s0, int s1, int s2, float* __restrict__ x,
                                                      uint8_t* __restrict__ labels, double* __restrict__ onehot) {
    const long long V = (long long)s0 * s1 * s2;
    const long long total = (long long)B * V;
    for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const int b = (int)(t / V);
        long long r = t - (long long)b * V;
        const int k = (int)(r % s2); r /= s2;
        const int j = (int)(r % s1);
        const int i = (int)(r / s1);
        const pb_augment_sample& sp = samples[b];
        // RandomFlip is the last geometric op: output (i,j,k) reads the pre-flip position p
        int p[3];
        p[0] = sp.flip[0] ? s0 - 1 - i : i;
        p[1] = sp.flip[1] ? s1 - 1 - j : j;
        p[2] = sp.flip[2] ? s2 - 1 - k : k;
        // scipy.ndimage.rotate(order=0, mode='constant', cval=-1, reshape=False) in the plane (a0, a1) of the crop
        const int a0 = sp.rot_axes[0], a1 = sp.rot_axes[1];
        const int n[3] = {s0, s1, s2};
        const double o0 = (double)p[a0], o1 = (double)p[a1];
        const double c0 = __dadd_rn(__dadd_rn(__dmul_rn(o0, sp.rot_m[0]), __dmul_rn(o1, sp.rot_m[1])), sp.rot_off[0]);
        const double c1 = __dadd_rn(__dadd_rn(__dmul_rn(o0, sp.rot_m[2]), __dmul_rn(o1, sp.rot_m[3])), sp.rot_off[1]);
        const bool inb = c0 >= 0.0 && c0 <= (double)(n[a0] - 1) && c1 >= 0.0 && c1 <= (double)(n[a1] - 1);
        float v[4] = {-1.f, -1.f, -1.f, -1.f};
        int lab = 0;                                                   // cval -1 clamps to 0 in an unsigned output
        if (inb) {
            int q[3] = {p[0], p[1], p[2]};
            q[a0] = (int)floor(__dadd_rn(c0, 0.5));
            q[a1] = (int)floor(__dadd_rn(c1, 0.5));
            const size_t src = ((size_t)(sp.start[0] + q[0]) * sp.shape[1] + (sp.start[1] + q[1])) * sp.shape[2] + (sp.start[2] + q[2]);
            const float4 f = __ldg(reinterpret_cast<const float4*>(sp.vol) + src);
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            lab = sp.seg ? (int)__ldg(sp.seg + src) : 0;
        }
        // RandomIntensityChange: factors indexed by (row of the rotated crop, channel), float32 * float64 + float64 -> float32
        const double* sc = scale + ((size_t)b * s0 + p[0]) * 4;
        const double* sh = shift + ((size_t)b * s0 + p[0]) * 4;
        const size_t o = (size_t)b * 4 * V + ((size_t)i * s1 + j) * s2 + k;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            x[o + (size_t)c * V] = (float)__dadd_rn(__dmul_rn((double)v[c], __ldg(sc + c)), __ldg(sh + c));
        if (labels) labels[(size_t)b * V + ((size_t)i * s1 + j) * s2 + k] = (uint8_t)lab;
        if (onehot) {
#pragma unroll
            for (int c = 0; c < 4; ++c) onehot[o + (size_t)c * V] = lab == c ? 1.0 : 0.0;
        }
    }
}

}  // namespace

extern "C" int pb_augment_sample_size(void) { return (int)sizeof(pb_augment_sample); }

extern "C" int pb_augment_batch(const pb_augment_sample* samples, const double* scale, const double* shift, int b, int s0,
                                int s1, int s2, float* x, uint8_t* labels, double* onehot, pb_stream_t stream) {
    PB_CHECK_ARG(samples && scale && shift && x && b >= 1 && s0 >= 1 && s1 >= 1 && s2 >= 1, "bad arguments");
    const long long total = (long long)b * s0 * s1 * s2;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    augment_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(samples, scale, shift, b, s0, s1, s2, x, labels, onehot);
    PB_CHECK_LAUNCH();
    return PB_OK;
}
