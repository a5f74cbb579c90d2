// Shared device helpers for the passion_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/passion_b200.h"

typedef __nv_bfloat16 bf16;

void pb_set_error(const char* fmt, ...);
void pb_count_launch(int n = 1);
// csrc/conv3d_wgrad_rs.cu: row-stacked tcgen05 weight gradient (PB_EUNSUPPORTED = class not covered)
int pb_wgrad_rs_launch(const pb_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw, int* err_flag, cudaStream_t st);

#define PB_CHECK_ARG(cond, msg)                                   \
    do { if (!(cond)) { pb_set_error("%s: %s", __func__, msg); return PB_EINVAL; } } while (0)

#define PB_CHECK_LAUNCH()                                                             \
    do { cudaError_t e_ = cudaGetLastError();                                         \
         if (e_ != cudaSuccess) { pb_set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e_)); \
                                  return PB_ECUDA; }                                  \
         pb_count_launch(); } while (0)

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- vector load/store of N contiguous channels (pointer aligned to N*sizeof(T)) ----
template <typename T, int N> struct VecIO;

template <int N> struct VecIO<float, N> {
    static __device__ __forceinline__ void load(const float* p, float* o) {
        if constexpr (N % 4 == 0) {
#pragma unroll
            for (int i = 0; i < N / 4; ++i) {
                float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
                o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w;
            }
        } else if constexpr (N == 2) {
            float2 v = __ldg(reinterpret_cast<const float2*>(p)); o[0] = v.x; o[1] = v.y;
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) o[i] = __ldg(p + i);
        }
    }
    static __device__ __forceinline__ void store(float* p, const float* o) {
        if constexpr (N % 4 == 0) {
#pragma unroll
            for (int i = 0; i < N / 4; ++i)
                reinterpret_cast<float4*>(p)[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        } else if constexpr (N == 2) {
            *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) p[i] = o[i];
        }
    }
};

__device__ __forceinline__ void bf2_unpack(uint32_t u, float& a, float& b) {
    a = __uint_as_float(u << 16);
    b = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ uint32_t bf2_pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int N> struct VecIO<bf16, N> {
    static __device__ __forceinline__ void load(const bf16* p, float* o) {
        if constexpr (N % 8 == 0) {
#pragma unroll
            for (int i = 0; i < N / 8; ++i) {
                uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + i);
                bf2_unpack(v.x, o[8 * i + 0], o[8 * i + 1]); bf2_unpack(v.y, o[8 * i + 2], o[8 * i + 3]);
                bf2_unpack(v.z, o[8 * i + 4], o[8 * i + 5]); bf2_unpack(v.w, o[8 * i + 6], o[8 * i + 7]);
            }
        } else if constexpr (N == 4) {
            uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
            bf2_unpack(v.x, o[0], o[1]); bf2_unpack(v.y, o[2], o[3]);
        } else if constexpr (N == 2) {
            uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
            bf2_unpack(v, o[0], o[1]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) o[i] = __bfloat162float(p[i]);
        }
    }
    static __device__ __forceinline__ void store(bf16* p, const float* o) {
        if constexpr (N % 8 == 0) {
#pragma unroll
            for (int i = 0; i < N / 8; ++i)
                reinterpret_cast<uint4*>(p)[i] = make_uint4(bf2_pack(o[8 * i], o[8 * i + 1]), bf2_pack(o[8 * i + 2], o[8 * i + 3]),
                                                            bf2_pack(o[8 * i + 4], o[8 * i + 5]), bf2_pack(o[8 * i + 6], o[8 * i + 7]));
        } else if constexpr (N == 4) {
            *reinterpret_cast<uint2*>(p) = make_uint2(bf2_pack(o[0], o[1]), bf2_pack(o[2], o[3]));
        } else if constexpr (N == 2) {
            *reinterpret_cast<uint32_t*>(p) = bf2_pack(o[0], o[1]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) p[i] = __float2bfloat16_rn(o[i]);
        }
    }
};

// A VEC-channel vector as it sits in memory: loads of several vectors are issued back to back in their packed form (one
// register per two bf16 values) and unpacked to fp32 only when consumed, which doubles the bytes a thread keeps in flight
// for the same register budget.
template <typename T, int VEC> struct RawVec {
    float v[VEC];
    __device__ __forceinline__ void load(const T* p) { VecIO<T, VEC>::load(p, v); }
    __device__ __forceinline__ void unpack(float* o) const {
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = v[j];
    }
};
template <> struct RawVec<bf16, 8> {
    uint4 r;
    __device__ __forceinline__ void load(const bf16* p) { r = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void unpack(float* o) const {
        bf2_unpack(r.x, o[0], o[1]); bf2_unpack(r.y, o[2], o[3]); bf2_unpack(r.z, o[4], o[5]); bf2_unpack(r.w, o[6], o[7]);
    }
};
template <> struct RawVec<bf16, 4> {
    uint2 r;
    __device__ __forceinline__ void load(const bf16* p) { r = __ldg(reinterpret_cast<const uint2*>(p)); }
    __device__ __forceinline__ void unpack(float* o) const { bf2_unpack(r.x, o[0], o[1]); bf2_unpack(r.y, o[2], o[3]); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 32 uniform bits per (seed, element counter) for the attention dropout masks (csrc/attn.cu, csrc/gemm_tc.cu): a 32-bit multiply /
// xor-shift mixer — the mask has to be reproducible within one forward pass (the backward reads the saved P and P', it never regenerates
// the mask) and independent between steps (a fresh device-side seed per call), not cryptographic.
__device__ __forceinline__ uint32_t pb_dropout_bits(unsigned long long seed, unsigned long long idx) {
    uint32_t x = ((uint32_t)idx ^ (uint32_t)seed) + ((uint32_t)(idx >> 32) + (uint32_t)(seed >> 32)) * 0x9E3779B9u;
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

// largest of {8,4,2,1} dividing c
static inline int pb_vec_width(int c) { return (c % 8 == 0) ? 8 : (c % 4 == 0) ? 4 : (c % 2 == 0) ? 2 : 1; }

// reflect index into [0, n) for one-voxel overhang (PyTorch 'reflect': -1 -> 1, n -> n-2)
__device__ __forceinline__ int reflect_idx(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}
