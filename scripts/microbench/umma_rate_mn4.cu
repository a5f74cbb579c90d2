// Micro-benchmark: how many cycles does one tcgen05.mma (kind::f16, bf16 -> fp32, cta_group::1, K = 16) cost as a function of
// M, N, the shared-memory layout of the operands and the accumulator dependency pattern?  One CTA per SM; one thread issues
// `iters` MMAs back to back, commits to an mbarrier and waits.  Operand contents are irrelevant (shared memory is zeroed).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate scripts/microbench/umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ULL << 46) | ((uint64_t)layout_type << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Cfg { int M, N, layout, nacc, iters, a_shift_rows, issuers, a_sbo, b_sbo, loopy, cols, same, commit_every, a_layout; };   // layout 9: both operands MN-major (LBO = 128 B per 8 K rows, SBO = MN-group stride)   // layout: 0 no-swizzle K-major planes, 2 SW128, 4 SW64, 6 SW32

__global__ void __launch_bounds__(128, 2) rate_kernel(Cfg c, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)c.cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && wid < c.issuers) {
        uint64_t& bar = bars[wid];
        const uint32_t tmem = tmem_base + (uint32_t)((c.same ? 0 : wid) * 2 * c.N);      // own accumulators per issuing warp (or all the same)
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
        uint64_t a, b;
        if (c.layout == 9) {
            a = umma_desc(a_addr, 128, c.a_sbo, 0);
            b = umma_desc(b_addr, 128, c.b_sbo, 0);
        } else if (c.layout == 0) {            // [K chunk of 8][row][16 B]: SBO = 128 B (8-row groups), LBO = chunk-plane pitch
            a = umma_desc(a_addr, 520 * 16, 128, 0);
            b = umma_desc(b_addr, c.N * 16, 128, 0);
        } else {                        // canonical swizzled K-major atoms: 8 rows x (32|64|128) B, SBO = 8 rows
            const uint32_t row = c.layout == 2 ? 128 : (c.layout == 4 ? 64 : 32);
            a = umma_desc(a_addr, 16, 8 * row, c.layout);
            b = umma_desc(b_addr, 16, 8 * row, c.layout);
        }
        const uint32_t idesc = umma_idesc(c.M, c.N, c.layout == 9, c.layout == 9);
        uint64_t ai[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) ai[t] = a + (uint64_t)(uint32_t)(t * c.a_shift_rows);   // tap-like start-address shifts
        uint64_t bi[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) bi[t] = b + (uint64_t)(uint32_t)(c.layout == 9 ? t * c.a_shift_rows : 0);
        const uint32_t d1 = tmem + (uint32_t)((c.nacc > 1 ? 1 : 0) * c.N);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            // lean issue loop: 9 MMAs per iteration, all operands already in registers, alternating accumulators if nacc = 2
            if (c.loopy == 2) {      // K-steps unrolled, descriptors advanced by adding to the low word only
                for (int i = 0; i < c.iters; i += 40)
                    for (int r = 0; r < 8; ++r) {
                        const uint32_t uc = c.a_layout == 6 ? 2 : (c.a_layout == 4 ? 4 : 1);
                        const uint64_t a0 = c.a_layout == 0 ? umma_desc(a_addr + r * 1408, 128, c.a_sbo, 0) : umma_desc(a_addr + r * 1408 * uc, 16 * uc, 128 * uc, c.a_layout);
                        const uint64_t b0 = umma_desc(b_addr + (r % 3) * 3 * c.b_sbo, 128, c.b_sbo, 0);
#pragma unroll
                        for (int ks = 0; ks < 5; ++ks) umma_f16(tmem, a0 + (uint64_t)(16 * uc * ks), b0 + (uint64_t)(16 * ks), idesc, 1u);
                        if (c.commit_every && ((r + 1) % c.commit_every) == 0) { umma_commit(&bars[3]); umma_commit(&bars[3]); }
                    }
            } else if (c.loopy) {      // the production loop shape: rows x K-steps with the descriptors advanced in the loop
                for (int i = 0; i < c.iters; i += 40)
                    for (int r = 0; r < 8; ++r) {
                        const uint64_t a0 = umma_desc(a_addr + r * 1408, 128, c.a_sbo, 0), b0 = umma_desc(b_addr + (r % 3) * 3 * c.b_sbo, 128, c.b_sbo, 0);
                        for (int ks = 0; ks < c.a_shift_rows / 16 * 5; ++ks) umma_f16(tmem, a0 + (uint64_t)(16 * ks), b0 + (uint64_t)(16 * ks), idesc, 1u);
                    }
            } else
            for (int i = 0; i < c.iters; i += 9) {
#pragma unroll
                for (int t = 0; t < 9; ++t) umma_f16((t & 1) ? d1 : tmem, ai[t], bi[t], idesc, 1u);
            }
            umma_commit(&bar);
            while (!mbar_try_wait(&bar, rep & 1)) {}
            const long long t1 = clock64();
            if (rep == 2 && wid == 0) cycles[blockIdx.x] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)c.cols) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 296 * sizeof(long long));
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    printf("both operands MN-major, K-step shifts of 16 rows between consecutive MMAs\n");
    printf("M N nacc issuers a_sbo b_sbo | cycles per MMA (aggregate) | MAC/clk/SM | %% of 4096\n");
    const int cfgs[][11] = {  // M, N, nacc, issuers, a_sbo, b_sbo, loopy, CTAs per SM, same accumulator, commits, A layout (0 none, 6 SW32, 4 SW64)
        {64, 88, 1, 3, 16, 1280, 2, 1, 1, 0, 0}, {64, 88, 1, 3, 16, 1280, 2, 1, 1, 0, 6}, {128, 88 + 8, 1, 3, 16, 1280, 2, 1, 1, 0, 4}, {128, 96, 1, 3, 16, 1280, 2, 1, 1, 0, 0},
        {64, 88, 1, 1, 16, 1280, 2, 1, 1, 0, 6}, {128, 96, 1, 1, 16, 1280, 2, 1, 1, 0, 4}, {64, 72, 1, 3, 16, 1280, 2, 1, 1, 0, 6},
    };
    for (auto& q : cfgs) {
        Cfg c{q[0], q[1], 9, q[2], 2040, 16, q[3], q[4], q[5], q[6], q[7] == 2 ? 256 : 512, q[8], q[9], q[10]};
        if ((c.same ? 1 : c.issuers) * 2 * c.N > c.cols) { printf("skip\n"); continue; }
        rate_kernel<<<148 * q[7], 128, 96 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M%d N%d: %s\n", c.M, c.N, cudaGetErrorString(e)); return 1; }
        long long h[296];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += (double)h[i];
        avg /= 148.0 * c.iters * c.issuers;
        const double mac = (double)c.M * c.N * 16 / avg;
        printf("A layout %d | ", q[10]);
        printf("%3d %3d %d %d %4d %4d | %7.1f | %7.0f | %5.1f\n", c.M, c.N, c.nacc, c.issuers, c.a_sbo, c.b_sbo, avg, mac, 100.0 * mac / 4096.0);
    }
    return 0;
}
