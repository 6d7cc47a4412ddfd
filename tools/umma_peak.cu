// tcgen05.mma kind::tf32 issue-rate / throughput probe (B200): one CTA per SM, one thread issues a tight unrolled
// stream of M=128 x N x K=8 MMAs on operands resident in shared memory (no global traffic), K-major (128B swizzle)
// or MN-major (128B swizzle, 32B atoms) descriptors.  Prints cycles per MMA and chip TFLOP/s per variant.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_peak tools/umma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: K-major SW128 both; 1: MN-major both; 2: MN-major with overlapping 128-byte column groups (the wgrad trick)
// 3: bf16 K-major (kind::f16, K = 16) for reference
template <int N, int MODE, int UNROLL>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* cycles_out) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (uint32_t o = threadIdx.x * 16; o < 160 * 1024; o += 128 * 16)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem0 + o), "r"(0u) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    if (threadIdx.x == 0) {
        const int a_mn = (MODE == 1 || MODE == 2), b_mn = a_mn;
        const uint32_t fmt = MODE == 3 ? 1u : 2u;      // bf16 : tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        // four operand stages of 32 KB (A 16 KB + B 16..32 KB), descriptors advance by 1024 B per k-step within a stage
        uint64_t ad[4], bd[4];
        for (int s = 0; s < 4; ++s) {
            const uint32_t a = smem0 + s * 40960, b = a + 16384;
            if (MODE == 0 || MODE == 3) { ad[s] = smem_desc(a, 16, 1024, 2); bd[s] = smem_desc(b, 16, 1024, 2); }
            else if (MODE == 1) { ad[s] = smem_desc(a, 4096, 512, 1); bd[s] = smem_desc(b, 4096, 512, 1); }
            else { ad[s] = smem_desc(a, 4096, 512, 1); bd[s] = smem_desc(b, 128, 512, 1); }
        }
        const long long t0 = clock64();
        for (int it = 0; it < iters; it += UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int s = u & 3;
                // K-major: k-step = +32 B (2 in 16-byte units); MN-major: +1024 B (64)
                const uint64_t step = (MODE == 0 || MODE == 3) ? (uint64_t)(2 * ((u >> 2) & 3)) : (uint64_t)(64 * ((u >> 2) & 3));
                if (MODE == 3) mma_f16(tmem, ad[s] + step, bd[s] + step, idesc, 1u);
                else mma_tf32(tmem, ad[s] + step, bd[s] + step, idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        const long long t1 = clock64();
        cycles_out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, int MODE, int UNROLL>
void run(const char* name, int sms, long long* dcyc) {
    auto k = probe<N, MODE, UNROLL>;
    const int smem = 170 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 16384;
    k<<<sms, 128, smem>>>(256, dcyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<sms, 128, smem>>>(iters, dcyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s N=%d: %s\n", name, N, cudaGetErrorString(err)); return; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, dcyc, sizeof(long long) * (sms < 148 ? sms : 148), cudaMemcpyDeviceToHost);
    const int kk = MODE == 3 ? 16 : 8;
    const double flops = 2.0 * 128 * N * kk * (double)iters * sms;
    printf("%-28s N=%3d unroll=%2d: %7.1f cycles/MMA (ideal %3d)  %8.1f TFLOP/s  (%.3f ms)\n", name, N, UNROLL, (double)h[0] / iters, N / 2,
           flops / (ms * 1e-3) / 1e12, ms);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* dcyc;
    cudaMalloc(&dcyc, sizeof(long long) * 256);
    run<32, 0, 16>("tf32 K-major SW128", sms, dcyc);
    run<64, 0, 16>("tf32 K-major SW128", sms, dcyc);
    run<96, 0, 16>("tf32 K-major SW128", sms, dcyc);
    run<128, 0, 16>("tf32 K-major SW128", sms, dcyc);
    run<256, 0, 16>("tf32 K-major SW128", sms, dcyc);
    run<32, 1, 16>("tf32 MN-major 32B atoms", sms, dcyc);
    run<64, 1, 16>("tf32 MN-major 32B atoms", sms, dcyc);
    run<96, 1, 16>("tf32 MN-major 32B atoms", sms, dcyc);
    run<128, 1, 16>("tf32 MN-major 32B atoms", sms, dcyc);
    run<256, 1, 16>("tf32 MN-major 32B atoms", sms, dcyc);
    run<96, 2, 16>("tf32 MN-major B LBO=128", sms, dcyc);
    run<96, 2, 1>("tf32 MN-major B LBO=128", sms, dcyc);
    run<32, 1, 1>("tf32 MN-major 32B atoms", sms, dcyc);
    run<32, 3, 16>("bf16 K-major SW128", sms, dcyc);
    run<128, 3, 16>("bf16 K-major SW128", sms, dcyc);
    run<256, 3, 16>("bf16 K-major SW128", sms, dcyc);
    return 0;
}
