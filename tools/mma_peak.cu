// Micro-benchmark: legacy mma.sync throughput on sm_100a (TF32 m16n8k8, BF16 m16n8k16) and ldmatrix rate.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) ldm(float* out, int iters) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += 256) sm[i] = i;
    __syncthreads();
    uint32_t acc = 0;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(&sm[warp * 640 + (lane & 7) * 20 + (lane >> 3) * 160]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t r0, r1, r2, r3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr + i * 16));
            acc += r0 ^ r1 ^ r2 ^ r3;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = 148 * bps;
        if (mode == 0) k<0><<<grid, 256>>>(out, 100); else k<1><<<grid, 256>>>(out, 100);
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, 256>>>(out, iters); else k<1><<<grid, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double macs = (double)grid * 8 * iters * 8 * (mode == 0 ? 16 * 8 * 8 : 16 * 8 * 16);
        printf("%s blocks/SM=%d: %.1f TFLOP/s  (%.3f ms)\n", mode == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16", bps, 2 * macs / ms / 1e9, ms);
    }
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = 148 * bps;
        ldm<<<grid, 256>>>(out, 100);
        cudaEventRecord(e0);
        ldm<<<grid, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double bytes = (double)grid * 8 * iters * 8 * 512;
        printf("ldmatrix.x4 blocks/SM=%d: %.1f TB/s smem (%.1f B/clk/SM at 1.965GHz)\n", bps, bytes / ms / 1e9, bytes / ms / 1e6 / 148 / 1965e3 * 1e3);
    }
    return 0;
}
