// TMA streaming-rate probe (B200): every CTA streams its contiguous share of a large fp32 buffer into a ring of R
// shared-memory slots with 2-D tiled TMA boxes of `rows` x 128 bytes (128-byte swizzle); a consumer thread only waits for
// "full" and releases the slot.  Reports GB/s per (ring depth, box rows): how many bytes must be in flight per SM to
// saturate HBM with TMA loads, and what one box row costs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream tools/tma_stream.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

__global__ void __launch_bounds__(64, 1) stream5d_kernel(const __grid_constant__ CUtensorMap tm, int R, int boxes_per_cta, int H) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t slot_bytes = 136u * 128u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < R; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int row0 = blockIdx.x * boxes_per_cta;
    if (threadIdx.x == 0) {
        for (int i = 0; i < boxes_per_cta; ++i) {
            const int s = i % R;
            if (i >= R) mbar_wait(smem_u32(&empty[s]), ((i / R) - 1) & 1);
            const uint32_t fb = smem_u32(&full[s]);
            const int r = row0 + i, n = r / H, h = r - n * H;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(slot_bytes), "r"(fb) : "memory");
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                         ::"r"(smem0 + s * slot_bytes), "l"(&tm), "r"(fb), "r"(0), "r"(-1), "r"(0), "r"(h), "r"(n) : "memory");
        }
    } else if (threadIdx.x == 32) {
        for (int i = 0; i < boxes_per_cta; ++i) {
            const int s = i % R;
            mbar_wait(smem_u32(&full[s]), (i / R) & 1);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
    }
}

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int R, int rows, int boxes_per_cta, int row_bytes) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t slot_bytes = (uint32_t)rows * (uint32_t)row_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < R; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int row0 = blockIdx.x * boxes_per_cta * rows;
    if (threadIdx.x == 0) {                       // producer
        for (int i = 0; i < boxes_per_cta; ++i) {
            const int s = i % R;
            if (i >= R) mbar_wait(smem_u32(&empty[s]), ((i / R) - 1) & 1);
            const uint32_t fb = smem_u32(&full[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(slot_bytes), "r"(fb) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem0 + s * slot_bytes), "l"(&tm), "r"(fb), "r"(0), "r"(row0 + i * rows) : "memory");
        }
    } else if (threadIdx.x == 32) {               // consumer
        for (int i = 0; i < boxes_per_cta; ++i) {
            const int s = i % R;
            mbar_wait(smem_u32(&full[s]), (i / R) & 1);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)f;
    const size_t total_rows = (size_t)sms * 8192;             // x 128 B = 148 x 1 MB
    float* buf;
    cudaMalloc(&buf, total_rows * 128 * 4);                   // 4 rounds worth so that consecutive runs do not hit L2
    cudaMemset(buf, 0, total_rows * 128 * 4);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct Var { int inner; int rows; CUtensorMapSwizzle swz; const char* name; int resident; };
    const Var vars[] = {{32, 136, CU_TENSOR_MAP_SWIZZLE_128B, "128B rows, SW128, HBM", 0},
                        {32, 136, CU_TENSOR_MAP_SWIZZLE_128B, "128B rows, SW128, L2-resident", 1},
                        {32, 136, CU_TENSOR_MAP_SWIZZLE_NONE, "128B rows, no swizzle, HBM", 0},
                        {64, 68, CU_TENSOR_MAP_SWIZZLE_NONE, "256B rows, no swizzle, HBM", 0},
                        {256, 17, CU_TENSOR_MAP_SWIZZLE_NONE, "1KB rows, no swizzle, HBM", 0},
                        {256, 17, CU_TENSOR_MAP_SWIZZLE_NONE, "1KB rows, no swizzle, L2-resident", 1}};
    for (const Var& v : vars) {
        const int row_bytes = v.inner * 4;
        const size_t rows_total = (size_t)sms * 8192 * 128 / row_bytes;          // 1 MB per CTA
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)v.inner, (cuuint64_t)rows_total * (v.resident ? 1 : 4)};
        const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
        const cuuint32_t box[2] = {(cuuint32_t)v.inner, (cuuint32_t)v.rows};
        const cuuint32_t es[2] = {1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.swz,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        for (int R = 2; R <= 10; R += 4) {
            const int per_cta = v.resident ? 64 * 1024 : 1024 * 1024;             // L2-resident: 148 x 64 KB = 9.5 MB, looped 16 times
            const int boxes = per_cta / (v.rows * row_bytes);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            const int reps = v.resident ? 16 : 1;
            stream_kernel<<<sms, 64, 210 * 1024>>>(tm, R, v.rows, boxes, row_bytes);
            if (!v.resident) cudaMemset(buf + total_rows * 32, 0, total_rows * 128 * 3);      // evict L2
            cudaEventRecord(e0);
            for (int k = 0; k < reps; ++k) stream_kernel<<<sms, 64, 210 * 1024>>>(tm, R, v.rows, boxes, row_bytes);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("%s R %d: %s\n", v.name, R, cudaGetErrorString(err)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)sms * boxes * v.rows * row_bytes * reps;
            printf("%-36s ring %2d (%6.1f KB in flight): %7.1f GB/s  %6.1f GB/s per SM  (%.1f us per launch)\n", v.name, R,
                   R * v.rows * row_bytes / 1024.0, bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e9 / sms, ms * 1e3 / reps);
        }
    }
    {   // 5-D map as the row-ring kernels use it: {32 ch, W = 128, 1 plane, H = 256, N}, box {32, 136, 1, 1, 1} at w = -1
        cudaFuncSetAttribute(stream5d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        const int H = 256, W = 128, boxes = 64;                               // 64 rows x 16 KB = 1 MB per CTA
        const size_t N = ((size_t)sms * boxes + H - 1) / H;
        CUtensorMap tm;
        const cuuint64_t dims[5] = {32, (cuuint64_t)W, 1, (cuuint64_t)H, N * 4};
        const cuuint64_t strides[4] = {128, 128, (cuuint64_t)W * 128, (cuuint64_t)W * 128 * H};
        const cuuint32_t box[5] = {32, 136, 1, 1, 1};
        const cuuint32_t es[5] = {1, 1, 1, 1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode 5d failed\n"); return 1; }
        for (int R = 2; R <= 10; R += 4) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            stream5d_kernel<<<sms, 64, 210 * 1024>>>(tm, R, boxes, H);
            cudaMemset(buf + total_rows * 32, 0, total_rows * 128 * 3);
            cudaEventRecord(e0);
            stream5d_kernel<<<sms, 64, 210 * 1024>>>(tm, R, boxes, H);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("5d R %d: %s\n", R, cudaGetErrorString(err)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)sms * boxes * 128 * 128;
            printf("%-36s ring %2d (%6.1f KB in flight): %7.1f GB/s  %6.1f GB/s per SM  (%.1f us per launch)\n", "5-D map, 136-pixel row boxes, HBM", R,
                   R * 17.0, bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e9 / sms, ms * 1e3);
        }
    }
    return 0;
}
