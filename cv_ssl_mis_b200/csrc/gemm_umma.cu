// Token-matrix GEMMs (nn.Linear forward / data gradient / weight gradient, 1x1 convolutions) on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), accumulators
// in TMEM.  One persistent warp-specialised kernel serves the three products of a Linear layer
//     forward  y  = [x0|x1] W^T + b      A = x  (K-major),   B = W  (K-major)
//     dgrad    dx = dy W                 A = dy (K-major),   B = W  (MN-major: the reduction index o is W's slow dim)
//     wgrad    dW = dy^T [x0|x1]         A = dy (MN-major),  B = x  (MN-major), split over token ranges
// straight from the row-major fp32 tensors the reference holds (swin_transformer_unet_skip_expand_decoder_sys.py:19-25,
// 115-150, 336-346, 378, 405): no weight packing, no transposed copies -- the MN-major cases use the transposing
// operand fetch of the UMMA descriptor (instruction-descriptor bits 15/16).
//
//   warp 0     TMA producer (one lane): per 32-deep k-block ONE box of A and ONE box of B -- 2-D {32, rows} for K-major
//              operands, 3-D {32 columns, 32 k-rows, column groups} for MN-major ones (per-32-column boxes only as the
//              fallback for ragged / concatenated operands) -- into a ring of shared-memory stages, mbarrier expect_tx
//   warp 1     MMA issuer (one lane): 4 x tcgen05.mma (K = 8) per k-block, tcgen05.commit frees the stage;
//              owns the TMEM allocation (two accumulator buffers of BN columns)
//   warps 2-9  epilogue: tcgen05.ld 32x32b.x32 -> + bias -> swizzled shared-memory tile -> TMA store (or fp32 reduce-add
//              for accumulate) through a {column, row, split} tensor map of the destination
#include "common.cuh"
#include <cuda.h>
#include <cstring>
#include "../../include/b200ssl.h"

namespace {

constexpr int BK = 32;                   // reduction elements per stage (= one 128-byte swizzle row of fp32)
constexpr int BM = 128;
constexpr int A_BYTES = BM * BK * 4;     // 16 KB
constexpr int G_THREADS = 320;           // TMA warp + MMA warp + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int G_MAX_STAGES = 8;

struct GemmP {
    int M, N, K;                 // output rows, output columns, reduction length
    int BN;                      // tile columns (multiple of 32, <= 256)
    int a_mn, b_mn;              // operand is MN-major in memory
    int a_3d, b_3d;              // MN-major operand fetched with ONE 3-D box {32 columns, BK rows, column groups} per stage
    int a_k0;                    // reduction elements served by tensor map a0 (the rest by a1)   [A K-major only]
    int b_n0;                    // output columns served by tensor map b0 (the rest by b1)       [B MN-major only]
    int m_tiles, n_tiles, splits, kb_total;
    int nstages;
    float* dst0;
    float* dst1;
    int ld0, ld1, ncol0;         // columns [0, ncol0) -> dst0 (row stride ld0), the rest -> dst1 (ld1)
    long long split_stride;      // floats between the partial results of consecutive splits in dst0
    const float* bias;
    int accumulate;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        if (++spins > (1u << 28)) __trap();      // a lost arrival must fail loudly, not hang the device
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared -> global tile store / reduce-add (fp32) through the tensor map {column, row, split}; clipped at the matrix edge
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int add) {
    if (add)
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// shared-memory matrix descriptor (sm_100 version bit).  layout 2 = 128-byte swizzle of 16-byte chunks (K-major
// operands); layout 1 = 128-byte swizzle of 32-byte chunks, the only swizzled layout tcgen05 accepts for MN-major
// 32-bit (TF32) operands -- TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__global__ void __launch_bounds__(G_THREADS) gemm_umma_kernel(const __grid_constant__ CUtensorMap ta0,
                                                              const __grid_constant__ CUtensorMap ta1,
                                                              const __grid_constant__ CUtensorMap tb0,
                                                              const __grid_constant__ CUtensorMap tb1,
                                                              const __grid_constant__ CUtensorMap tc0,
                                                              const __grid_constant__ CUtensorMap tc1, const GemmP p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[G_MAX_STAGES], empty_bar[G_MAX_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (role branches on the uniform datapath)
    const int S = p.nstages, BN = p.BN;
    const int stage_bytes = A_BYTES + BN * BK * 4;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;      // the 128-byte swizzle wants 1024-byte alignment
    const int nitems = p.m_tiles * p.n_tiles * p.splits;
    int tmem_cols = 32;
    while (tmem_cols < 2 * BN) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&acc_full[a]), 1);
            mbar_init(smem_u32(&acc_empty[a]), EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        // whole warp in warp-uniform control flow, one elected lane issues (see the MMA issuer below)
        {
            uint32_t leader_u;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(leader_u));
            const bool leader = leader_u != 0;
            int s = 0, eph = 0, it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int nt = item % p.n_tiles, mt = (item / p.n_tiles) % p.m_tiles, sp = item / (p.n_tiles * p.m_tiles);
                const int kb0 = (int)((long long)sp * p.kb_total / p.splits), kb1 = (int)((long long)(sp + 1) * p.kb_total / p.splits);
                const int m0 = mt * BM, n0 = nt * BN;
                // boxes entirely outside the matrix are skipped (their rows / columns are never stored)
                const int a_boxes = p.a_mn ? min(4, (p.M - m0 + 31) / 32) : 1;
                const int b_boxes = p.b_mn ? min(BN / 32, (p.N - n0 + 31) / 32) : 1;
                // a 3-D box always delivers its full size (out-of-range column groups are zero-filled)
                const uint32_t tx = ((p.a_mn && !p.a_3d) ? a_boxes * 4096u : (uint32_t)A_BYTES) +
                                    ((p.b_mn && !p.b_3d) ? b_boxes * 4096u : (uint32_t)(BN * BK * 4));
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    if (it >= S) mbar_wait(smem_u32(&empty_bar[s]), eph ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    const uint32_t a_dst = smem0 + s * stage_bytes, b_dst = a_dst + A_BYTES;
                    const int k = kb * BK;
                    if (leader) {
                        mbar_expect_tx(fb, tx);
                        if (!p.a_mn) {
                            if (k < p.a_k0) tma_load_2d(a_dst, &ta0, k, m0, fb);
                            else tma_load_2d(a_dst, &ta1, k - p.a_k0, m0, fb);
                        } else if (p.a_3d) {
                            tma_load_3d(a_dst, &ta0, 0, k, m0 >> 5, fb);
                        } else {
                            for (int j = 0; j < a_boxes; ++j) tma_load_2d(a_dst + j * 4096, &ta0, m0 + 32 * j, k, fb);
                        }
                        if (!p.b_mn) {
                            tma_load_2d(b_dst, &tb0, k, n0, fb);
                        } else if (p.b_3d) {
                            tma_load_3d(b_dst, &tb0, 0, k, n0 >> 5, fb);
                        } else {
                            for (int j = 0; j < b_boxes; ++j) {
                                const int col = n0 + 32 * j;
                                if (col < p.b_n0) tma_load_2d(b_dst + j * 4096, &tb0, col, k, fb);
                                else tma_load_2d(b_dst + j * 4096, &tb1, col - p.b_n0, k, fb);
                            }
                        }
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; eph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop (warp-uniform control flow keeps the descriptors in uniform registers; under
        // `if (lane == 0)` every tcgen05.mma is wrapped in a uniformisation loop) and ONE elected lane issues.  The two
        // descriptors of a stage are built once, a k-step adds 32 B (K-major) or 1024 B (MN-major) to the address field.
        {
            uint32_t leader_u;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(leader_u));
            const bool leader = leader_u != 0;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t a_step = p.a_mn ? 64 : 2, b_step = p.b_mn ? 64 : 2;       // in units of 16 B
            int tl = 0, s = 0, fph = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++tl) {
                const int sp = item / (p.n_tiles * p.m_tiles);
                const int kb0 = (int)((long long)sp * p.kb_total / p.splits), kb1 = (int)((long long)(sp + 1) * p.kb_total / p.splits);
                const int as = tl & 1;
                if (tl >= 2) mbar_wait(smem_u32(&acc_empty[as]), ((tl >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol = tmem_base + as * BN;
                uint32_t acc = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&full_bar[s]), fph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_base = smem0 + s * stage_bytes, b_base = a_base + A_BYTES;
                    // K-major: 8 rows x 128 B swizzle atoms, 1024 B apart (SBO); a k-step is 32 bytes along the row.
                    // MN-major: a 32-column box is [32 k][128 B]; 4 k-rows = one 512-byte atom (SBO), the next
                    // 32 columns are the next 4096-byte box (LBO); a k-step (8 k-rows) is 1024 bytes.
                    const uint64_t ad = p.a_mn ? smem_desc(a_base, 4096, 512, 1) : smem_desc(a_base, 16, 1024, 2);
                    const uint64_t bd = p.b_mn ? smem_desc(b_base, 4096, 512, 1) : smem_desc(b_base, 16, 1024, 2);
                    if (leader) {
                        umma_tf32(dcol, ad, bd, idesc, acc);
                        umma_tf32(dcol, ad + a_step, bd + b_step, idesc, 1u);
                        umma_tf32(dcol, ad + 2 * a_step, bd + 2 * b_step, idesc, 1u);
                        umma_tf32(dcol, ad + 3 * a_step, bd + 3 * b_step, idesc, 1u);
                        umma_commit(smem_u32(&empty_bar[s]));
                    }
                    __syncwarp();
                    acc = 1;
                    if (++s == S) { s = 0; fph ^= 1; }
                }
                if (leader) umma_commit(smem_u32(&acc_full[as]));
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps)
        // TMEM lane = output row: a thread holds 32 consecutive columns of its row.  Each warp adds the bias, writes its
        // 32x32 block into a 128-byte-swizzled shared-memory tile and one lane hands it to the TMA unit (plain store,
        // or fp32 reduce-add for accumulate) -- whole 128-byte lines, clipped at the matrix edge by the tensor map.
        // Two warps share each 32-lane TMEM quarter and alternate over the 32-column blocks.
        const int ew = warp - 2, lslice = (warp & 3) * 32, half = ew >> 2;
        const uint32_t stg = smem0 + (uint32_t)S * stage_bytes + (uint32_t)ew * 4096;
        int tl = 0;
        bool pending = false;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++tl) {
            const int nt = item % p.n_tiles, mt = (item / p.n_tiles) % p.m_tiles, sp = item / (p.n_tiles * p.m_tiles);
            const int as = tl & 1;
            const int row0 = mt * BM + lslice;
            mbar_wait(smem_u32(&acc_full[as]), (tl >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int j = half; j < BN / 32; j += 2) {
                const int c0 = nt * BN + j * 32;
                if (row0 >= p.M || c0 >= p.N) continue;                 // warp-uniform: nothing of this block is stored
                float4 bv[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    bv[c] = (p.bias && c0 + 4 * c < p.N) ? ldg4(p.bias + c0 + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t rg[32];
                tmem_ld32(tmem_base + ((uint32_t)lslice << 16) + (uint32_t)(as * BN + j * 32), rg);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (pending) {                                           // the previous store must have read the tile
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t dst = stg + (uint32_t)lane * 128 + (uint32_t)((c ^ (lane & 7)) * 16);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(__uint_as_float(rg[4 * c]) + bv[c].x),
                                 "f"(__uint_as_float(rg[4 * c + 1]) + bv[c].y), "f"(__uint_as_float(rg[4 * c + 2]) + bv[c].z),
                                 "f"(__uint_as_float(rg[4 * c + 3]) + bv[c].w) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (c0 < p.ncol0) tma_store_3d(&tc0, stg, c0, row0, sp, p.accumulate);
                    else tma_store_3d(&tc1, stg, c0 - p.ncol0, row0, 0, p.accumulate);
                }
                pending = true;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// out[i] (+)= sum over splits of part[s][i], fixed order (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits, long long n4,
                                                            float* __restrict__ out, int accumulate) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        float4 s = ldg4(part + q * 4);
        for (int k = 1; k < splits; ++k) {
            const float4 v = ldg4(part + ((size_t)k * n4 + q) * 4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        if (accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(out + q * 4);
            s.x += old.x; s.y += old.y; s.z += old.z; s.w += old.w;
        }
        stg4(out + q * 4, s);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// row-major fp32 matrix [rows][cols] with row stride ld (elements); box = box_rows x 32 columns, 128-byte swizzle
int make_map(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int box_rows, bool mn_major,
             const char* who) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { b200_set_error("%s: cuTensorMapEncodeTiled is unavailable", who); return B200_ERR_CUDA; }
    if (((uintptr_t)base & 15) || (ld & 3)) { b200_set_error("%s: operand must be 16-byte aligned with a row stride multiple of 4", who); return B200_ERR_ARG; }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b200_set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r); return B200_ERR_CUDA; }
    return B200_OK;
}

// MN-major operand as ONE box per stage: the row-major matrix [rows = reduction][cols] viewed as
// {32 columns, rows, cols / 32 column groups}; box {32, BK, groups} lands as [group][BK rows][128 B], the layout the
// per-box path builds with `groups` separate copies (a single elected thread issues every TMA, and the issue rate of
// 4 KB boxes -- not bandwidth -- was what bounded dgrad / wgrad).  Needs cols % 32 == 0.
int make_map_mn3d(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int groups, const char* who) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { b200_set_error("%s: cuTensorMapEncodeTiled is unavailable", who); return B200_ERR_CUDA; }
    if (((uintptr_t)base & 15) || (ld & 3) || (cols & 31)) { b200_set_error("%s: bad MN-major operand", who); return B200_ERR_ARG; }
    const cuuint64_t dims[3] = {32u, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128u};
    const cuuint32_t box[3] = {32u, (cuuint32_t)BK, (cuuint32_t)groups};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b200_set_error("%s: cuTensorMapEncodeTiled (3-D MN-major) failed (%d)", who, (int)r); return B200_ERR_CUDA; }
    return B200_OK;
}

// output [splits][rows][cols] fp32, row stride ld, split stride ss (elements); 32x32 boxes, 128-byte swizzle
int make_out_map(CUtensorMap* m, float* base, long long rows, long long cols, long long ld, long long splits, long long ss,
                 const char* who) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { b200_set_error("%s: cuTensorMapEncodeTiled is unavailable", who); return B200_ERR_CUDA; }
    if (((uintptr_t)base & 15) || (ld & 3) || (ss & 3)) { b200_set_error("%s: output must be 16-byte aligned with strides multiple of 4", who); return B200_ERR_ARG; }
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)splits};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(splits > 1 ? ss : rows * ld) * 4};
    const cuuint32_t box[3] = {32u, 32u, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b200_set_error("%s: cuTensorMapEncodeTiled (output) failed (%d)", who, (int)r); return B200_ERR_CUDA; }
    return B200_OK;
}

// tile width (multiple of 32, <= 256): minimise waves x per-tile cost, where a tile's k-block moves 128 + BN rows
int choose_bn(int N, long long m_tiles, int splits) {
    int best = 32;
    double best_cost = -1;
    const int sms = b200_num_sms();
    for (int bn = 256; bn >= 32; bn -= 32) {
        const long long tiles = m_tiles * ((N + bn - 1) / bn) * splits;
        const double cost = (double)((tiles + sms - 1) / sms) * (128 + bn);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

int wgrad_splits(long long tokens, int O, int I) {
    const long long m_tiles = (O + BM - 1) / BM;
    long long tiles = m_tiles * ((I + 255) / 256);
    const long long kb = (tokens + BK - 1) / BK;
    long long s = b200_num_sms() / tiles;
    if (s > kb / 8) s = kb / 8;                      // at least 8 k-blocks (256 tokens) per split
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int)s;
}

int launch_gemm(GemmP& p, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                cudaStream_t st, const char* who) {
    p.m_tiles = (p.M + BM - 1) / BM;
    p.n_tiles = (p.N + p.BN - 1) / p.BN;
    p.kb_total = (p.K + BK - 1) / BK;
    const int stage = A_BYTES + p.BN * BK * 4;
    int S = (190 * 1024) / stage;
    if (S > G_MAX_STAGES) S = G_MAX_STAGES;
    p.nstages = S;
    const int bytes = S * stage + 1024 + EPI_WARPS * 4096;   // + alignment slack + one 32x32 staging tile per epilogue warp
    CUtensorMap c0, c1;
    if (int rc = make_out_map(&c0, p.dst0, p.M, p.ncol0, p.ld0, p.splits, p.split_stride, who)) return rc;
    c1 = c0;
    if (p.dst1) if (int rc = make_out_map(&c1, p.dst1, p.M, p.N - p.ncol0, p.ld1, 1, 0, who)) return rc;
    static int attr_bytes = 0;
    if (bytes > attr_bytes) {
        cudaFuncSetAttribute(gemm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        attr_bytes = bytes;
    }
    const long long items = (long long)p.m_tiles * p.n_tiles * p.splits;
    const int grid = (int)(items < b200_num_sms() ? items : b200_num_sms());
    gemm_umma_kernel<<<grid, G_THREADS, bytes, st>>>(a0, a1, b0, b1, c0, c1, p);
    B200_CHECK_LAUNCH(who);
    return B200_OK;
}

bool dims_ok(long long M, int O, int c0, int c1) {
    const int I = c0 + c1;
    return M > 0 && M < (1ll << 31) && O >= 16 && (O & 3) == 0 && c0 > 0 && (c0 & 3) == 0 && c1 >= 0 && (c1 & 3) == 0 &&
           (c1 == 0 || (c0 & 31) == 0) && I >= 4;
}

}  // namespace

B200_API int b200_linear_supported(long long M, int O, int c0, int c1) { return dims_ok(M, O, c0, c1) ? 1 : 0; }

B200_API int b200_linear_fwd(const float* x0, const float* x1, int c0, int c1, const float* w, const float* bias, float* y,
                             long long M, int O, cudaStream_t st) {
    B200_REQUIRE(dims_ok(M, O, c0, c1), "linear_fwd: unsupported shape (M=%lld O=%d c0=%d c1=%d)", M, O, c0, c1);
    B200_REQUIRE(x0 && w && y && (c1 == 0 || x1), "linear_fwd: null pointer");
    const int I = c0 + c1;
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.M = (int)M; p.N = O; p.K = I; p.BN = choose_bn(O, (M + BM - 1) / BM, 1);
    p.a_k0 = c0; p.b_n0 = O; p.splits = 1;
    p.dst0 = y; p.ld0 = O; p.ncol0 = O; p.bias = bias;
    CUtensorMap a0, a1, b0;
    if (int rc = make_map(&a0, x0, M, c0, c0, BM, false, "linear_fwd")) return rc;
    a1 = a0;
    if (c1) if (int rc = make_map(&a1, x1, M, c1, c1, BM, false, "linear_fwd")) return rc;
    if (int rc = make_map(&b0, w, O, I, I, p.BN, false, "linear_fwd")) return rc;
    return launch_gemm(p, a0, a1, b0, b0, st, "linear_fwd");
}

B200_API int b200_linear_dgrad(const float* dy, const float* w, float* dx0, float* dx1, int c0, int c1, int accumulate,
                               long long M, int O, cudaStream_t st) {
    B200_REQUIRE(dims_ok(M, O, c0, c1), "linear_dgrad: unsupported shape (M=%lld O=%d c0=%d c1=%d)", M, O, c0, c1);
    B200_REQUIRE(dy && w && dx0 && (c1 == 0 || dx1), "linear_dgrad: null pointer");
    const int I = c0 + c1;
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.M = (int)M; p.N = I; p.K = O; p.BN = choose_bn(I, (M + BM - 1) / BM, 1);
    p.b_mn = 1; p.a_k0 = O; p.b_n0 = I; p.splits = 1;
    p.dst0 = dx0; p.ld0 = c0; p.ncol0 = c0; p.dst1 = dx1; p.ld1 = c1; p.accumulate = accumulate;
    CUtensorMap a0, b0;
    if (int rc = make_map(&a0, dy, M, O, O, BM, false, "linear_dgrad")) return rc;
    if ((I & 31) == 0) {                                                              // rows = reduction index o
        p.b_3d = 1;
        if (int rc = make_map_mn3d(&b0, w, O, I, I, p.BN / 32, "linear_dgrad")) return rc;
    } else if (int rc = make_map(&b0, w, O, I, I, BK, true, "linear_dgrad")) return rc;   // 32x32 boxes
    return launch_gemm(p, a0, a0, b0, b0, st, "linear_dgrad");
}

B200_API long long b200_linear_wgrad_workspace_bytes(long long M, int O, int I) {
    const int s = wgrad_splits(M, O, I);
    return s > 1 ? (long long)s * O * I * (long long)sizeof(float) : 0;
}

B200_API int b200_linear_wgrad(const float* x0, const float* x1, int c0, int c1, const float* dy, float* dw, int accumulate,
                               float* workspace, long long workspace_bytes, long long M, int O, cudaStream_t st) {
    B200_REQUIRE(dims_ok(M, O, c0, c1), "linear_wgrad: unsupported shape (M=%lld O=%d c0=%d c1=%d)", M, O, c0, c1);
    B200_REQUIRE(x0 && dy && dw && (c1 == 0 || x1), "linear_wgrad: null pointer");
    const int I = c0 + c1;
    const int splits = wgrad_splits(M, O, I);
    if (splits > 1 && (!workspace || workspace_bytes < b200_linear_wgrad_workspace_bytes(M, O, I))) {
        b200_set_error("linear_wgrad: workspace too small (%lld < %lld bytes)", workspace_bytes, b200_linear_wgrad_workspace_bytes(M, O, I));
        return B200_ERR_WORKSPACE;
    }
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.M = O; p.N = I; p.K = (int)M; p.BN = choose_bn(I, (O + BM - 1) / BM, splits);
    p.a_mn = 1; p.b_mn = 1; p.a_k0 = (int)M; p.b_n0 = c0; p.splits = splits;
    p.dst0 = splits > 1 ? workspace : dw; p.ld0 = I; p.ncol0 = I; p.split_stride = (long long)O * I;
    p.accumulate = splits > 1 ? 0 : accumulate;
    CUtensorMap a0, b0, b1;
    if ((O & 31) == 0) {
        p.a_3d = 1;
        if (int rc = make_map_mn3d(&a0, dy, M, O, O, BM / 32, "linear_wgrad")) return rc;
    } else if (int rc = make_map(&a0, dy, M, O, O, BK, true, "linear_wgrad")) return rc;
    if (c1 == 0 && (c0 & 31) == 0) {
        p.b_3d = 1;
        if (int rc = make_map_mn3d(&b0, x0, M, c0, c0, p.BN / 32, "linear_wgrad")) return rc;
    } else if (int rc = make_map(&b0, x0, M, c0, c0, BK, true, "linear_wgrad")) return rc;
    b1 = b0;
    if (c1) if (int rc = make_map(&b1, x1, M, c1, c1, BK, true, "linear_wgrad")) return rc;
    if (int rc = launch_gemm(p, a0, a0, b0, b1, st, "linear_wgrad")) return rc;
    if (splits > 1) {
        const long long n4 = (long long)O * I / 4;
        const int grid = (int)((n4 + 255) / 256 < 4 * b200_num_sms() ? (n4 + 255) / 256 : 4 * b200_num_sms());
        splitk_reduce_kernel<<<grid, 256, 0, st>>>(workspace, splits, n4, dw, accumulate);
        B200_CHECK_LAUNCH("linear_wgrad_reduce");
    }
    return B200_OK;
}
