// 3x3 stride-1 "same" convolution (2D) on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) with the
// accumulators in TMEM -- forward and data gradient (same kernel, flipped weights).
//
// Formulation.  The batch is treated as one tall zero-padded image of N*(H+2) rows x (W+2) columns.  A CTA owns
// a tile of TH output rows x TW output columns; it stages the (TH+2) x (TW+2) halo of 16 input channels in shared
// memory ONCE per channel chunk with cp.async (zero-fill at the borders = the convolution padding), laid out
// k-chunk-major:   halo[kq = c/4][q = rr*(TW+2) + cc][4 floats]
// so that 8 consecutive halo pixels x 16 bytes form one contiguous 128-byte UMMA "core matrix".  With the
// no-swizzle K-major canonical layout ((8,m),(T,2)):((1T,SBO),(1,LBO)) a tap (kh,kw) is then nothing but a
// shift of the A descriptor's start address by (kh*(TW+2) + kw) pixels: 128 consecutive halo-linear positions are
// the M = 128 rows of one MMA (the 2 halo columns per row produce junk rows that the epilogue skips).  No im2col,
// no per-tap copies, no address arithmetic on the data path: 9 taps x 2 k-steps = 18 MMAs per 16 channels and
// 128-pixel block, issued by one thread; tcgen05.commit arrives on an mbarrier when they retire, which both
// frees the shared-memory stage and releases the epilogue.
//
// Epilogue: each warp reads its 32 TMEM lanes (= 32 pixels) with tcgen05.ld 32x32b.x16, adds the bias and writes
// whole pixels (16 channels = 64 contiguous bytes per thread).
#include "conv_common.cuh"
#include "../../include/b200ssl.h"

namespace {

constexpr int UKC = 16;          // channels per chunk (two K = 8 TF32 MMAs per tap)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    uint32_t spins = 0;
    do {
        if (++spins > (1u << 28)) __trap();      // a lost arrival must fail loudly, not hang the device
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 B contiguous,
// SBO = byte distance between 8-row groups, LBO = byte distance between the two 16-byte K chunks of one MMA
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
    return d;                                    // layout_type (bits 61-63) = 0: SWIZZLE_NONE
}

// D[tmem] (+)= A[smem] * B[smem]^T, M = 128, N from idesc, K = 8 (TF32)
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

__global__ void pack_umma_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int dgrad, int O, int I,
                                         int T, int rows, int cols, int colsP, int chunks) {
    const long long total = (long long)chunks * T * 4 * colsP * 4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx & 3);
        long long r = idx >> 2;
        const int col = (int)(r % colsP); r /= colsP;
        const int kq = (int)(r & 3); r >>= 2;
        const int tap = (int)(r % T);
        const int chunk = (int)(r / T);
        const int row = chunk * UKC + kq * 4 + j;
        float v = 0.f;
        if (row < rows && col < cols) {
            v = dgrad ? w[((size_t)row * I + col) * T + (T - 1 - tap)] : w[((size_t)col * I + row) * T + tap];
            v = __uint_as_float(f2tf32(v));
        }
        out[idx] = v;
    }
}

inline int round16(int v) { return (v + 15) / 16 * 16; }


int umma_supported(const b200_conv_desc* d) {
    return d && d->stride == 1 && d->kd == 1 && d->kh == 3 && d->kw == 3 && d->pd == 0 && d->ph == 1 && d->pw == 1 &&
           (d->c0 & 3) == 0 && (d->c1 & 3) == 0 && d->id == 1;
}

}  // namespace

B200_API int b200_conv_umma_supported(const b200_conv_desc* d, int for_dgrad) {
    if (!umma_supported(d)) return 0;
    if (for_dgrad) return (d->cout & 3) == 0;        // dy is the staged operand; dx channel counts are already multiples of 4
    return d->cout <= 4 || (d->cout & 3) == 0;
}

B200_API long long b200_conv_umma_packed_floats(int dgrad, int O, int I, int T) {
    const int rows = dgrad ? O : I, cols = dgrad ? I : O;
    return (long long)((rows + UKC - 1) / UKC) * T * 4 * round16(cols) * 4;
}

B200_API int b200_conv_umma_pack_weights(const float* w, float* out, int dgrad, int O, int I, int T, cudaStream_t st) {
    B200_REQUIRE(w && out && O > 0 && I > 0 && T > 0, "conv_umma_pack_weights: bad arguments");
    const int rows = dgrad ? O : I, cols = dgrad ? I : O;
    const int chunks = (rows + UKC - 1) / UKC, colsP = round16(cols);
    const long long total = (long long)chunks * T * 4 * colsP * 4;
    int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    pack_umma_weights_kernel<<<blocks, 256, 0, st>>>(w, out, dgrad, O, I, T, rows, cols, colsP, chunks);
    B200_CHECK_LAUNCH("conv_umma_pack_weights");
    return B200_OK;
}

// =====================================================================================================
// The convolution kernel: persistent and warp-specialised (the simpler one-tile-per-CTA first version is in the git history).
//   warps 0-3  producers: per 16-channel chunk the halo arrives with coalesced 16-byte cp.async (four lanes cover the
//              64 bytes of one pixel and scatter them into the four 4-channel planes the UMMA descriptor wants;
//              zero-fill = conv padding); cp.async.mbarrier.arrive.noinc signals the stage's "full" barrier when a
//              thread's copies have landed.  One lane adds the packed weights with cp.async.bulk (expect_tx).
//              (A 4D TMA tiled load of the same box was measured first: its inner extent is only 16 bytes, which
//              costs ~4 cycles per 16-byte row -- 2.5 us per tile -- so the LSU path is used for the halo.)
//   warp 4     MMA issuer (one lane): waits "full", issues MB x 9 taps x 2 tcgen05.mma, tcgen05.commit -> "empty";
//              also owns the TMEM allocation (two accumulator buffers)
//   warps 5-8  epilogue: wait for the accumulator barrier, tcgen05.ld, bias, vector stores, release the buffer
// Each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; the shared-memory ring (up to 8 stages) and the
// two TMEM buffers run continuously across tiles, so the loads of the next tiles are in flight while the tensor
// core works and the epilogue drains the previous accumulator.
// =====================================================================================================
namespace {

struct Umma2P {
    const float* src0;
    const float* src1;
    int C0, C1, Cin;
    int N, H, W;
    int TH, TW, HW, MB;          // tile rows / cols, halo width, 128-row accumulator blocks per tile
    int NPIXA;
    int tiles_w, tiles_h;
    const float* wt;
    int Cout, CoutP;
    const float* bias;
    float* dst0;
    float* dst1;
    int D0, D1;
    int out_nchw;
    int accumulate;
    int nstages;
    FastDiv fd_hw;
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cpasync_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int MAX_STAGES = 8;
constexpr int PRODUCER_THREADS = 128;
constexpr int U2_THREADS = 288;          // 4 producer warps + 1 MMA warp + 4 epilogue warps

template <int BN>
__global__ void __launch_bounds__(U2_THREADS) conv_umma2_kernel(const Umma2P p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (role branches on the uniform datapath)
    const int n0 = blockIdx.y * BN;
    const int plane_a = p.NPIXA * 16;
    const int halo_bytes = 4 * plane_a;
    const int stage_bytes = halo_bytes + 36 * BN * 16;
    const int nchunks = (p.Cin + UKC - 1) / UKC;
    const int S = p.nstages;
    const int tiles_img = p.tiles_h * p.tiles_w;
    const int ntiles = p.N * tiles_img;
    const int acc_cols = p.MB * BN;                       // columns of one accumulator buffer
    int tmem_cols = 32;
    while (tmem_cols < 2 * acc_cols) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&full_bar[s]), PRODUCER_THREADS + 1);   // 128 cp.async arrivals + the weight expect_tx
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&acc_full[a]), 1);
            mbar_init(smem_u32(&acc_empty[a]), 4);        // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t smem0 = smem_u32(smem);

    if (warp < 4) {
        // ------------------------------------------------------------------ producers
        const int npix_real = (p.TH + 2) * p.HW;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int n = tile / tiles_img, r = tile % tiles_img;
            const int h0 = (r / p.tiles_w) * p.TH, w0 = (r % p.tiles_w) * p.TW;
            for (int chunk = 0; chunk < nchunks; ++chunk, ++it) {
                const int s = it % S;
                if (it >= S) mbar_wait(smem_u32(&empty_bar[s]), ((it / S) - 1) & 1);
                const uint32_t fb = smem_u32(&full_bar[s]);
                const uint32_t st = smem0 + s * stage_bytes;
                if (tid == 0) {
                    const float* wsrc = p.wt + (size_t)chunk * 36 * p.CoutP * 4;
                    mbar_expect_tx(fb, 36u * BN * 16u);
                    if (BN == p.CoutP) {
                        bulk_load(st + halo_bytes, wsrc, 36u * BN * 16u, fb);
                    } else {
                        for (int tk = 0; tk < 36; ++tk)
                            bulk_load(st + halo_bytes + tk * BN * 16, wsrc + ((size_t)tk * p.CoutP + n0) * 4, BN * 16u, fb);
                    }
                }
                for (int sl = tid; sl < npix_real * 4; sl += PRODUCER_THREADS) {
                    const int kq = sl & 3, q = sl >> 2;
                    uint32_t rr, cc;
                    p.fd_hw.divmod((uint32_t)q, rr, cc);
                    const int ih = h0 + (int)rr - 1, iw = w0 + (int)cc - 1, ch = chunk * UKC + kq * 4;
                    const bool ok = (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W && ch < p.Cin;
                    const float* src = p.src0;
                    if (ok) {
                        const size_t pix = ((size_t)n * p.H + ih) * p.W + iw;
                        src = ch < p.C0 ? p.src0 + pix * p.C0 + ch : p.src1 + pix * p.C1 + (ch - p.C0);
                    }
                    cp16(st + kq * plane_a + q * 16, src, ok);
                }
                cpasync_arrive_noinc(fb);                  // arrives once this thread's copies have landed
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop (warp-uniform control flow keeps the descriptors in uniform registers; under
        // `if (lane == 0)` every tcgen05.mma is wrapped in a uniformisation loop) and ONE elected lane issues.  The
        // descriptors of a stage are built once; a tap / k-step / row block only adds to their 14-bit address field
        // (units of 16 B = one halo pixel of a 4-channel plane).  tools/umma_peak.cu: an N = 64 TF32 MMA retires every
        // 48 cycles, the previous 40-instruction loop body issued one every ~250.
        {
            uint32_t leader_u;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(leader_u));
            const bool leader = leader_u != 0;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t a_ks = (uint64_t)((2 * plane_a) >> 4);          // second k-step: two 4-channel planes further
            const uint64_t row1 = (uint64_t)p.HW, row2 = (uint64_t)(2 * p.HW);
            int it = 0, tl = 0;
            int s = 0, fph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
                const int as = tl & 1;
                // the epilogue must have drained this accumulator buffer (used two tiles ago)
                if (tl >= 2) mbar_wait(smem_u32(&acc_empty[as]), ((tl >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol = tmem_base + as * acc_cols;
                for (int chunk = 0; chunk < nchunks; ++chunk, ++it) {
                    mbar_wait(smem_u32(&full_bar[s]), fph);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // cp.async (generic proxy) data -> tensor core
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sh = smem0 + s * stage_bytes, sw = sh + halo_bytes;
                    const uint64_t ad0 = umma_desc(sh, plane_a, 128);
                    const uint64_t bd0 = umma_desc(sw, BN * 16, 128);
                    const uint32_t first = chunk == 0 ? 0u : 1u;
#pragma unroll 1
                    for (int b = 0; b < p.MB; ++b) {
                        const uint64_t ab = ad0 + (uint64_t)(b * 128);
                        const uint32_t d = dcol + b * BN;
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int kh = tap / 3, kw = tap % 3;
                            const uint64_t at = ab + (kh == 0 ? 0 : (kh == 1 ? row1 : row2)) + (uint64_t)kw;
                            const uint64_t bt = bd0 + (uint64_t)(tap * 4 * BN);
                            if (leader) {
                                umma_tf32(d, at, bt, idesc, tap == 0 ? first : 1u);
                                umma_tf32(d, at + a_ks, bt + (uint64_t)(2 * BN), idesc, 1u);
                            }
                        }
                    }
                    if (leader) umma_commit(smem_u32(&empty_bar[s]));     // frees the stage when these MMAs retire
                    __syncwarp();
                    if (++s == S) { s = 0; fph ^= 1; }
                }
                if (leader) umma_commit(smem_u32(&acc_full[as]));          // accumulator of this tile complete
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int lslice = (warp & 3) * 32;                    // TMEM lanes this warp may read
        const size_t S_img = (size_t)p.H * p.W;
        int tl = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
            const int as = tl & 1;
            const int n = tile / tiles_img, r = tile % tiles_img;
            const int h0 = (r / p.tiles_w) * p.TH, w0 = (r % p.tiles_w) * p.TW;
            mbar_wait(smem_u32(&acc_full[as]), (tl >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int b = 0; b < p.MB; ++b) {
                const int q = b * 128 + lslice + lane;
                uint32_t rr, cc;
                p.fd_hw.divmod((uint32_t)q, rr, cc);
                const int oh = h0 + (int)rr, ow = w0 + (int)cc;
                const bool valid = (int)rr < p.TH && (int)cc < p.TW && oh < p.H && ow < p.W;
                const size_t sp = (size_t)oh * p.W + ow;
                const size_t pix = (size_t)n * S_img + sp;
#pragma unroll 1
                for (int j = 0; j < BN / 16; ++j) {
                    uint32_t rg[16];
                    tmem_ld16(tmem_base + ((uint32_t)lslice << 16) + (uint32_t)(as * acc_cols + b * BN + j * 16), rg);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const int c0 = n0 + j * 16;
                    if (!valid || c0 >= p.Cout) continue;
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        v[e] = __uint_as_float(rg[e]);
                        if (p.bias && c0 + e < p.Cout) v[e] += __ldg(p.bias + c0 + e);
                    }
                    if (p.out_nchw) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (c0 + e < p.Cout) {
                                float* o = p.dst0 + ((size_t)n * p.Cout + c0 + e) * S_img + sp;
                                *o = p.accumulate ? *o + v[e] : v[e];
                            }
                    } else {
#pragma unroll
                        for (int e4 = 0; e4 < 4; ++e4) {
                            const int c = c0 + e4 * 4;
                            if (c >= p.Cout) break;
                            float* o = c < p.D0 ? p.dst0 + pix * p.D0 + c : p.dst1 + pix * p.D1 + (c - p.D0);
                            float4 val = make_float4(v[e4 * 4], v[e4 * 4 + 1], v[e4 * 4 + 2], v[e4 * 4 + 3]);
                            if (p.accumulate) {
                                const float4 old = *reinterpret_cast<const float4*>(o);
                                val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                            }
                            *reinterpret_cast<float4*>(o) = val;
                        }
                    }
                }
            }
            // all TMEM reads of this warp are complete (wait::ld above): hand the buffer back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// tile geometry: maximise the fraction of accumulator rows that are real outputs, prefer more rows per tile
void choose_tile2(int H, int W, int BN, int& TW, int& TH, int& MB, int& S, int& NPIXA) {
    const int cand[7] = {W, 14, 30, 32, 62, 64, 126};
    double best = -1;
    for (int mb = 1; mb <= 4; ++mb) {
        if (2 * mb * BN > 512) break;                      // two TMEM accumulator buffers
        for (int i = 0; i < 7; ++i) {
            const int tw = cand[i] > W ? W : cand[i];
            const int hw = tw + 2;
            int th = (mb * 128) / hw;
            if (th < 1) continue;
            if (th > H) th = H;
            const int npixa = (mb * 128 + 2 * hw + 2 + 7) / 8 * 8;
            const int stage = 4 * npixa * 16 + 36 * BN * 16;
            int s = (200 * 1024) / stage;                     // persistent CTA: use the shared memory for a deep ring
            if (s > MAX_STAGES) s = MAX_STAGES;
            if (s < 2) continue;
            const double cover = (double)H * W / ((double)((H + th - 1) / th) * ((W + tw - 1) / tw) * mb * 128);
            const double score = cover + 0.01 * mb;             // tie-break: bigger tiles reuse the weights more
            if (score > best) { best = score; TW = tw; TH = th; MB = mb; S = s; NPIXA = npixa; }
        }
    }
}

template <int BN>
int launch_umma2(const Umma2P& p, cudaStream_t st) {
    auto kern = conv_umma2_kernel<BN>;
    const int bytes = p.nstages * (4 * p.NPIXA * 16 + 36 * BN * 16) + 1024;
    static int attr_bytes = 0;
    if (bytes > attr_bytes) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        attr_bytes = bytes;
    }
    const int coltiles = (p.Cout + BN - 1) / BN;
    const int ntiles = p.N * p.tiles_h * p.tiles_w;
    int gx = b200_num_sms() / coltiles;                       // one persistent CTA per SM in total
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    dim3 grid(gx, coltiles);
    kern<<<grid, U2_THREADS, bytes, st>>>(p);
    B200_CHECK_LAUNCH("conv_umma2");
    return B200_OK;
}

int run_umma2(Umma2P& p, cudaStream_t st) {
    const int BN = p.Cout <= 16 ? 16 : (p.Cout <= 32 ? 32 : (p.Cout <= 64 ? 64 : 128));
    p.TW = 0;
    choose_tile2(p.H, p.W, BN, p.TW, p.TH, p.MB, p.nstages, p.NPIXA);
    if (p.TW == 0) { b200_set_error("conv_umma: no tile geometry fits"); return B200_ERR_ARG; }
    p.HW = p.TW + 2;
    p.tiles_w = (p.W + p.TW - 1) / p.TW;
    p.tiles_h = (p.H + p.TH - 1) / p.TH;
    p.fd_hw.init(p.HW);
    if (BN == 16) return launch_umma2<16>(p, st);
    if (BN == 32) return launch_umma2<32>(p, st);
    if (BN == 64) return launch_umma2<64>(p, st);
    return launch_umma2<128>(p, st);
}

}  // namespace

B200_API int b200_conv_umma2_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wt,
                                 const float* bias, float* dst, int out_nchw, cudaStream_t st) {
    B200_REQUIRE(b200_conv_umma_supported(d, 0), "conv_umma2_fwd: unsupported convolution");
    B200_REQUIRE(src0 && wt && dst && (d->c1 == 0 || src1), "conv_umma2_fwd: null pointer");
    B200_REQUIRE(out_nchw || (d->cout & 3) == 0, "conv_umma2_fwd: channels-last output needs cout % 4 == 0");
    Umma2P p;
    memset(&p, 0, sizeof(p));
    p.src0 = src0; p.src1 = src1; p.C0 = d->c0; p.C1 = d->c1; p.Cin = d->c0 + d->c1;
    p.N = d->n; p.H = d->ih; p.W = d->iw;
    p.wt = wt; p.Cout = d->cout; p.CoutP = round16(d->cout); p.bias = bias;
    p.dst0 = dst; p.D0 = d->cout; p.D1 = 0; p.out_nchw = out_nchw;
    return run_umma2(p, st);
}

B200_API int b200_conv_umma2_dgrad(const b200_conv_desc* d, const float* dy, const float* wt_dgrad, float* dx0, float* dx1,
                                   int accumulate, cudaStream_t st) {
    B200_REQUIRE(b200_conv_umma_supported(d, 1), "conv_umma2_dgrad: unsupported convolution");
    B200_REQUIRE(dy && wt_dgrad && dx0 && (d->c1 == 0 || dx1), "conv_umma2_dgrad: null pointer");
    Umma2P p;
    memset(&p, 0, sizeof(p));
    p.src0 = dy; p.C0 = d->cout; p.C1 = 0; p.Cin = d->cout;
    p.N = d->n; p.H = d->ih; p.W = d->iw;
    p.wt = wt_dgrad; p.Cout = d->c0 + d->c1; p.CoutP = round16(p.Cout);
    p.dst0 = dx0; p.dst1 = dx1; p.D0 = d->c0; p.D1 = d->c1; p.accumulate = accumulate;
    return run_umma2(p, st);
}
