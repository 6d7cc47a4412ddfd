// Forward / data gradient of the wide-image 2D 3x3 stride-1 pad-1 convolutions (code/networks/unet.py:37,41 at 256^2 and
// 128^2: the 16- and 32-channel levels that hold most of the bytes of the UNet) on the 5th-generation tensor cores, with
// the BatchNorm batch statistics of the output taken in the epilogue (code/networks/unet.py:38,42 train mode).
//
// Formulation ("row ring", K-major twin of conv_row_wgrad.cu).  A persistent CTA owns a contiguous range of output
// rows (n, h).  Input image rows are staged ONCE by TMA as [plane][P pixels][RB bytes] (plane = 32 channels / 128-byte
// rows / 128-byte swizzle, or one 16-channel tensor / 64-byte rows / 64-byte swizzle; one zero pixel on the left and
// >= 1 on the right and whole zero rows above / below the image come from the TMA out-of-bounds fill = the padding)
// into a ring of R rows: moving down one output row loads one new input row.  The weights of the whole layer
// ([tap][plane][cout][RB], <= 74 KB for these layers) are loaded once per CTA and stay resident.
// The A operand of tap (kh, kw) for the 128 output pixels [128b, 128b+127] of a row is simply the ring row h+kh-1
// starting kw pixels further: the UMMA swizzle is a function of the absolute shared-memory address, so a tap is a
// descriptor start-address shift -- no im2col, no per-tap copies, and M = 128 is exactly 128 real pixels (W % 128 == 0).
//   warp 0     TMA producer (elected lane): weights once, then one box per source per new input row
//   warp 1     MMA issuer (elected lane): 9 taps x planes x (2|4) k-steps of tcgen05.mma kind::tf32 into one of four
//              TMEM accumulators; tcgen05.commit publishes the accumulator and releases ring rows
//   warps 2-5  epilogue: tcgen05.ld -> + bias -> swizzled shared-memory tile -> ONE TMA store (or reduce-add) per
//              destination; per-channel sum / sum of squares of the tile from shared memory into per-thread running sums,
//              written once per CTA as fp64 partials for bn_finalize (deterministic).
// The data gradient is the same kernel on dy with flipped / transposed packed weights and up to two destinations
// (the two sources of a virtual concat).
//
// 16-channel tensors (the 256^2 level) run in PIXEL-PAIR mode: a 64-byte pixel makes a 64-byte TMA box row, and the TMA
// unit moves one box row per ~7 cycles whatever its length (measured: 1.9 TB/s over the chip, for loads and stores alike).
// Two neighbouring pixels x 16 channels are one 128-byte row instead; the GEMM then runs on pairs: M = 128 pairs (a whole
// 256-pixel row), K = (pixel parity pb, cin), N = (pixel parity pa, cout), and the horizontal taps become three pair
// shifts s with kw = 2 (s - 1) + pb - pa + 1 baked into expanded weights (conv_row_pack.cuh); the k-steps whose weight
// block is all zero (half of them for s = 0 and s = 2) are skipped.  The accumulator row (pa, cout) IS the output pair
// in memory, so the epilogue is that of a 32-channel layer.
#include "umma_common.cuh"
#include "conv_row_pack.cuh"
#include <cstdlib>
#include <cstring>
#include "../../include/b200ssl.h"

namespace {

using namespace umma;

constexpr int CR_THREADS = 320;      // TMA warp, MMA warp, 4 epilogue warps, 4 statistics warps
constexpr int MAX_R = 8;
constexpr int NACC = 4;

struct RowP {
    int N, H, W, P;              // images, rows, k-pixels per row (pixels, or pixel pairs in pair mode); k-pixels per staged row
                                 // (multiple of 8, >= W + 2)
    int NP, NP0;                 // planes in total / served by the first source
    int P1;                      // pixels of the first box of a row (a second box brings P - P1 when P > 256)
    int Cout;                    // GEMM columns (16 / 32 / 64)
    int nblk;                    // 128-pixel blocks per row
    int R;                       // ring rows
    int plane_bytes, slot_bytes, w_bytes;
    int G, ngrp, grp0;           // channels per output staging group (16 / 32), groups, groups going to the first destination
    int fold;                    // statistics: column groups that are the same channel (2 in pair mode)
    int bias_mask;               // bias index = column & bias_mask
    const float* bias;
    double* stats;               // [gridDim.x][2][Cout] or null
    int accumulate;
    int debug;                   // profiling only: 1 no statistics, 2 no store, 4 epilogue releases the accumulator and nothing else, 8 no MMAs
};

template <int CPP, bool PAIR>
__global__ void __launch_bounds__(CR_THREADS, 1) conv_row_kernel(const __grid_constant__ CUtensorMap tx0a,
                                                                 const __grid_constant__ CUtensorMap tx0b,
                                                                 const __grid_constant__ CUtensorMap tx1a,
                                                                 const __grid_constant__ CUtensorMap tx1b,
                                                                 const __grid_constant__ CUtensorMap tw,
                                                                 const __grid_constant__ CUtensorMap ty0,
                                                                 const __grid_constant__ CUtensorMap ty1, const RowP p) {
    constexpr int RB = CPP * 4;                       // bytes of one pixel of a plane
    constexpr int KS = CPP / 8;                       // k-steps per plane
    constexpr uint32_t LAYOUT = CPP == 32 ? 2u : 4u;  // 128-byte / 64-byte swizzle, K-major
    constexpr uint32_t SBO = 8 * RB;                  // 8-row core-matrix group
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_r[MAX_R], empty_r[MAX_R], w_full, acc_full[NACC], acc_empty[NACC];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (role branches on the uniform datapath)
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = smem0;
    const uint32_t ring = w_base + (uint32_t)((p.w_bytes + 1023) & ~1023);
    const uint32_t stg_bytes = (uint32_t)p.ngrp * 128u * (uint32_t)p.G * 4u;
    const uint32_t stg = (ring + (uint32_t)p.R * p.slot_bytes + 1023u) & ~1023u;      // two staging buffers
    const int Rtot = p.N * p.H;
    const int r0 = (int)((long long)blockIdx.x * Rtot / gridDim.x), r1 = (int)((long long)(blockIdx.x + 1) * Rtot / gridDim.x);
    const int tmem_cols = NACC * p.Cout < 32 ? 32 : (NACC * p.Cout <= 64 ? 64 : (NACC * p.Cout <= 128 ? 128 : 256));

    if (tid == 0) {
        for (int s = 0; s < p.R; ++s) { mbar_init(smem_u32(&full_r[s]), 1); mbar_init(smem_u32(&empty_r[s]), 1); }
        mbar_init(smem_u32(&w_full), 1);
        for (int a = 0; a < NACC; ++a) { mbar_init(smem_u32(&acc_full[a]), 1); mbar_init(smem_u32(&acc_empty[a]), 4); }
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        const bool leader = elect_one();
        if (leader) {
            tma_prefetch_desc(&tx0a); tma_prefetch_desc(&tw); tma_prefetch_desc(&ty0);
            mbar_expect_tx(smem_u32(&w_full), (uint32_t)p.w_bytes);
            for (int i = 0; i < (PAIR ? 6 : 9) * p.NP; ++i)
                tma_load_2d(w_base + (uint32_t)i * p.Cout * RB, &tw, 0, i * p.Cout, smem_u32(&w_full));
        }
        __syncwarp();
        int slot = 0, ph = 0, cnt = 0;
        int n = r0 / p.H, h = r0 - n * p.H;
        const int np1 = p.NP - p.NP0;
        const uint32_t row_tx = (uint32_t)p.slot_bytes;
        for (int r = r0; r < r1; ++r) {
            const bool new_seg = (r == r0) || (h == 0);
            const int first = new_seg ? h - 1 : h + 1, nload = new_seg ? 3 : 1;
            for (int k = 0; k < nload; ++k, ++cnt) {
                if (cnt >= p.R) mbar_wait(smem_u32(&empty_r[slot]), ph ^ 1);
                const uint32_t fb = smem_u32(&full_r[slot]);
                const uint32_t dst = ring + (uint32_t)slot * p.slot_bytes;
                const int row = first + k;
                if (leader) {
                    mbar_expect_tx(fb, row_tx);
                    tma_load_5d(dst, &tx0a, 0, -1, 0, row, n, fb);
                    if (p.P1 < p.P) tma_load_5d(dst + (uint32_t)p.P1 * RB, &tx0b, 0, p.P1 - 1, 0, row, n, fb);
                    if (np1) {
                        const uint32_t d1 = dst + (uint32_t)p.NP0 * p.plane_bytes;
                        tma_load_5d(d1, &tx1a, 0, -1, 0, row, n, fb);
                        if (p.P1 < p.P) tma_load_5d(d1 + (uint32_t)p.P1 * RB, &tx1b, 0, p.P1 - 1, 0, row, n, fb);
                    }
                }
                __syncwarp();
                if (++slot == p.R) { slot = 0; ph ^= 1; }
            }
            if (++h == p.H) { h = 0; ++n; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
        const bool leader = elect_one();
        const uint32_t idesc = idesc_tf32(128, p.Cout, 0, 0);
        const uint64_t plane16 = (uint64_t)(p.plane_bytes >> 4);
        const uint64_t wplane16 = (uint64_t)((p.Cout * RB) >> 4);             // one (tap, plane) weight block
        const uint64_t wtap16 = wplane16 * (uint64_t)p.NP;
        const uint64_t wd0 = smem_desc(w_base, 16, SBO, LAYOUT);
        mbar_wait(smem_u32(&w_full), 0);
        int wslot = 0;                         // ring slot of input row h-1
        int fslot = 0, fph = 0, ahead = 0;     // next slot to wait for / its parity / rows of the window already waited
        int buf = 0, bph = 0, nb = 0;          // accumulator buffer / parity of its "empty" barrier / blocks issued
        int h = r0 % p.H;
        for (int r = r0; r < r1; ++r) {
            for (; ahead < 3; ++ahead) {
                mbar_wait(smem_u32(&full_r[fslot]), fph);
                if (++fslot == p.R) { fslot = 0; fph ^= 1; }
            }
            tc_fence_after();
            const int w1 = wslot + 1 == p.R ? 0 : wslot + 1, w2 = w1 + 1 == p.R ? 0 : w1 + 1;
            const uint64_t a0 = smem_desc(ring + (uint32_t)wslot * p.slot_bytes, 16, SBO, LAYOUT);
            const uint64_t a1 = smem_desc(ring + (uint32_t)w1 * p.slot_bytes, 16, SBO, LAYOUT);
            const uint64_t a2 = smem_desc(ring + (uint32_t)w2 * p.slot_bytes, 16, SBO, LAYOUT);
            for (int b = 0; b < p.nblk; ++b, ++nb) {
                if (nb >= NACC) mbar_wait(smem_u32(&acc_empty[buf]), bph ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * p.Cout);
                const uint64_t boff = (uint64_t)((128 * b * RB) >> 4);
                uint32_t acc = 0;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    uint64_t ap = (kh == 0 ? a0 : (kh == 1 ? a1 : a2)) + boff;
                    uint64_t wp = wd0 + (uint64_t)((PAIR ? 2 : 3) * kh) * wtap16;
#pragma unroll 1
                    for (int pl = 0; pl < p.NP; ++pl, ap += plane16, wp += wplane16) {
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            // pair mode: shift -1 only meets the odd input pixel (k-steps 2, 3), shift +1 only the even one
                            constexpr int KLO[3] = {PAIR ? 2 : 0, 0, 0}, KHI[3] = {KS, KS, PAIR ? 2 : KS};
#pragma unroll
                            for (int ks = KLO[kw]; ks < KHI[kw]; ++ks) {
                                if (leader && !(p.debug & 8))
                                    mma_tf32(d, ap + (uint64_t)(kw * (RB >> 4) + 2 * ks),
                                             wp + (uint64_t)(PAIR ? (kw == 1 ? 0 : 1) : kw) * wtap16 + (uint64_t)(2 * ks), idesc, acc);
                                acc = 1;
                            }
                        }
                    }
                }
                if (leader) mma_commit(smem_u32(&acc_full[buf]));
                __syncwarp();
                if (++buf == NACC) { buf = 0; bph ^= 1; }
            }
            const bool last = (h == p.H - 1) || (r == r1 - 1);
            if (leader) {
                mma_commit(smem_u32(&empty_r[wslot]));
                if (last) { mma_commit(smem_u32(&empty_r[w1])); mma_commit(smem_u32(&empty_r[w2])); }
            }
            __syncwarp();
            if (last) { wslot = w2 + 1 == p.R ? 0 : w2 + 1; ahead = 0; }
            else { wslot = w1; ahead = 2; }
            h = (h == p.H - 1) ? 0 : h + 1;
        }
    } else if (warp >= 6) {
        // ------------------------------------------------------------------ statistics warps (BatchNorm sum / sum of squares)
        // The epilogue warps are the throughput limit of this kernel (profiles/r2_ncu_full_conv_row.csv: ~3400 warp
        // instructions per 128-pixel block, 60 % of them the statistics pass), so the statistics of a staged tile are taken
        // by four more warps while the epilogue warps already convert the next accumulator.  Hand-off through named
        // barriers per staging buffer: 6 + sb "tile staged" (epilogue arrives, statistics warps wait), 4 + sb "tile read"
        // (statistics warps arrive, the epilogue waits before overwriting the buffer two blocks later).
        const int st_ = (warp - 6) * 32 + lane;
        const int Nc = p.Cout, G = p.G;
        const int nsub = 128 / Nc;                                // threads sharing a channel
        const int sc = st_ % Nc, ssub = st_ / Nc;
        const bool use_s = p.stats && !(p.debug & 5);
        if (use_s) {
            double d1 = 0.0, d2 = 0.0;
            const int total = (r1 - r0) * p.nblk;
            const int gi = sc / G, cg = sc % G;
            for (int k = 0; k < total; ++k) {
                const int sb = k & 1;
                const uint32_t sbase = stg + (uint32_t)sb * stg_bytes;
                if (sb == 0) asm volatile("bar.sync 6, 256;" ::: "memory"); else asm volatile("bar.sync 7, 256;" ::: "memory");
                // channel sc over rows [ssub * Nc, ssub * Nc + Nc) of the staged tile (nsub * Nc == 128)
                const uint32_t gb = sbase + (uint32_t)gi * (128u * G * 4u) + (uint32_t)((cg & 3) * 4);
                float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                for (int i = 0; i < Nc; ++i) {
                    const int row = ssub * Nc + i;
                    const int srow = G == 32 ? row : (row >> 1), chunk = (G == 32 ? 0 : (row & 1) * 4) + (cg >> 2);
                    float v;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(gb + (uint32_t)srow * 128u + (uint32_t)((chunk ^ (srow & 7)) * 16)));
                    s1 += v;
                    s2 += v * v;
                }
                d1 += (double)s1;
                d2 += (double)s2;
                if (k + 2 < total) {                              // somebody will wait for this buffer again
                    if (sb == 0) asm volatile("bar.arrive 4, 256;" ::: "memory"); else asm volatile("bar.arrive 5, 256;" ::: "memory");
                }
            }
            // cross-thread reduction in the (now idle) first staging buffer: [sub][2][64] doubles = 8 KB
            asm volatile("bar.sync 3, 256;" ::: "memory");        // the epilogue's last TMA store has read the staging buffers
            double* sred = reinterpret_cast<double*>(smem_raw + (stg - smem_u32(smem_raw)));
            sred[(ssub * 2 + 0) * 64 + sc] = d1;
            sred[(ssub * 2 + 1) * 64 + sc] = d2;
            asm volatile("bar.sync 8, 128;" ::: "memory");
            const int Cr = Nc / p.fold;                          // real channels (pair mode: columns (pa, c) fold onto c)
            if (st_ < 2 * Cr) {
                const int which = st_ / Cr, c = st_ % Cr;
                double s = 0.0;
                for (int k = 0; k < nsub; ++k)
                    for (int fo = 0; fo < p.fold; ++fo) s += sred[(k * 2 + which) * 64 + fo * Cr + c];
                p.stats[(size_t)blockIdx.x * 2 * Cr + which * Cr + c] = s;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (4 warps = 128 accumulator rows)
        const int q = warp & 3, et = q * 32 + lane;              // TMEM lane quarter; thread index within the epilogue
        const int Nc = p.Cout, G = p.G;
        const bool use_s = p.stats && !(p.debug & 5);            // statistics warps active
        int buf = 0, fph = 0, sb = 0, nstore = 0, kblk = 0;
        int n = r0 / p.H, h = r0 - n * p.H;
        for (int r = r0; r < r1; ++r) {
            for (int b = 0; b < p.nblk; ++b) {
                mbar_wait(smem_u32(&acc_full[buf]), fph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Nc);
                const uint32_t sbase = stg + (uint32_t)sb * stg_bytes;
                if (p.debug & 4) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
                    if (++buf == NACC) { buf = 0; fph ^= 1; }
                    continue;
                }
                // staging buffer sb was handed to the TMA unit two blocks ago: its reads must be done
                if (et == 0 && nstore >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (use_s && kblk >= 2) {                         // ... and the statistics warps must have read it
                    if (sb == 0) asm volatile("bar.sync 4, 256;" ::: "memory"); else asm volatile("bar.sync 5, 256;" ::: "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
                for (int c0 = 0; c0 < Nc; c0 += 16) {
                    uint32_t rg[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(rg[0]), "=r"(rg[1]), "=r"(rg[2]), "=r"(rg[3]), "=r"(rg[4]), "=r"(rg[5]), "=r"(rg[6]), "=r"(rg[7]),
                          "=r"(rg[8]), "=r"(rg[9]), "=r"(rg[10]), "=r"(rg[11]), "=r"(rg[12]), "=r"(rg[13]), "=r"(rg[14]), "=r"(rg[15])
                        : "r"(taddr + (uint32_t)c0));
                    tmem_ld_wait();
                    const int gi = c0 / G, cg0 = c0 % G;                   // staging group and first channel within it
                    // staging rows are 128 bytes with the 128-byte TMA swizzle: one pixel of a 32-channel group, or a pixel
                    // PAIR of a 16-channel destination (row = et >> 1, the odd pixel in the upper 64 bytes)
                    const uint32_t gbase = sbase + (uint32_t)gi * (128u * G * 4u);
                    const int srow = G == 32 ? et : (et >> 1), cofs = G == 32 ? 0 : (et & 1) * 4;
                    const uint32_t rowb = gbase + (uint32_t)srow * 128u;
                    const int f = srow & 7;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b4 = p.bias ? ldg4(p.bias + ((c0 + 4 * j) & p.bias_mask)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const int chunk = cofs + (cg0 >> 2) + j;
                        const uint32_t dst = rowb + (uint32_t)((chunk ^ f) * 16);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                                     "f"(__uint_as_float(rg[4 * j]) + b4.x), "f"(__uint_as_float(rg[4 * j + 1]) + b4.y),
                                     "f"(__uint_as_float(rg[4 * j + 2]) + b4.z), "f"(__uint_as_float(rg[4 * j + 3]) + b4.w)
                                     : "memory");
                    }
                }
                // the accumulator is in registers / shared memory: hand the TMEM buffer back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
                fence_proxy_async();
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (use_s) {
                    if (sb == 0) asm volatile("bar.arrive 6, 256;" ::: "memory"); else asm volatile("bar.arrive 7, 256;" ::: "memory");
                }
                if (et == 0 && !(p.debug & 2)) {
                    // G = 32: box {32 channels, 128 k-pixels}; G = 16: box {one 128-byte pixel pair, 64 pairs}
                    const int pix0 = G == 32 ? (n * p.H + h) * p.W + 128 * b : ((n * p.H + h) * p.W + 128 * b) >> 1;
                    for (int gi = 0; gi < p.ngrp; ++gi) {
                        const uint32_t src = sbase + (uint32_t)gi * (128u * G * 4u);
                        const bool first = gi < p.grp0;
                        const CUtensorMap* tm = first ? &ty0 : &ty1;
                        const int c = G == 32 ? (first ? gi : gi - p.grp0) * G : 0;
                        if (p.accumulate)
                            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(tm), "r"(src), "r"(c), "r"(pix0) : "memory");
                        else
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(tm), "r"(src), "r"(c), "r"(pix0) : "memory");
                    }
                    tma_commit_group();
                    ++nstore;
                }
                ++kblk;
                sb ^= 1;
                if (++buf == NACC) { buf = 0; fph ^= 1; }
            }
            if (++h == p.H) { h = 0; ++n; }
        }
        if (use_s) {
            if (et == 0) tma_wait_group_read0();
            asm volatile("bar.sync 3, 256;" ::: "memory");        // hands the staging buffers to the statistics warps' reduction
        }
        if (et == 0) tma_wait_group0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

__global__ void __launch_bounds__(256) conv_row_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int mode, int O, int I,
                                                            int total) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
        out[idx] = row_pack_elem(w, mode, O, I, 9, idx);
}

struct RGeo {
    int cpp, pair, NP, NP0, Ncols, G, ngrp, grp0, fold, Wk, P, P1, nblk, R, plane, slot, wbytes, smem, mode;
};

// dgrad = 0: A = [src0|src1] (c0 + c1 channels), columns = cout.  dgrad = 1: A = dy (cout channels), columns = c0 + c1.
bool rgeometry(const b200_conv_desc* d, int dgrad, RGeo& g) {
    if (d->id != 1 || d->kd != 1 || d->kh != 3 || d->kw != 3 || d->stride != 1 || d->ph != 1 || d->pw != 1 || d->pd != 0) return false;
    if (d->n < 1 || d->ih < 1 || d->iw < 128) return false;
    const int a0 = dgrad ? d->cout : d->c0, a1 = dgrad ? 0 : d->c1;      // channels of the A-operand sources
    const int n0 = dgrad ? d->c0 : d->cout, n1 = dgrad ? d->c1 : 0;      // channels of the destinations
    if (n1 != 0 && n1 != n0) return false;
    g.pair = a0 == 16 && (a1 == 0 || a1 == 16) && n0 == 16 && d->iw % 256 == 0;
    if (const char* e = getenv("B200_ROW_NOPAIR")) if (e[0] == '1') g.pair = 0;       // A/B runs only
    if (g.pair) {
        g.cpp = 32; g.NP0 = 1; g.NP = a1 ? 2 : 1;
        g.Wk = d->iw / 2;
        g.Ncols = 2 * (n0 + n1); g.G = 32; g.ngrp = g.Ncols / 32; g.grp0 = 1; g.fold = 2;
    } else {
        if (a0 % 32 == 0 && a1 % 32 == 0 && a0 > 0) { g.cpp = 32; g.NP0 = a0 / 32; g.NP = (a0 + a1) / 32; }
        else if (a0 == 16 && (a1 == 0 || a1 == 16)) { g.cpp = 16; g.NP0 = 1; g.NP = a1 ? 2 : 1; }
        else return false;
        g.Wk = d->iw;
        g.Ncols = n0 + n1;
        if (g.Ncols != 16 && g.Ncols != 32 && g.Ncols != 64) return false;
        g.G = n0 >= 32 ? 32 : 16;
        if (n0 % g.G != 0) return false;
        g.ngrp = g.Ncols / g.G; g.grp0 = n0 / g.G; g.fold = 1;
    }
    if (g.Wk % 128 != 0 || g.Wk > 384) return false;
    g.mode = (dgrad ? 1 : 0) | (g.cpp == 16 ? 2 : 0) | (g.pair ? 4 : 0);
    g.P = (g.Wk + 2 + 7) / 8 * 8;
    g.P1 = g.P <= 256 ? g.P : 136;
    if (g.P - g.P1 > 256) return false;
    const int pa0 = g.pair ? 32 : a0, pa1 = g.pair ? (a1 ? 32 : 0) : a1;           // channels per k-pixel of each source
    if (g.P1 < g.P && (pa0 > g.cpp || pa1 > g.cpp)) return false;                 // a split row needs one plane per source
    g.nblk = g.Wk / 128;
    const int RB = g.cpp * 4;
    g.plane = g.P * RB;
    g.slot = g.NP * g.plane;
    g.wbytes = (g.pair ? 6 : 9) * g.NP * g.Ncols * RB;
    const int stage = 2 * g.ngrp * 128 * g.G * 4;
    g.R = 0;
    for (int r = MAX_R; r >= 4; --r) {
        const int bytes = 1024 + ((g.wbytes + 1023) & ~1023) + r * g.slot + 1024 + stage;
        if (bytes <= 224 * 1024) { g.R = r; g.smem = bytes; break; }
    }
    return g.R != 0;
}

}  // namespace

// 0 = this convolution is not served by the row kernels; otherwise 8 + the packing mode (bit 0 data gradient,
// bit 1 16-channel planes, bit 2 pixel-pair mode) -- the mode goes into the batch packer's job table
B200_API int b200_conv_row_supported(const b200_conv_desc* d, int dgrad) {
    RGeo g;
    return (d && rgeometry(d, dgrad, g)) ? 8 + g.mode : 0;
}

B200_API long long b200_conv_row_packed_floats(const b200_conv_desc* d, int dgrad) {
    RGeo g;
    if (!d || !rgeometry(d, dgrad, g)) return 0;
    return row_pack_total(g.mode, d->cout, d->c0 + d->c1);
}

B200_API int b200_conv_row_pack_weights(const b200_conv_desc* d, int dgrad, const float* w, float* out, cudaStream_t st) {
    RGeo g;
    B200_REQUIRE(d && w && out && rgeometry(d, dgrad, g), "conv_row_pack_weights: unsupported convolution");
    const int total = (int)row_pack_total(g.mode, d->cout, d->c0 + d->c1);
    conv_row_pack_kernel<<<(total + 255) / 256 < 64 ? (total + 255) / 256 : 64, 256, 0, st>>>(w, out, g.mode, d->cout, d->c0 + d->c1, total);
    B200_CHECK_LAUNCH("conv_row_pack_weights");
    return B200_OK;
}

B200_API long long b200_conv_row_stats_blocks(const b200_conv_desc* d) {
    if (!d) return 0;
    const long long rows = (long long)d->n * d->ih;
    return rows < b200_num_sms() ? rows : b200_num_sms();
}

template <int CPP, bool PAIR>
static void launch_row(int grid, int smem, cudaStream_t st, const CUtensorMap& tx0a, const CUtensorMap& tx0b, const CUtensorMap& tx1a,
                       const CUtensorMap& tx1b, const CUtensorMap& tw, const CUtensorMap& ty0, const CUtensorMap& ty1, const RowP& p) {
    static int attr = 0;                                       // per instantiation
    if (smem > attr) { cudaFuncSetAttribute(conv_row_kernel<CPP, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = smem; }
    conv_row_kernel<CPP, PAIR><<<grid, CR_THREADS, smem, st>>>(tx0a, tx0b, tx1a, tx1b, tw, ty0, ty1, p);
}

static int run_row(const b200_conv_desc* d, int dgrad, const float* a0, const float* a1, const float* wpk, const float* bias,
                   float* dst0, float* dst1, double* stats, int accumulate, cudaStream_t st, const char* who) {
    RGeo g;
    B200_REQUIRE(d && rgeometry(d, dgrad, g), "%s: unsupported convolution", who);
    const int N = d->n, H = d->ih, W = d->iw;
    const int ca0 = dgrad ? d->cout : d->c0, ca1 = dgrad ? 0 : d->c1;
    const int cn0 = dgrad ? d->c0 : d->cout, cn1 = dgrad ? d->c1 : 0;
    B200_REQUIRE(a0 && wpk && dst0 && (ca1 == 0 || a1) && (cn1 == 0 || dst1), "%s: null pointer", who);
    CUtensorMap tx0a, tx0b, tx1a, tx1b, tw, ty0, ty1;
    const CUtensorMapSwizzle swz = g.cpp == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    auto make_in = [&](CUtensorMap* m, const float* base, int C, int boxp) -> int {
        // channels-last [N][H][W][C] as {cpp floats, k-pixels of a row, planes, H, N}; pair mode: a k-pixel is 2 pixels x 16 channels
        const int Ck = g.pair ? 32 : C;                       // floats per k-pixel
        const int cg = Ck / g.cpp;
        const cuuint64_t rowb = (cuuint64_t)W * C * 4;
        const cuuint64_t dims[5] = {(cuuint64_t)g.cpp, (cuuint64_t)g.Wk, (cuuint64_t)cg, (cuuint64_t)H, (cuuint64_t)N};
        const cuuint64_t strides[4] = {(cuuint64_t)Ck * 4, (cuuint64_t)g.cpp * 4, rowb, rowb * H};
        const cuuint32_t box[5] = {(cuuint32_t)g.cpp, (cuuint32_t)boxp, (cuuint32_t)cg, 1u, 1u};
        return make_tmap(m, base, 5, dims, strides, box, swz, who);
    };
    if (int rc = make_in(&tx0a, a0, ca0, g.P1)) return rc;
    tx0b = tx0a;
    if (g.P1 < g.P) if (int rc = make_in(&tx0b, a0, ca0, g.P - g.P1)) return rc;
    tx1a = tx0a; tx1b = tx0b;
    if (ca1) {
        if (int rc = make_in(&tx1a, a1, ca1, g.P1)) return rc;
        tx1b = tx1a;
        if (g.P1 < g.P) if (int rc = make_in(&tx1b, a1, ca1, g.P - g.P1)) return rc;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)g.cpp, (cuuint64_t)((g.pair ? 6 : 9) * g.NP * g.Ncols)};
        const cuuint64_t strides[1] = {(cuuint64_t)g.cpp * 4};
        const cuuint32_t box[2] = {(cuuint32_t)g.cpp, (cuuint32_t)g.Ncols};
        if (int rc = make_tmap(&tw, wpk, 2, dims, strides, box, swz, who)) return rc;
    }
    auto make_out = [&](CUtensorMap* m, float* base, int C) -> int {
        // 16-channel destinations are stored through their pixel-pair view (128-byte rows): pair mode has 128 pairs per
        // accumulator block, the plain modes 64
        const bool pv = C == 16;
        const cuuint64_t dims[2] = {(cuuint64_t)(pv ? 32 : C), (cuuint64_t)N * H * W / (pv ? 2 : 1)};
        const cuuint64_t strides[1] = {(cuuint64_t)(pv ? 32 : C) * 4};
        const cuuint32_t box[2] = {32u, (pv && !g.pair) ? 64u : 128u};
        return make_tmap(m, base, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, who);
    };
    if (int rc = make_out(&ty0, dst0, cn0)) return rc;
    ty1 = ty0;
    if (cn1) if (int rc = make_out(&ty1, dst1, cn1)) return rc;

    RowP p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.H = H; p.W = g.Wk; p.P = g.P; p.NP = g.NP; p.NP0 = g.NP0; p.P1 = g.P1; p.Cout = g.Ncols; p.nblk = g.nblk; p.R = g.R;
    p.plane_bytes = g.plane; p.slot_bytes = g.slot; p.w_bytes = g.wbytes; p.G = g.G; p.ngrp = g.ngrp; p.grp0 = g.grp0;
    p.bias = bias; p.stats = stats; p.accumulate = accumulate; p.fold = g.fold; p.bias_mask = (g.pair ? 16 : g.Ncols) - 1;
    { const char* e = getenv("B200_ROW_DEBUG"); p.debug = e ? atoi(e) : 0; }
    const long long rows = (long long)N * H;
    const int grid = (int)(rows < b200_num_sms() ? rows : b200_num_sms());
    if (g.pair) launch_row<32, true>(grid, g.smem, st, tx0a, tx0b, tx1a, tx1b, tw, ty0, ty1, p);
    else if (g.cpp == 32) launch_row<32, false>(grid, g.smem, st, tx0a, tx0b, tx1a, tx1b, tw, ty0, ty1, p);
    else launch_row<16, false>(grid, g.smem, st, tx0a, tx0b, tx1a, tx1b, tw, ty0, ty1, p);
    B200_CHECK_LAUNCH(who);
    return B200_OK;
}

B200_API int b200_conv_row_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wpk, const float* bias,
                               float* dst, double* stats_part, cudaStream_t st) {
    return run_row(d, 0, src0, src1, wpk, bias, dst, nullptr, stats_part, 0, st, "conv_row_fwd");
}

B200_API int b200_conv_row_dgrad(const b200_conv_desc* d, const float* dy, const float* wpk_dgrad, float* dx0, float* dx1,
                                 int accumulate, cudaStream_t st) {
    return run_row(d, 1, dy, nullptr, wpk_dgrad, nullptr, dx0, dx1, nullptr, accumulate, st, "conv_row_dgrad");
}
