// Forward / data gradient of the narrow-image 2D 3x3 stride-1 pad-1 convolutions (code/networks/unet.py:37,41 at the
// 64^2 / 32^2 / 16^2 levels: 32..256 channels, the tensor-bound half of the UNet) on the 5th-generation tensor cores.
//
// Formulation.  A work item is (image, block of TR output rows, 64 output channels).  For every 32-channel plane of the
// input the (TR + 2) x (W + 2) halo of the row block is staged ONCE by one TMA box (128-byte rows = one pixel of the
// plane, 128-byte swizzle; the zero padding is the box's out-of-bounds fill) and the packed weights of the plane arrive
// per filter row (3 taps x 64 x 128 B) through their own ring.  The halo is halo-linear: position q = r (W + 2) + c, so
// tap (kh, kw) of the 128 consecutive output positions of an accumulator block is the SAME shared-memory plane starting
// kh (W + 2) + kw positions further -- the UMMA swizzle is a function of the absolute shared-memory address, so a tap is
// a descriptor start-address shift (the 2 of every W + 2 accumulator rows that are halo columns are junk the epilogue
// skips).  All <= 3 accumulator blocks of an item live in TMEM at once (two item buffers), so a weight chunk is read from
// shared memory for 3 x 128 positions and from L2 once per item.
//   warp 0     TMA producer (elected lane): halo planes (ring of 2), weight chunks (ring of 2-3), mbarrier expect_tx
//   warp 1     MMA issuer (elected lane): per (plane, filter row) 3 blocks x 3 taps x 4 k-steps of tcgen05.mma kind::tf32
//   warps 2-5  epilogue: tcgen05.ld -> + bias -> per-warp swizzled staging tile -> coalesced 16-byte stores of the valid
//              pixels (two destinations for the data gradient of a virtual concat); BatchNorm (sum, sum^2) of the tile from
//              the staged values, accumulated per CTA in fp64 and written as one partial per CTA for b200_bn_finalize (code/networks/unet.py:38,42)
// Packed weights: the row-kernel layout [tap][plane][column][32] (conv_row_pack.cuh, mode = data-gradient flag).
//
// 3D (code/networks/vnet.py:28, 3x3x3 stride 1 pad 1, >= 32 channels): the same kernel with a depth axis on the halo --
// an item is (volume, TD output planes, TR output rows, 64 channels), its halo box is (TD + 2) x (TR + 2) x (W + 2)
// positions, position q = (dd (TR + 2) + r) (W + 2) + c, and tap (kd, kh, kw) is the start-address shift
// (kd (TR + 2) + kh) (W + 2) + kw; weight chunks arrive per (plane, kd, kh).  (TD, TR) per layer come from a small cost
// model (tensor-pipe cycles of the accumulator blocks against the TMA bytes of halo + weights, times the number of waves).
#include "umma_common.cuh"
#include "conv_row_pack.cuh"
#include <cstdlib>
#include <cstring>
#include "../../include/b200ssl.h"

namespace {

using namespace umma;

constexpr int CB_THREADS = 320;      // TMA warp + MMA warp + 8 epilogue warps
constexpr int MAX_NW = 4;
constexpr int MAX_NB = 3;

struct BlkP {
    int N, D, H, W, P;           // images, planes (1 in 2D), rows, columns, halo pitch W + 2
    int KD, TD, tiles_d, R2;     // depth taps (1 or 3), output planes per item, plane blocks, halo rows per plane TR + 2
    int TR, tiles_r, nb;         // output rows per item, row blocks per image, 128-position accumulator blocks per item
    int NP, NP0;                 // 32-channel planes of the reduction operand in total / served by the first source
    int Nt, ntiles_n, Ntot;      // GEMM columns per item, column tiles, total GEMM columns
    int a_bytes, a_box_bytes;    // halo ring slot, bytes one halo box delivers
    int w_bytes, NW;             // weight ring slot (3 taps x Nt x 128 B), slots
    const float* bias;
    float* dst0;
    float* dst1;
    int D0, D1;                  // channels of the destinations (columns [0, D0) -> dst0, the rest -> dst1)
    int accumulate;
    double* stats;               // [gridDim.x][2][Ntot] or null (forward only: Ntot = cout <= 256)
    int debug;                   // profiling only: 1 no MMAs, 2 epilogue only releases the accumulator, 4 no TMA loads / waits,
                                 // 8 every tap reads the unshifted halo (wrong results: cost of atom-misaligned A operands)
    FastDiv fdP, fdR2;
};

// CPP: channels per halo row -- 32 (128-byte rows, 128-byte swizzle) or 16 (16-channel tensors: 64-byte rows, 64-byte swizzle)
template <int CPP>
__global__ void __launch_bounds__(CB_THREADS, 1) conv_blk_kernel(const __grid_constant__ CUtensorMap tx0,
                                                                 const __grid_constant__ CUtensorMap tx1,
                                                                 const __grid_constant__ CUtensorMap tw, const BlkP p) {
    constexpr int RB = CPP * 4;                       // bytes of one halo position of a plane
    constexpr int KS = CPP / 8;                       // k-steps per plane
    constexpr uint32_t LAYOUT = CPP == 32 ? 2u : 4u;  // K-major, 128-byte / 64-byte swizzle
    constexpr uint32_t SBO = 8 * RB;                  // 8-row core-matrix group
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], w_full[MAX_NW], w_empty[MAX_NW], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float sred[4][2][64];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler: the role branches and their loops run on the uniform datapath
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_ring = smem0;
    const uint32_t w_ring = a_ring + 2u * p.a_bytes;
    const uint32_t stg = w_ring + (uint32_t)p.NW * p.w_bytes;
    const int nitems = p.N * p.tiles_d * p.tiles_r * p.ntiles_n;

    // the accumulator blocks read past the halo box (junk rows): those bytes must be finite, and no box ever writes them
    {
        const uint32_t tail0 = (uint32_t)p.a_box_bytes & ~15u, tail = (uint32_t)p.a_bytes - tail0;
        for (uint32_t o = (uint32_t)tid * 16; o < 2u * tail; o += CB_THREADS * 16) {
            const uint32_t slot = o >= tail ? 1u : 0u, off = o - slot * tail;
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_ring + slot * p.a_bytes + tail0 + off), "r"(0u) : "memory");
        }
    }
    fence_proxy_async();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1);
            mbar_init(smem_u32(&acc_full[s]), 1); mbar_init(smem_u32(&acc_empty[s]), 8);
        }
        for (int s = 0; s < p.NW; ++s) { mbar_init(smem_u32(&w_full[s]), 1); mbar_init(smem_u32(&w_empty[s]), 1); }
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        const bool leader = elect_one();
        if (leader) { tma_prefetch_desc(&tx0); tma_prefetch_desc(&tx1); tma_prefetch_desc(&tw); }
        int as = 0, aph = 0, acnt = 0, ws = 0, wph = 0, wcnt = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int nt = item % p.ntiles_n, t = item / p.ntiles_n, tr = t % p.tiles_r, t2 = t / p.tiles_r;
            const int td = t2 % p.tiles_d, n = t2 / p.tiles_d;
            const int h0 = tr * p.TR, z0 = td * p.TD - (p.KD == 3 ? 1 : 0);
            for (int pl = 0; pl < p.NP; ++pl) {
                if (acnt >= 2) mbar_wait(smem_u32(&a_empty[as]), aph ^ 1);
                if (leader && !(p.debug & 4)) {
                    const uint32_t fb = smem_u32(&a_full[as]);
                    mbar_expect_tx(fb, (uint32_t)p.a_box_bytes);
                    if (pl < p.NP0) tma_load_5d(a_ring + (uint32_t)as * p.a_bytes, &tx0, CPP * pl, -1, h0 - 1, z0, n, fb);
                    else tma_load_5d(a_ring + (uint32_t)as * p.a_bytes, &tx1, CPP * (pl - p.NP0), -1, h0 - 1, z0, n, fb);
                }
                __syncwarp();
                ++acnt;
                if (++as == 2) { as = 0; aph ^= 1; }
                for (int kh = 0; kh < 3 * p.KD; ++kh) {              // (kd, kh) filter rows
                    if (wcnt >= p.NW) mbar_wait(smem_u32(&w_empty[ws]), wph ^ 1);
                    if (leader && !(p.debug & 4)) {
                        const uint32_t fb = smem_u32(&w_full[ws]);
                        mbar_expect_tx(fb, (uint32_t)p.w_bytes);
                        for (int kw = 0; kw < 3; ++kw)
                            tma_load_2d(w_ring + (uint32_t)ws * p.w_bytes + (uint32_t)(kw * p.Nt * RB), &tw, 0,
                                        ((kh * 3 + kw) * p.NP + pl) * p.Ntot + nt * p.Nt, fb);
                    }
                    __syncwarp();
                    ++wcnt;
                    if (++ws == p.NW) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
        const bool leader = elect_one();
        const uint32_t idesc = idesc_tf32(128, p.Nt, 0, 0);
        const uint64_t wtap16 = (uint64_t)((p.Nt * RB) >> 4);
        const uint64_t row16 = (uint64_t)(p.P * (RB >> 4));           // one halo row in 16-byte units
        int as = 0, aph = 0, ws = 0, wph = 0, il = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++il) {
            const int buf = il & 1;
            if (il >= 2) mbar_wait(smem_u32(&acc_empty[buf]), ((il >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)(buf * 256);
            for (int pl = 0; pl < p.NP; ++pl) {
                if (!(p.debug & 4)) mbar_wait(smem_u32(&a_full[as]), aph);
                const uint64_t ad0 = smem_desc(a_ring + (uint32_t)as * p.a_bytes, 16, SBO, LAYOUT);
                for (int kk = 0, kd = 0, kh = 0; kk < 3 * p.KD; ++kk) {
                    if (!(p.debug & 4)) mbar_wait(smem_u32(&w_full[ws]), wph);
                    tc_fence_after();
                    const uint64_t wd0 = smem_desc(w_ring + (uint32_t)ws * p.w_bytes, 16, SBO, LAYOUT);
                    const uint32_t first = (pl | kk) == 0 ? 0u : 1u;
                    const uint64_t tap16 = (p.debug & 8) ? 0ull : (uint64_t)(kd * p.R2 + kh) * row16;
                    if (++kh == 3) { kh = 0; ++kd; }
#pragma unroll 1
                    for (int b = 0; b < p.nb; ++b) {
                        const uint64_t ab = ad0 + (uint64_t)(b * 8 * RB) + tap16;                   // 128 positions
                        const uint32_t d = d0 + (uint32_t)(b * p.Nt);
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) {
                                if (leader && !(p.debug & 1))
                                    mma_tf32(d, ab + (uint64_t)(((p.debug & 8) ? 0 : kw * (RB >> 4)) + 2 * ks), wd0 + (uint64_t)kw * wtap16 + (uint64_t)(2 * ks), idesc,
                                             (kw | ks) == 0 ? first : 1u);
                            }
                        }
                    }
                    if (leader) mma_commit(smem_u32(&w_empty[ws]));
                    __syncwarp();
                    if (++ws == p.NW) { ws = 0; wph ^= 1; }
                }
                if (leader) mma_commit(smem_u32(&a_empty[as]));
                __syncwarp();
                if (++as == 2) { as = 0; aph ^= 1; }
            }
            if (leader) mma_commit(smem_u32(&acc_full[buf]));
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps)
        // Two warps share each 32-lane TMEM quarter (a warp may only read the quarter warp_id % 4) and split its columns:
        // a warp owns 32 accumulator rows x Nt/2 columns of every block, stages them in its own swizzled tile and writes
        // the valid pixels with coalesced 16-byte stores (TPP lanes cover the Nt/2 * 4 contiguous bytes of one pixel).
        const int ew = warp - 2, q = warp & 3, half = ew >> 2;
        const int Nt = p.Nt, Nh = Nt == 16 ? 16 : Nt >> 1;           // columns per item / per warp
        const bool idle = Nt == 16 && half == 1;                     // 16-column items: one warp per lane quarter does the work
        const int TPP = Nh >> 2, RPI = 32 / TPP;                     // lanes per pixel, pixels per store instruction
        const int chunk = lane % TPP, rsub = lane / TPP;
        const uint32_t rowb = (uint32_t)(Nh * 4);
        const uint32_t stw = stg + (uint32_t)ew * 32u * rowb;        // this warp's staging tile: 32 rows x Nh floats
        const int et = ew * 32 + lane;
        double acc1 = 0.0, acc2 = 0.0;                               // BatchNorm sums of GEMM column et over this CTA's items
        int il = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++il) {
            const int nt = item % p.ntiles_n, t = item / p.ntiles_n, tr = t % p.tiles_r, t2 = t / p.tiles_r;
            const int td = t2 % p.tiles_d, n = t2 / p.tiles_d;
            const int h0 = tr * p.TR, z0 = td * p.TD, buf = il & 1;
            const int col0 = nt * Nt;
            const bool second = col0 >= p.D0;
            float* dst = second ? p.dst1 : p.dst0;
            const int ldc = second ? p.D1 : p.D0, cofs = (second ? col0 - p.D0 : col0) + half * Nh + chunk * 4;
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
            mbar_wait(smem_u32(&acc_full[buf]), (il >> 1) & 1);
            tc_fence_after();
            if (p.debug & 2) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
                continue;
            }
            for (int b = 0; b < (idle ? 0 : p.nb); ++b) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + b * Nt + half * Nh);
#pragma unroll 1
                for (int c0 = 0; c0 < Nh; c0 += 16) {
                    uint32_t rg[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(rg[0]), "=r"(rg[1]), "=r"(rg[2]), "=r"(rg[3]), "=r"(rg[4]), "=r"(rg[5]), "=r"(rg[6]), "=r"(rg[7]),
                          "=r"(rg[8]), "=r"(rg[9]), "=r"(rg[10]), "=r"(rg[11]), "=r"(rg[12]), "=r"(rg[13]), "=r"(rg[14]), "=r"(rg[15])
                        : "r"(taddr + (uint32_t)c0));
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b4 = p.bias ? ldg4(p.bias + col0 + half * Nh + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const int ch = (c0 >> 2) + j;
                        const uint32_t dsts = stw + (uint32_t)lane * rowb + (uint32_t)((ch ^ (lane & (TPP - 1) & 7)) * 16);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dsts),
                                     "f"(__uint_as_float(rg[4 * j]) + b4.x), "f"(__uint_as_float(rg[4 * j + 1]) + b4.y),
                                     "f"(__uint_as_float(rg[4 * j + 2]) + b4.z), "f"(__uint_as_float(rg[4 * j + 3]) + b4.w)
                                     : "memory");
                    }
                }
                if (b == p.nb - 1) {                           // every TMEM read of this item is done: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
                }
                __syncwarp();
                // coalesced pass, four pixels in flight per lane; (r, c) of the halo-linear position by one division per pixel
                const uint32_t pos0 = (uint32_t)(b * 128 + q * 32);
#pragma unroll 1
                for (int it = 0; it < 32; it += 4 * RPI) {
                    float4 v[4];
                    uint32_t rr[4], cc[4], dd[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int row = it + u * RPI + rsub;
                        uint32_t hr;
                        p.fdP.divmod(pos0 + (uint32_t)row, hr, cc[u]);
                        p.fdR2.divmod(hr, dd[u], rr[u]);
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                                     : "r"(stw + (uint32_t)row * rowb + (uint32_t)((chunk ^ (row & (TPP - 1) & 7)) * 16)));
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (it + u * RPI + rsub < 32 && (int)cc[u] < p.W && (int)rr[u] < p.TR && h0 + (int)rr[u] < p.H &&
                            (int)dd[u] < p.TD && z0 + (int)dd[u] < p.D) {
                            float* o = dst + (((size_t)(n * p.D + z0 + (int)dd[u]) * p.H + h0 + (int)rr[u]) * p.W + cc[u]) * ldc + cofs;
                            if (p.accumulate) {
                                const float4 old = *reinterpret_cast<const float4*>(o);
                                v[u].x += old.x; v[u].y += old.y; v[u].z += old.z; v[u].w += old.w;
                            }
                            stg4(o, v[u]);
                            s1[0] += v[u].x; s1[1] += v[u].y; s1[2] += v[u].z; s1[3] += v[u].w;
                            s2[0] += v[u].x * v[u].x; s2[1] += v[u].y * v[u].y; s2[2] += v[u].z * v[u].z; s2[3] += v[u].w * v[u].w;
                        }
                    }
                }
                __syncwarp();
            }
            if (idle) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
            }
            if (p.stats) {
                // lanes with the same chunk (same four channels) across the RPI pixel groups, then the four quarter-warps
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    for (int o = TPP; o < 32; o <<= 1) {
                        s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
                        s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
                    }
                asm volatile("bar.sync 1, 256;" ::: "memory");          // sred of the previous item has been consumed
                if (rsub == 0 && !idle) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        sred[q][0][half * Nh + chunk * 4 + k] = s1[k];
                        sred[q][1][half * Nh + chunk * 4 + k] = s2[k];
                    }
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                // epilogue thread et owns GEMM column et for the whole launch (cout <= 256): items are visited in a fixed
                // order, so the per-CTA fp64 sums are deterministic
                if (et >= col0 && et < col0 + Nt) {
                    const int c = et - col0;
                    acc1 += (double)sred[0][0][c] + (double)sred[1][0][c] + (double)sred[2][0][c] + (double)sred[3][0][c];
                    acc2 += (double)sred[0][1][c] + (double)sred[1][1][c] + (double)sred[2][1][c] + (double)sred[3][1][c];
                }
            }
        }
        if (p.stats && et < p.Ntot) {
            p.stats[((size_t)blockIdx.x * 2 + 0) * p.Ntot + et] = acc1;
            p.stats[((size_t)blockIdx.x * 2 + 1) * p.Ntot + et] = acc2;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void __launch_bounds__(256) conv_blk_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int mode, int O, int I,
                                                            int T, int total) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
        out[idx] = row_pack_elem(w, mode, O, I, T, idx);
}

struct BGeo {
    int KD, P, TD, tiles_d, R2, TR, tiles_r, nb, NP, NP0, Nt, ntiles_n, Ntot, a_bytes, a_box, w_bytes, NW, smem, D0, D1, cpp;
};

// shared-memory plan of one (TD, TR) candidate; false when it does not fit
bool bfit(BGeo& g, int KD, int W, int TD, int TR) {
    g.TD = TD; g.TR = TR; g.R2 = TR + 2;
    const int span = ((TD - 1) * g.R2 + TR - 1) * g.P + W;           // halo-linear positions from the first to the last output
    g.nb = (span + 127) / 128;
    if (g.nb * g.Nt > 256) return false;                             // TMEM: two item buffers of 256 columns
    const int box_pos = (KD == 3 ? TD + 2 : 1) * g.R2 * g.P;
    if (g.R2 > 256 || TD + 2 > 256) return false;
    g.a_box = box_pos * g.cpp * 4;
    const int need = g.nb * 128 + ((KD == 3 ? 2 * g.R2 : 0) + 2) * g.P + 2;      // positions an accumulator block may touch
    g.a_bytes = ((need > box_pos ? need : box_pos) * g.cpp * 4 + 1023) & ~1023;
    const int stage = 128 * g.Nt * 4;
    g.NW = 0;
    for (int nw = 3; nw >= 2; --nw) {
        const int bytes = 1024 + 2 * g.a_bytes + nw * g.w_bytes + stage;
        if (bytes <= 222 * 1024) { g.NW = nw; g.smem = bytes; break; }
    }
    return g.NW != 0;
}

// dgrad = 0: A = [src0|src1], columns = cout.  dgrad = 1: A = dy (cout channels), columns = c0 + c1 (two destinations).
bool bgeometry(const b200_conv_desc* d, int dgrad, BGeo& g) {
    const bool is3 = d->kd == 3 && d->pd == 1 && d->id >= 1;
    const bool is2 = d->kd == 1 && d->pd == 0 && d->id == 1;
    if ((!is2 && !is3) || d->kh != 3 || d->kw != 3 || d->stride != 1 || d->ph != 1 || d->pw != 1) return false;
    if (d->n < 1 || d->ih < 1 || d->iw < 4 || d->iw > 96) return false;
    const int a0 = dgrad ? d->cout : d->c0, a1 = dgrad ? 0 : d->c1;
    const int n0 = dgrad ? d->c0 : d->cout, n1 = dgrad ? d->c1 : 0;
    if (a0 <= 0 || a0 % 16 != 0 || a1 % 16 != 0) return false;
    g.cpp = (a0 % 32 == 0 && a1 % 32 == 0) ? 32 : 16;               // 16-channel tensors: 64-byte halo rows
    if (g.cpp == 16 && !is3) return false;                           // (2D 16-channel layers: conv_row's pixel-pair mode)
    g.KD = is3 ? 3 : 1;
    g.NP0 = a0 / g.cpp; g.NP = (a0 + a1) / g.cpp;
    g.Ntot = n0 + n1; g.D0 = n0; g.D1 = n1;
    if (g.Ntot % 16 != 0 || (g.Ntot % 32 != 0 && !is3)) return false;
    g.Nt = g.Ntot % 64 == 0 && (n1 == 0 || n0 % 64 == 0) ? 64 : (g.Ntot % 32 == 0 && (n1 == 0 || n0 % 32 == 0) ? 32 : 16);
    if (n1 != 0 && n0 % g.Nt != 0) return false;
    g.ntiles_n = g.Ntot / g.Nt;
    g.P = d->iw + 2;
    g.w_bytes = 3 * g.Nt * g.cpp * 4;
    BGeo best = g;
    double best_cost = -1;
    if (!is3) {
        // rows per item: best fraction of real outputs among the accumulator rows, at most 3 blocks
        for (int tr = 1; tr <= d->ih; ++tr) {
            BGeo c = g;
            if (!bfit(c, 1, d->iw, 1, tr) || c.nb > MAX_NB) break;
            const int tiles = (d->ih + tr - 1) / tr;
            const double cost = (double)tiles * c.nb * 128 / ((double)d->ih * d->iw);      // accumulator rows per real output
            if (best_cost < 0 || cost < best_cost - 1e-9) { best_cost = cost; best = c; }
        }
    } else {
        // cost model per item, calibrated on tools/bench_conv3d.py sweeps: tensor-pipe cycles of its accumulator blocks
        // (27 taps x 4 k-steps per plane; ~55 / 58 cycles per N = 32 / 64 instruction as issued here) against the TMA
        // bytes of halo and weight chunks (~26 B / cycle / SM when they hit L2), plus a fixed hand-off per item; items
        // run in waves over the SMs
        const double cyc = g.Nt == 64 ? 58.0 : (g.Nt == 32 ? 55.0 : 50.0), sms = (double)b200_num_sms();
        const char* etd = getenv("B200_BLK_TD");                     // tuning override: fixed (TD, TR)
        const char* etr = getenv("B200_BLK_TR");
        for (int td = 1; td <= d->id && td <= 16; ++td)
            for (int tr = 1; tr <= d->ih; ++tr) {
                BGeo c = g;
                if (!bfit(c, 3, d->iw, td, tr)) break;
                if (etd && etr && (td != atoi(etd) || tr != atoi(etr))) continue;
                const double mma = (double)c.nb * 27 * (c.cpp / 8) * c.NP * cyc;
                const double tma = (double)c.NP * (c.a_box + 9.0 * c.w_bytes) / 26.0;
                const double items = (double)d->n * ((d->id + td - 1) / td) * ((d->ih + tr - 1) / tr) * c.ntiles_n;
                const double waves = items <= sms ? 1.0 : items / sms;
                const double cost = waves * ((mma > tma ? mma : tma) + 3000.0);
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = c; }
            }
    }
    if (best_cost < 0) return false;
    g = best;
    g.tiles_r = (d->ih + g.TR - 1) / g.TR;
    g.tiles_d = is3 ? (d->id + g.TD - 1) / g.TD : 1;
    return true;
}

}  // namespace

B200_API int b200_conv_blk_supported(const b200_conv_desc* d, int dgrad) {
    BGeo g;
    if (!d || !bgeometry(d, dgrad, g)) return 0;
    return 8 + (dgrad ? 1 : 0) + (g.cpp == 16 ? 2 : 0);              // 8 + weight-pack mode (conv_row_pack.cuh)
}

B200_API long long b200_conv_blk_stats_blocks(const b200_conv_desc* d) {
    BGeo g;
    if (!d || !bgeometry(d, 0, g)) return 0;
    const long long items = (long long)d->n * g.tiles_d * g.tiles_r * g.ntiles_n;
    return items < b200_num_sms() ? items : b200_num_sms();         // one partial per CTA
}

B200_API int b200_conv_blk_pack_weights(const float* w, float* out, int mode, int O, int I, int taps, cudaStream_t st) {
    B200_REQUIRE(w && out && O > 0 && I > 0 && mode >= 0 && mode < 4 && ((mode & 1) ? O : I) % ((mode & 2) ? 16 : 32) == 0 &&
                 (taps == 9 || taps == 27), "conv_blk_pack_weights: bad arguments");
    const int total = taps * O * I;
    conv_blk_pack_kernel<<<(total + 255) / 256 < 64 ? (total + 255) / 256 : 64, 256, 0, st>>>(w, out, mode, O, I, taps, total);
    B200_CHECK_LAUNCH("conv_blk_pack_weights");
    return B200_OK;
}

static int run_blk(const b200_conv_desc* d, int dgrad, const float* a0, const float* a1, const float* wpk, const float* bias,
                   float* dst0, float* dst1, double* stats, int accumulate, cudaStream_t st, const char* who) {
    BGeo g;
    B200_REQUIRE(d && bgeometry(d, dgrad, g), "%s: unsupported convolution", who);
    const int N = d->n, D = d->id, H = d->ih, W = d->iw;
    const int ca0 = dgrad ? d->cout : d->c0, ca1 = dgrad ? 0 : d->c1;
    B200_REQUIRE(a0 && wpk && dst0 && (ca1 == 0 || a1) && (g.D1 == 0 || dst1), "%s: null pointer", who);
    B200_REQUIRE(!stats || g.Ntot <= 256, "%s: fused statistics need cout <= 256", who);
    CUtensorMap tx0, tx1, tw;
    const CUtensorMapSwizzle swz = g.cpp == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    auto make_in = [&](CUtensorMap* m, const float* base, int C) -> int {
        const cuuint64_t rowb = (cuuint64_t)W * C * 4;
        const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t strides[4] = {(cuuint64_t)C * 4, rowb, rowb * H, rowb * H * D};
        const cuuint32_t box[5] = {(cuuint32_t)g.cpp, (cuuint32_t)g.P, (cuuint32_t)g.R2, (cuuint32_t)(g.KD == 3 ? g.TD + 2 : 1), 1u};
        return make_tmap(m, base, 5, dims, strides, box, swz, who);
    };
    if (int rc = make_in(&tx0, a0, ca0)) return rc;
    tx1 = tx0;
    if (ca1) if (int rc = make_in(&tx1, a1, ca1)) return rc;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)g.cpp, (cuuint64_t)(9 * g.KD * g.NP * g.Ntot)};
        const cuuint64_t strides[1] = {(cuuint64_t)g.cpp * 4};
        const cuuint32_t box[2] = {(cuuint32_t)g.cpp, (cuuint32_t)g.Nt};
        if (int rc = make_tmap(&tw, wpk, 2, dims, strides, box, swz, who)) return rc;
    }
    BlkP p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.D = D; p.H = H; p.W = W; p.P = g.P; p.KD = g.KD; p.TD = g.TD; p.tiles_d = g.tiles_d; p.R2 = g.R2;
    p.TR = g.TR; p.tiles_r = g.tiles_r; p.nb = g.nb; p.NP = g.NP; p.NP0 = g.NP0;
    p.Nt = g.Nt; p.ntiles_n = g.ntiles_n; p.Ntot = g.Ntot; p.a_bytes = g.a_bytes; p.a_box_bytes = g.a_box; p.w_bytes = g.w_bytes; p.NW = g.NW;
    p.bias = bias; p.dst0 = dst0; p.dst1 = dst1; p.D0 = g.D0; p.D1 = g.D1; p.accumulate = accumulate; p.stats = stats;
    p.fdP.init(g.P);
    p.fdR2.init(g.R2);
    { const char* e = getenv("B200_BLK_DEBUG"); p.debug = e ? atoi(e) : 0; }
    const long long items = (long long)N * g.tiles_d * g.tiles_r * g.ntiles_n;
    const int grid = (int)(items < b200_num_sms() ? items : b200_num_sms());
    if (g.cpp == 32) {
        static int attr = 0;
        if (g.smem > attr) { cudaFuncSetAttribute(conv_blk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem); attr = g.smem; }
        conv_blk_kernel<32><<<grid, CB_THREADS, g.smem, st>>>(tx0, tx1, tw, p);
    } else {
        static int attr = 0;
        if (g.smem > attr) { cudaFuncSetAttribute(conv_blk_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem); attr = g.smem; }
        conv_blk_kernel<16><<<grid, CB_THREADS, g.smem, st>>>(tx0, tx1, tw, p);
    }
    B200_CHECK_LAUNCH(who);
    return B200_OK;
}

B200_API int b200_conv_blk_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wpk, const float* bias,
                               float* dst, double* stats_part, cudaStream_t st) {
    return run_blk(d, 0, src0, src1, wpk, bias, dst, nullptr, stats_part, 0, st, "conv_blk_fwd");
}

B200_API int b200_conv_blk_dgrad(const b200_conv_desc* d, const float* dy, const float* wpk_dgrad, float* dx0, float* dx1,
                                 int accumulate, cudaStream_t st) {
    return run_blk(d, 1, dy, nullptr, wpk_dgrad, nullptr, dx0, dx1, nullptr, accumulate, st, "conv_blk_dgrad");
}
