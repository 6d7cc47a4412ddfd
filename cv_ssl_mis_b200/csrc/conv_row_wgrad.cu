// Weight gradient of the 2D 3x3 stride-1 pad-1 convolutions (backward of code/networks/unet.py:37,41) on the
// 5th-generation tensor cores: tcgen05.mma kind::tf32, both operands MN-major straight from the channels-last
// activations (TMA, 128-byte swizzle of 32-byte atoms), accumulators in TMEM.
//
//     dW[kh][kw][ci][co] = sum over (n, h, w) of  dy[n][h][w][co] * x[n][h+kh-1][w+kw-1][ci]
//
// Formulation ("row ring").  A shared-memory "k-row" is 128 bytes: one pixel of a 32-channel group, or a PAIR of
// pixels of a 16-channel tensor (ppr = 2).  An image row is staged by ONE TMA box as P k-rows (one zero k-row on the
// left, >= 1 on the right: out-of-bounds box elements are zero-filled = the convolution padding; rows -1 and H are
// entirely out of bounds = zero rows).  The reduction index of the MMAs is the k-row, so
//   * the B operand (x, N = channels) of tap kw is the SAME shared-memory row shifted by kw-1 k-rows.  Because the UMMA
//     swizzle is a function of the absolute shared-memory address, a shift is just another descriptor start address;
//     with the "leading byte offset" (distance between 32-column groups) set to ONE k-row (128 B) the three taps
//     kw = 0,1,2 become the three column groups of one N = 96 MMA;
//   * the A operand (dy, M = channels) of tap kh is the dy row h = X+1-kh for the x row X being processed: the three
//     dy rows of the rolling window sit in consecutive ring slots, so with LBO = plane stride they become the row groups
//     of one M = 128 MMA (3 x 32 channels; 2 x 64 + 1 x 64; or three M = 128 MMAs for 128 channels).
// One x row therefore costs P/8 k-steps x (1..3) MMAs for ALL nine taps, x is read from HBM once and dy once per
// 32-channel group of x.  For 16-channel tensors the pixel pairs give accumulator blocks (pa, co) x (pixel offset, ci);
// the epilogue picks the blocks whose offset difference is a tap (the rest is discarded work on an HBM-bound layer).
//
//   warp 0     TMA producer (one lane): dy rows into a ring of RD slots (+2 mirror slots so that every window of three
//              consecutive rows is contiguous), x rows into a ring of RX slots; mbarrier expect_tx
//   warp 1     MMA issuer (one lane) + TMEM allocation; tcgen05.commit releases the slots
//   warps 2-5  epilogue: tcgen05.ld -> partial dW of this (row range, ci group, co block) into the workspace
// A second kernel reduces the row-range partials in fixed order (deterministic) into the framework layout [co][ci][3][3].
//
// 3D (3x3x3, backward of code/networks/vnet.py:28 and of the UNETR residual blocks): for a fixed depth tap kd the sum
// over (n, d) of the 2D problem between x plane d + kd - 1 and dy plane d, so kd joins the work-item index (ci group,
// co block, kd, row range); the "images" of the row walk are the N * D planes of x, the dy rows come from the plane one
// depth tap away (planes -1 and D are out of bounds of the tensor map = zero).
#include "umma_common.cuh"
#include <cstdlib>
#include <cstring>
#include "../../include/b200ssl.h"

namespace {

using namespace umma;

constexpr int WG_THREADS = 192;
constexpr int MAX_RX = 8, MAX_RD = 8;

struct WgP {
    int N, H, P;             // images (2D) or planes N * D (3D), rows per image, k-rows per staged image row (multiple of 8)
    int D, KD;               // planes per volume (1 in 2D), depth taps (1 or 3)
    int ppr;                 // pixels per k-row: 1 (channel counts multiple of 32) or 2 (16-channel tensors)
    int G, G0;               // x channel groups in total / served by the first source
    int MB, MG;              // co blocks (128 channels), 128-byte planes per dy slot
    int KHM, NMMA;           // window rows stacked per MMA, MMAs per k-step
    int S;                   // row-range splits
    int RX, RD;              // ring slots
    int xs_bytes, ds_bytes;  // slot sizes
    int debug;               // profiling only: bit 0 = no MMAs (TMA pipeline alone), bit 1 = no TMA (MMA pipeline alone)
    int Cin, Cout;
    float* ws;
};

__global__ void __launch_bounds__(WG_THREADS, 1) conv_row_wgrad_kernel(const __grid_constant__ CUtensorMap tx0,
                                                                       const __grid_constant__ CUtensorMap tx1,
                                                                       const __grid_constant__ CUtensorMap tdy, const WgP p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_x[MAX_RX], empty_x[MAX_RX], full_d[MAX_RD], empty_d[MAX_RD], acc_full, acc_empty;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (role branches on the uniform datapath)
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t d_base = smem0 + 1024;                                       // [guard][dy ring + 2 mirrors][guard][x ring][guard]
    const uint32_t x_base = d_base + (uint32_t)(p.RD + 2) * p.ds_bytes + 1024;
    const uint32_t total = x_base + (uint32_t)p.RX * p.xs_bytes + 1024 - smem0;
    const int R = p.N * p.H;
    const int nitems = p.G * p.MB * p.S * p.KD;

    // every byte an MMA may touch (guards, the row groups beyond the real channels) must hold finite data
    for (uint32_t o = (uint32_t)tid * 16; o < total; o += WG_THREADS * 16)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem0 + o), "r"(0u) : "memory");
    fence_proxy_async();
    if (tid == 0) {
        for (int s = 0; s < p.RX; ++s) { mbar_init(smem_u32(&full_x[s]), 1); mbar_init(smem_u32(&empty_x[s]), 1); }
        for (int s = 0; s < p.RD; ++s) { mbar_init(smem_u32(&full_d[s]), 1); mbar_init(smem_u32(&empty_d[s]), 1); }
        mbar_init(smem_u32(&acc_full), 1);
        mbar_init(smem_u32(&acc_empty), 4);
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        // whole warp in warp-uniform control flow, one elected lane issues (see the MMA issuer below)
        if (!(p.debug & 2)) {
            const bool leader = elect_one();
            if (leader) { tma_prefetch_desc(&tx0); tma_prefetch_desc(&tx1); tma_prefetch_desc(&tdy); }
            int xs = 0, xph = 0, xcnt = 0;            // x ring slot / parity of the fill being overwritten / fills so far
            int ds = 0, dph = 0, dcnt = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int g = item % p.G, t = item / p.G, mb = t % p.MB, t2 = t / p.MB, kd = t2 % p.KD, s = t2 / p.KD;
                const int r0 = (int)((long long)s * R / p.S), r1 = (int)((long long)(s + 1) * R / p.S);
                const bool second = g >= p.G0;
                const int gl = second ? g - p.G0 : g;
                const int img = r0 / p.H;
                int n = img / p.D, xd = img - n * p.D, X = r0 - img * p.H;
                const int dshift = p.KD == 3 ? 1 - kd : 0;          // dy plane = x plane + 1 - kd
                const uint32_t plane_b = (uint32_t)p.P * 128;
                for (int r = r0; r < r1; ++r) {
                    const bool new_seg = (r == r0) || (X == 0);
                    const int first = new_seg ? X - 1 : X + 1, cnt = new_seg ? 3 : 1;
                    for (int k = 0; k < cnt; ++k, ++dcnt) {
                        if (dcnt >= p.RD) mbar_wait(smem_u32(&empty_d[ds]), dph ^ 1);
                        const uint32_t fb = smem_u32(&full_d[ds]);
                        if (leader) {
                            mbar_expect_tx(fb, (uint32_t)p.ds_bytes * (ds < 2 ? 2u : 1u));
                            for (int j = 0; j < p.MG; ++j) {
                                const int c0 = 32 * (mb * p.MG + j);
                                tma_load_5d(d_base + (uint32_t)ds * p.ds_bytes + (uint32_t)j * plane_b, &tdy, c0, -1, first + k, xd + dshift, n, fb);
                                if (ds < 2)
                                    tma_load_5d(d_base + (uint32_t)(p.RD + ds) * p.ds_bytes + (uint32_t)j * plane_b, &tdy, c0, -1, first + k,
                                                xd + dshift, n, fb);
                            }
                        }
                        __syncwarp();
                        if (++ds == p.RD) { ds = 0; dph ^= 1; }
                    }
                    if (xcnt >= p.RX) mbar_wait(smem_u32(&empty_x[xs]), xph ^ 1);
                    const uint32_t fb = smem_u32(&full_x[xs]);
                    if (leader) {
                        mbar_expect_tx(fb, (uint32_t)p.xs_bytes);
                        if (second) tma_load_5d(x_base + (uint32_t)xs * p.xs_bytes, &tx1, 32 * gl, -1, X, xd, n, fb);
                        else tma_load_5d(x_base + (uint32_t)xs * p.xs_bytes, &tx0, 32 * gl, -1, X, xd, n, fb);
                    }
                    __syncwarp();
                    ++xcnt;
                    if (++xs == p.RX) { xs = 0; xph ^= 1; }
                    if (++X == p.H) { X = 0; if (++xd == p.D) { xd = 0; ++n; } }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // One thread feeds the tensor core, so the loop body must stay a handful of instructions: the descriptors of a
        // row are built once and a k-step only adds 1024 B (64 units of 16 B) to their 14-bit address field
        // (tools/umma_peak.cu: an N = 96 TF32 MMA retires every 56 cycles; a 40-instruction loop body issued one every 250).
        // The WHOLE warp runs the loop (warp-uniform control flow keeps descriptors in uniform registers; under
        // `if (lane == 0)` the compiler wraps every tcgen05.mma in a uniformisation loop); one elected lane issues.
        {
            const bool leader = elect_one();
            const uint32_t idesc96 = idesc_tf32(128, 96, 1, 1);
            const uint32_t plane = (uint32_t)p.P * 128;
            const int ksteps = p.P / 8;
            const uint64_t a_inc = (uint64_t)(((uint32_t)p.KHM * (uint32_t)p.ds_bytes) >> 4);
            const uint32_t t1 = tmem_base + 128, t2 = tmem_base + 256;
            // ring positions and mbarrier phases are carried as counters (no divisions on the issue path)
            int xslot = 0, xph = 0;                 // x ring: slot / parity of the row being consumed
            int wslot = 0;                          // dy ring: slot of the oldest row of the window
            int fslot = 0, fph = 0, ahead = 0;      // dy ring: next slot to wait for / its parity / rows already waited beyond wslot
            int il = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++il) {
                const int s = item / (p.G * p.MB * p.KD);
                const int r0 = (int)((long long)s * R / p.S), r1 = (int)((long long)(s + 1) * R / p.S);
                if (il >= 1) mbar_wait(smem_u32(&acc_empty), (il - 1) & 1);
                tc_fence_after();
                uint32_t acc = 0;
                int X = r0 % p.H;
                for (int r = r0; r < r1; ++r) {
                    if (!(p.debug & 2)) {
                        for (; ahead < 3; ++ahead) {          // the window [wslot, wslot + 3) must have landed
                            mbar_wait(smem_u32(&full_d[fslot]), fph);
                            if (++fslot == p.RD) { fslot = 0; fph ^= 1; }
                        }
                        mbar_wait(smem_u32(&full_x[xslot]), xph);
                    }
                    tc_fence_after();
                    uint64_t ad = smem_desc(d_base + (uint32_t)wslot * p.ds_bytes, plane, 512, 1);
                    uint64_t bd = smem_desc(x_base + (uint32_t)xslot * p.xs_bytes - 128, 128, 512, 1);
                    const int nk = (p.debug & 1) ? 0 : ksteps;
                    if (p.NMMA == 1) {
#pragma unroll 4
                        for (int ks = 0; ks < nk; ++ks, ad += 64, bd += 64) {
                            if (leader) mma_tf32(tmem_base, ad, bd, idesc96, acc);
                            acc = 1;
                        }
                    } else if (p.NMMA == 2) {
#pragma unroll 4
                        for (int ks = 0; ks < nk; ++ks, ad += 64, bd += 64) {
                            if (leader) mma_tf32(tmem_base, ad, bd, idesc96, acc);
                            if (leader) mma_tf32(t1, ad + a_inc, bd, idesc96, acc);
                            acc = 1;
                        }
                    } else {
#pragma unroll 4
                        for (int ks = 0; ks < nk; ++ks, ad += 64, bd += 64) {
                            if (leader) mma_tf32(tmem_base, ad, bd, idesc96, acc);
                            if (leader) mma_tf32(t1, ad + a_inc, bd, idesc96, acc);
                            if (leader) mma_tf32(t2, ad + 2 * a_inc, bd, idesc96, acc);
                            acc = 1;
                        }
                    }
                    const bool last = (X == p.H - 1) || (r == r1 - 1);
                    const int w1 = wslot + 1 == p.RD ? 0 : wslot + 1, w2 = w1 + 1 == p.RD ? 0 : w1 + 1;
                    if (leader) {
                        mma_commit(smem_u32(&empty_x[xslot]));
                        mma_commit(smem_u32(&empty_d[wslot]));
                        if (last) {
                            mma_commit(smem_u32(&empty_d[w1]));
                            mma_commit(smem_u32(&empty_d[w2]));
                        }
                    }
                    __syncwarp();
                    if (++xslot == p.RX) { xslot = 0; xph ^= 1; }
                    if (last) { wslot = w2 + 1 == p.RD ? 0 : w2 + 1; ahead = 0; }     // next segment starts a fresh window
                    else { wslot = w1; ahead = 2; }
                    X = (X == p.H - 1) ? 0 : X + 1;
                }
                if (leader) mma_commit(smem_u32(&acc_full));
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (4 warps, one TMEM lane quarter each)
        const int q = warp & 3;
        int il = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++il) {
            const int g = item % p.G, t = item / p.G, mb = t % p.MB, t2 = t / p.MB, kd = t2 % p.KD, s = t2 / p.KD;
            const int T = 9 * p.KD, tap0 = kd * 9;
            mbar_wait(smem_u32(&acc_full), il & 1);
            tc_fence_after();
            for (int a = 0; a < p.NMMA; ++a) {
                const int jw = q / p.MG, pl = q % p.MG;            // window row and channel plane of this lane quarter
                const int j = a * p.KHM + jw;
                const bool row_ok = jw < p.KHM && j < 3;
                const int kh = 2 - j;
#pragma unroll 1
                for (int c = 0; c < 3; ++c) {
                    uint32_t rg[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 128 + c * 32), rg);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    if (p.ppr == 1) {
                        const int co = mb * 128 + pl * 32 + lane;
                        if (co < p.Cout) {
                            float* dst = p.ws + ((size_t)((s * T + tap0 + kh * 3 + c) * p.Cin + g * 32)) * p.Cout + co;
#pragma unroll
                            for (int i = 0; i < 32; ++i) dst[(size_t)i * p.Cout] = __uint_as_float(rg[i]);
                        }
                    } else {
                        const int pa = lane >> 4, co = lane & 15;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kw = 2 * c + (i >> 4) - 1 - pa;            // pixel offset of this column block minus pa
                            if (kw >= 0 && kw <= 2)
                                p.ws[((size_t)(((s * 2 + pa) * T + tap0 + kh * 3 + kw) * p.Cin + g * 16 + (i & 15))) * p.Cout + co] =
                                    __uint_as_float(rg[i]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dw[co][ci][tap] (+)= sum over partials of ws[s][tap][ci][co]: a block covers 64 float4 columns (co fastest) with 4
// interleaved partial sums each, combined in a fixed order (deterministic).  db_zero: bias gradient of a convolution that
// feeds a train-mode BatchNorm -- identically zero (the BatchNorm backward removes the per-channel mean of its gradient).
__global__ void __launch_bounds__(256) conv_row_wgrad_reduce_kernel(const float* __restrict__ ws, int nparts, int Cin, int Cout, int T,
                                                                    float* __restrict__ dw, int accumulate, float* __restrict__ db_zero) {
    __shared__ float4 part[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int E = T * Cin * Cout;
    const int e = (blockIdx.x * 64 + tx) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < E) {
#pragma unroll 4
        for (int s = ty; s < nparts; s += 4) {
            const float4 v = ldg4(ws + (size_t)s * E + e);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && e < E) {
        float4 v = part[0][tx];
#pragma unroll
        for (int k = 1; k < 4; ++k) { v.x += part[k][tx].x; v.y += part[k][tx].y; v.z += part[k][tx].z; v.w += part[k][tx].w; }
        const int co = e % Cout, r = e / Cout, ci = r % Cin, tap = r / Cin;
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float* o = dw + ((size_t)(co + k) * Cin + ci) * T + tap;
            *o = accumulate ? *o + vv[k] : vv[k];
        }
    }
    if (db_zero && !accumulate && blockIdx.x == 0)
        for (int c = threadIdx.x; c < Cout; c += 256) db_zero[c] = 0.f;
}

struct Geo {
    int ppr, Wk, P, G, G0, MB, MG, KHM, NMMA, S, RX, RD, xs, ds, smem, KD;
};

bool geometry(const b200_conv_desc* d, Geo& g) {
    const bool is3 = d->kd == 3 && d->pd == 1 && d->id >= 1, is2 = d->kd == 1 && d->pd == 0 && d->id == 1;
    if ((!is2 && !is3) || d->kh != 3 || d->kw != 3 || d->stride != 1 || d->ph != 1 || d->pw != 1) return false;
    g.KD = is3 ? 3 : 1;
    const int c0 = d->c0, c1 = d->c1, co = d->cout;
    if (d->n < 1 || d->ih < 1 || d->iw < 1) return false;
    if (c0 % 32 == 0 && c1 % 32 == 0 && c0 > 0 && co % 32 == 0 && (co <= 128 || co % 128 == 0)) {
        g.ppr = 1;
        g.G0 = c0 / 32; g.G = (c0 + c1) / 32;
        g.MB = (co + 127) / 128;
        g.MG = (co < 128 ? co : 128) / 32;
    } else if (c0 == 16 && (c1 == 0 || c1 == 16) && co == 16 && d->iw % 2 == 0) {
        g.ppr = 2;
        g.G0 = 1; g.G = c1 ? 2 : 1;
        g.MB = 1; g.MG = 1;
    } else {
        return false;
    }
    g.Wk = d->iw / g.ppr;
    g.P = (g.Wk + 2 + 7) / 8 * 8;
    if (g.P > 256) return false;
    g.KHM = g.MG == 1 ? 3 : (g.MG == 2 ? 2 : 1);
    if (g.MG == 3) { g.KHM = 1; }                         // 96 channels: one window row per MMA
    const char* e = getenv("B200_WGRAD_KHM1");
    if (e && e[0] == '1') g.KHM = 1;
    g.NMMA = (3 + g.KHM - 1) / g.KHM;
    g.xs = g.P * 128;
    g.ds = g.MG * g.P * 128;
    g.RD = 0;
    for (int rd = MAX_RD; rd >= 4; --rd) {
        const int rx = rd - 2 < 3 ? 3 : rd - 2;
        const int bytes = 1024 + 1024 + (rd + 2) * g.ds + 1024 + rx * g.xs + 1024;
        if (bytes <= 224 * 1024) { g.RD = rd; g.RX = rx; g.smem = bytes; break; }
    }
    if (!g.RD) return false;
    // the row groups beyond the stacked window read up to 4 planes past the last mirror slot: they must stay inside the x ring
    if ((long long)g.RX * g.xs + 1024 < 4ll * g.P * 128) return false;
    const long long R = (long long)d->n * d->id * d->ih;
    long long S = b200_num_sms() / ((long long)g.G * g.MB * g.KD);
    if (S > R) S = R;
    if (S < 1) S = 1;
    g.S = (int)S;
    return true;
}

}  // namespace

B200_API int b200_conv_row_wgrad_supported(const b200_conv_desc* d) {
    Geo g;
    return (d && geometry(d, g)) ? 1 : 0;
}

B200_API long long b200_conv_row_wgrad_workspace_bytes(const b200_conv_desc* d) {
    Geo g;
    if (!d || !geometry(d, g)) return 0;
    return (long long)g.S * g.ppr * 9 * g.KD * (d->c0 + d->c1) * d->cout * (long long)sizeof(float);
}

B200_API int b200_conv_row_wgrad(const b200_conv_desc* d, const float* src0, const float* src1, const float* dy, float* workspace,
                                 long long workspace_bytes, float* dw, float* db_zero, int accumulate, cudaStream_t st) {
    Geo g;
    B200_REQUIRE(d && geometry(d, g), "conv_row_wgrad: unsupported convolution");
    B200_REQUIRE(src0 && dy && dw && workspace && (d->c1 == 0 || src1), "conv_row_wgrad: null pointer");
    if (workspace_bytes < b200_conv_row_wgrad_workspace_bytes(d)) {
        b200_set_error("conv_row_wgrad: workspace too small (%lld < %lld bytes)", workspace_bytes, b200_conv_row_wgrad_workspace_bytes(d));
        return B200_ERR_WORKSPACE;
    }
    const int N = d->n, D = d->id, H = d->ih, W = d->iw, Cin = d->c0 + d->c1, Cout = d->cout;
    CUtensorMap tx0, tx1, tdy;
    auto make = [&](CUtensorMap* m, const float* base, int C) -> int {
        // channels-last [N][D][H][W][C] viewed as {channels (a box takes one 128-byte group), k-rows of an image row, H, D, N};
        // 16-channel tensors: the 128-byte k-row is a pixel pair
        const cuuint64_t rowb = (cuuint64_t)W * C * 4;
        const cuuint64_t dims[5] = {g.ppr == 1 ? (cuuint64_t)C : 32u, (cuuint64_t)g.Wk, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t strides[4] = {g.ppr == 1 ? (cuuint64_t)C * 4 : 128u, rowb, rowb * H, rowb * H * D};
        const cuuint32_t box[5] = {32u, (cuuint32_t)g.P, 1u, 1u, 1u};
        return make_tmap(m, base, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "conv_row_wgrad");
    };
    if (int rc = make(&tx0, src0, d->c0)) return rc;
    tx1 = tx0;
    if (d->c1) if (int rc = make(&tx1, src1, d->c1)) return rc;
    if (int rc = make(&tdy, dy, Cout)) return rc;

    WgP p;
    memset(&p, 0, sizeof(p));
    p.N = N * D; p.D = D; p.KD = g.KD; p.H = H; p.P = g.P; p.ppr = g.ppr; p.G = g.G; p.G0 = g.G0; p.MB = g.MB; p.MG = g.MG; p.KHM = g.KHM; p.NMMA = g.NMMA;
    p.S = g.S; p.RX = g.RX; p.RD = g.RD; p.xs_bytes = g.xs; p.ds_bytes = g.ds; p.Cin = Cin; p.Cout = Cout; p.ws = workspace;
    const char* e = getenv("B200_WGRAD_DEBUG");
    p.debug = e ? atoi(e) : 0;
    static int attr_bytes = 0;
    if (g.smem > attr_bytes) {
        cudaFuncSetAttribute(conv_row_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem);
        attr_bytes = g.smem;
    }
    const int nitems = g.G * g.MB * g.S * g.KD;
    const int grid = nitems < b200_num_sms() ? nitems : b200_num_sms();
    conv_row_wgrad_kernel<<<grid, WG_THREADS, g.smem, st>>>(tx0, tx1, tdy, p);
    B200_CHECK_LAUNCH("conv_row_wgrad");
    const int total = 9 * g.KD * Cin * Cout;
    conv_row_wgrad_reduce_kernel<<<(total / 4 + 63) / 64, 256, 0, st>>>(workspace, g.S * g.ppr, Cin, Cout, 9 * g.KD, dw, accumulate, db_zero);
    B200_CHECK_LAUNCH("conv_row_wgrad_reduce");
    return B200_OK;
}
