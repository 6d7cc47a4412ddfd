// Tile-based 3x3 / 3x3x3 stride-1 "same" convolution on the tensor cores (TF32 mma.sync, fp32 accumulate):
// forward, data gradient (same kernel, flipped weights) and weight gradient.
//
// Unlike the generic implicit GEMM in conv.cu, the im2col operand is never gathered: each CTA stages one
// spatial halo tile of 16 input channels in shared memory ONCE per channel chunk (cp.async, zero-filled at the
// image border) and every tap reads its shifted view of that tile straight into MMA fragments with ldmatrix.
// HBM traffic is therefore ~1.27x the input (halo overlap) instead of 9x/27x through L1/L2, there is no address
// arithmetic in the inner loop, and the weights arrive pre-rounded to TF32 in the exact order ldmatrix wants.
//
// Layout facts used throughout (all measured/derived in DESIGN.md):
//   * shared-memory rows of 16 fp32 channels are padded to 20 floats: 8 consecutive rows x 16 B then cover all
//     32 banks, so every ldmatrix phase is conflict-free;
//   * one ldmatrix.x4 of 8x8 b16 matrices == one m16k8 TF32 A fragment (rows 0-7/8-15 x k 0-3/4-7);
//   * B fragments come from weights stored [tap][cout][16 cin] (cin contiguous): one ldmatrix.x4 yields b0,b1
//     of both k8 steps of a 16-channel chunk.
#include "conv_common.cuh"
#include "../../include/b200ssl.h"

#define KC 16                 // input channels per chunk
#define PSTR 20               // padded shared-memory row length (floats) of a KC-channel pixel / weight row

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const float* p) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}

struct TileP {
    const float* src0;      // input (fwd) / dy (dgrad) / x (wgrad), channels-last
    const float* src1;      // second concat source or null
    int C0, C1, Cin;        // Cin = C0 + C1 (channels of the staged operand)
    int N, D, H, W;         // spatial dims (D = 1 for 2D); 'same' conv: output dims are identical
    int tiles_d, tiles_h, tiles_w;
    const float* wt;        // packed weights [chunks][taps][CoutP][16], TF32-rounded
    int Cout, CoutP;        // output channels, padded to a multiple of 16 in the packing
    const float* bias;
    float* dst0;
    float* dst1;
    int D0, D1;             // output channel split [dst0 | dst1]
    int out_nchw;
    int accumulate;
    // wgrad only
    const float* g;         // dy [pixels][NG]
    int NG;
    float* part;            // [splits][taps*Cin][NG]
    float* part_colsum;     // [splits][NG] or null
    int tiles_per_split;
};

template <int DIMS>
struct TileShape {
    static constexpr int TD = DIMS == 3 ? 4 : 1, TH = DIMS == 3 ? 8 : 16, TW = DIMS == 3 ? 8 : 16;
    static constexpr int HD = DIMS == 3 ? TD + 2 : 1, HH = TH + 2, HW = TW + 2;
    static constexpr int PIX = TD * TH * TW;            // 256 output pixels per tile
    static constexpr int HPIX = HD * HH * HW;           // halo pixels
    static constexpr int KD = DIMS == 3 ? 3 : 1;
    static constexpr int TAPS2 = 9;                     // taps per kd plane
    static constexpr int TAPS = KD * 9;
};

// tile-linear pixel index -> halo offset (in pixels) of the tap-(0,0,0) input pixel
template <int DIMS>
__device__ __forceinline__ int pix_to_halo(int p) {
    using S = TileShape<DIMS>;
    const int w = p % S::TW, h = (p / S::TW) % S::TH, d = p / (S::TW * S::TH);
    return (d * S::HH + h) * S::HW + w;
}

// stage the halo tile of channels [c0, c0+16) into shared memory (cp.async, zero fill outside the image / Cin)
template <int DIMS, int STRIDE = PSTR>
__device__ __forceinline__ void load_halo(const TileP& p, float* sh, int n, int d0, int h0, int w0, int c0, int tid) {
    using S = TileShape<DIMS>;
    for (int s = tid; s < S::HPIX * 4; s += 256) {
        const int hp = s >> 2, piece = s & 3;
        const int hw = hp % S::HW, hh = (hp / S::HW) % S::HH, hd = hp / (S::HW * S::HH);
        const int id = d0 + hd - (DIMS == 3 ? 1 : 0), ih = h0 + hh - 1, iw = w0 + hw - 1;
        const int ch = c0 + piece * 4;
        bool ok = (unsigned)id < (unsigned)p.D && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W && ch < p.Cin;
        const float* src = p.src0;
        if (ok) {
            const size_t pix = (((size_t)n * p.D + id) * p.H + ih) * p.W + iw;
            src = ch < p.C0 ? p.src0 + pix * p.C0 + ch : p.src1 + pix * p.C1 + (ch - p.C0);
        }
        cp_async16(sh + hp * STRIDE + piece * 4, src, ok);
    }
}

// =====================================================================================
// forward / dgrad: out[pixel][n] = bias[n] + sum_{tap, c} halo[pixel + tap][c] * W[tap][n][c]
// =====================================================================================
template <int DIMS, int BN>
struct FwdSmem {
    using S = TileShape<DIMS>;
    static constexpr int HALO_F = S::HPIX * PSTR;
    static constexpr int W_F = S::TAPS2 * BN * PSTR;     // weights of one kd plane
    static constexpr int STAGE_F = HALO_F + W_F;
    static constexpr size_t BYTES = 2 * (size_t)STAGE_F * sizeof(float);
};

template <int DIMS, int BN>
__global__ void __launch_bounds__(256, (DIMS == 2 && BN == 16) ? 4 : 1) conv_tile_fwd_kernel(const TileP p) {
    using S = TileShape<DIMS>;
    using SM = FwdSmem<DIMS, BN>;
    constexpr int NT = BN / 8;                           // n8 tiles
    extern __shared__ __align__(16) float smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    int tile = blockIdx.x;
    const int tw = tile % p.tiles_w; tile /= p.tiles_w;
    const int th = tile % p.tiles_h; tile /= p.tiles_h;
    const int td = tile % p.tiles_d;
    const int n = tile / p.tiles_d;
    const int d0 = td * S::TD, h0 = th * S::TH, w0 = tw * S::TW;
    const int n0 = blockIdx.y * BN;

    // per-lane ldmatrix row offsets (floats) inside the halo / weight stage
    int a_off[2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const int r = (lane & 7) + 8 * ((lane >> 3) & 1);
        a_off[mi] = pix_to_halo<DIMS>((warp * 2 + mi) * 16 + r) * PSTR + 4 * (lane >> 4);
    }
    const int b_off = (lane & 7) * PSTR + 4 * (lane >> 3);

    float acc[2][NT][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    const int nchunks = (p.Cin + KC - 1) / KC;
    const int nsteps = nchunks * S::KD;                  // one step = (channel chunk, kd plane)

    auto issue = [&](int step) {
        const int chunk = step / S::KD, kd = step % S::KD;
        float* st = smem + (step & 1) * SM::STAGE_F;
        // the halo of a chunk is loaded with its first kd plane and shared by the following ones
        if (kd == 0) load_halo<DIMS>(p, smem + (chunk & 1) * SM::STAGE_F, n, d0, h0, w0, chunk * KC, tid);
        float* sw = st + SM::HALO_F;
        const float* wsrc = p.wt + ((size_t)(chunk * S::TAPS + kd * S::TAPS2) * p.CoutP) * KC;
        for (int s = tid; s < S::TAPS2 * BN * 4; s += 256) {
            const int piece = s & 3, row = s >> 2;        // row = tap2 * BN + nn
            const int tap2 = row / BN, nn = row % BN;
            const bool ok = n0 + nn < p.CoutP;
            cp_async16(sw + row * PSTR + piece * 4, wsrc + ((size_t)tap2 * p.CoutP + (ok ? n0 + nn : 0)) * KC + piece * 4, ok);
        }
        cp_async_commit();
    };

    // For DIMS == 3 the halo buffer of chunk c lives in stage (c & 1) while weight planes alternate per step;
    // KD = 3 is odd, so give the halo its own double buffer by indexing it with the chunk parity (see issue()).
    issue(0);
    for (int step = 0; step < nsteps; ++step) {
        if (step + 1 < nsteps) {
            issue(step + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int chunk = step / S::KD, kd = step % S::KD;
        const float* sh = smem + (chunk & 1) * SM::STAGE_F;
        const float* sw = smem + (step & 1) * SM::STAGE_F + SM::HALO_F;
#pragma unroll
        for (int tap2 = 0; tap2 < 9; ++tap2) {
            const int kh = tap2 / 3, kw = tap2 % 3;
            const int tap_off = ((kd * S::HH + kh) * S::HW + kw) * PSTR;
            uint32_t bfr[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) ldsm4(bfr[j], sw + (tap2 * BN + j * 8) * PSTR + b_off);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t afr[2][4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) ldsm4(afr[mi], sh + a_off[mi] + tap_off + ks * 8);
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const uint32_t b2[2] = {bfr[j][2 * ks], bfr[j][2 * ks + 1]};
                        mma_tf32(acc[mi][j], afr[mi], b2);
                    }
            }
        }
        __syncthreads();
    }

    // ---------------- epilogue: bias, optional channel split / NCHW / accumulate
    const size_t S_img = (size_t)p.D * p.H * p.W;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int pt = (warp * 2 + mi) * 16 + g + 8 * hrow;
            const int w = w0 + pt % S::TW, h = h0 + (pt / S::TW) % S::TH, d = d0 + pt / (S::TW * S::TH);
            if (d >= p.D || h >= p.H || w >= p.W) continue;
            const size_t sp = ((size_t)d * p.H + h) * p.W + w;
            const size_t pix = (size_t)n * S_img + sp;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int nn = n0 + j * 8 + 2 * t;
                if (nn >= p.Cout) continue;
                float v0 = acc[mi][j][2 * hrow], v1 = acc[mi][j][2 * hrow + 1];
                if (p.bias) {
                    v0 += __ldg(p.bias + nn);
                    if (nn + 1 < p.Cout) v1 += __ldg(p.bias + nn + 1);
                }
                if (p.out_nchw) {
                    float* o = p.dst0 + ((size_t)n * p.Cout + nn) * S_img + sp;
                    o[0] = p.accumulate ? o[0] + v0 : v0;
                    if (nn + 1 < p.Cout) o[S_img] = p.accumulate ? o[S_img] + v1 : v1;
                } else {
                    // D0, D1 and nn are even, so a channel pair never straddles the split
                    float* o = nn < p.D0 ? p.dst0 + pix * p.D0 + nn : p.dst1 + pix * p.D1 + (nn - p.D0);
                    if (nn + 1 < p.Cout) {
                        float2 r = make_float2(v0, v1);
                        if (p.accumulate) { const float2 old = *reinterpret_cast<const float2*>(o); r.x += old.x; r.y += old.y; }
                        *reinterpret_cast<float2*>(o) = r;
                    } else {
                        o[0] = p.accumulate ? o[0] + v0 : v0;
                    }
                }
            }
        }
    }
}

// =====================================================================================
// wgrad: part[split][tap*Cin + c][n] = sum_{pixels of the split} halo[pixel + tap][c] * dy[pixel][n]
//   CTA = (16-channel chunk of x, BN columns of dy, [3D: one kd plane of taps,] a range of pixel tiles); warps =
//   (BN/16) column groups x pixel groups, each holding the 9 (kh, kw) taps of its 16x16 block in registers.  In 3D the
//   27 taps do not fit in registers, so blockIdx.y also enumerates kd and the CTA stages only the 4 depth planes
//   [d0 + kd - 1, d0 + kd + 3) of the halo.
// =====================================================================================
template <int DIMS, int BN>
struct WgSmem {
    using S = TileShape<DIMS>;
    static constexpr int GSTR = BN + 8;                  // dy row stride: (BN + 8) % 32 in {8, 24} -> conflict-free B frags
    static constexpr int XSTR = 24;                      // x row stride: lanes (pixel t, channel g) -> bank 24 t + g, all distinct
    static constexpr int HPIX_W = S::TD * S::HH * S::HW; // halo pixels staged per tile: the TD depth planes one kd needs
    static constexpr int HALO_F = HPIX_W * XSTR;
    static constexpr int G_F = S::PIX * GSTR;
    static constexpr int STAGE_F = HALO_F + G_F;
    static constexpr int CG = BN / 16, PG = 8 / CG;      // column groups, pixel groups
    static constexpr int RED_F = PG * S::TAPS2 * 16 * BN; // cross-pixel-group reduction buffer
    static constexpr int PIPE_F = 2 * STAGE_F;
    static constexpr size_t BYTES = (size_t)(PIPE_F > RED_F ? PIPE_F : RED_F) * sizeof(float);
};

template <int DIMS, int BN>
__global__ void __launch_bounds__(256) conv_tile_wgrad_kernel(const TileP p) {
    using S = TileShape<DIMS>;
    using SM = WgSmem<DIMS, BN>;
    constexpr int CG = SM::CG, PG = SM::PG, GSTR = SM::GSTR;
    constexpr int KSTEPS = S::PIX / 8 / PG;              // k8 (pixel) steps per warp per tile
    extern __shared__ __align__(16) float smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int cg = warp % CG, pg = warp / CG;
    const int kd = DIMS == 3 ? blockIdx.y % 3 : 0;
    const int chunk = blockIdx.x, n0 = (DIMS == 3 ? blockIdx.y / 3 : blockIdx.y) * BN;
    const int tiles_img = p.tiles_d * p.tiles_h * p.tiles_w;
    const int ntiles = p.N * tiles_img;
    const int tbeg = blockIdx.z * p.tiles_per_split;
    const int tend = min(ntiles, tbeg + p.tiles_per_split);

    float acc[S::TAPS2][2][4];
#pragma unroll
    for (int a = 0; a < S::TAPS2; ++a)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[a][j][e] = 0.f;
    float colsum = 0.f;
    const bool do_colsum = p.part_colsum != nullptr && chunk == 0 && kd == 0;

    auto issue = [&](int tile, int stage) {
        int q = tile;
        const int tw = q % p.tiles_w; q /= p.tiles_w;
        const int th = q % p.tiles_h; q /= p.tiles_h;
        const int td = q % p.tiles_d;
        const int n = q / p.tiles_d;
        const int d0 = td * S::TD, h0 = th * S::TH, w0 = tw * S::TW;
        float* st = smem + stage * SM::STAGE_F;
        if (DIMS == 2) {
            load_halo<DIMS, SM::XSTR>(p, st, n, d0, h0, w0, chunk * KC, tid);
        } else {
            // depth planes d0 + kd - 1 + [0, TD) only: local plane index hd pairs with output depth d0 + hd
            for (int s = tid; s < SM::HPIX_W * 4; s += 256) {
                const int hp = s >> 2, piece = s & 3;
                const int hw = hp % S::HW, hh = (hp / S::HW) % S::HH, hd = hp / (S::HW * S::HH);
                const int id = d0 + hd + kd - 1, ih = h0 + hh - 1, iw = w0 + hw - 1;
                const int ch = chunk * KC + piece * 4;
                const bool ok = (unsigned)id < (unsigned)p.D && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W && ch < p.Cin;
                const float* src = p.src0;
                if (ok) {
                    const size_t pix = (((size_t)n * p.D + id) * p.H + ih) * p.W + iw;
                    src = ch < p.C0 ? p.src0 + pix * p.C0 + ch : p.src1 + pix * p.C1 + (ch - p.C0);
                }
                cp_async16(st + hp * SM::XSTR + piece * 4, src, ok);
            }
        }
        float* sg = st + SM::HALO_F;
        for (int s = tid; s < S::PIX * (BN / 4); s += 256) {
            const int pt = s / (BN / 4), piece = s % (BN / 4);
            const int w = w0 + pt % S::TW, h = h0 + (pt / S::TW) % S::TH, d = d0 + pt / (S::TW * S::TH);
            const int col = n0 + piece * 4;
            const bool ok = d < p.D && h < p.H && w < p.W && col < p.NG;
            const size_t pix = (((size_t)n * p.D + d) * p.H + h) * p.W + w;
            cp_async16(sg + pt * GSTR + piece * 4, ok ? p.g + pix * p.NG + col : p.g, ok);
        }
        cp_async_commit();
    };

    if (tbeg < tend) issue(tbeg, 0);
    for (int tile = tbeg; tile < tend; ++tile) {
        const int stage = (tile - tbeg) & 1;
        if (tile + 1 < tend) {
            issue(tile + 1, stage ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* sh = smem + stage * SM::STAGE_F;
        const float* sg = sh + SM::HALO_F;
#pragma unroll 2
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const int p0 = (pg * KSTEPS + ks) * 8;          // first of 8 consecutive tile pixels (one row segment)
            const int hoff = pix_to_halo<DIMS>(p0);
            float bf[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float* b = sg + (p0 + t) * GSTR + cg * 16 + j * 8 + g;
                bf[j][0] = b[0];
                bf[j][1] = b[4 * GSTR];
            }
            uint32_t bu[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) { bu[j][0] = f2tf32(bf[j][0]); bu[j][1] = f2tf32(bf[j][1]); }
#pragma unroll
            for (int tap = 0; tap < S::TAPS2; ++tap) {
                const int kh = tap / 3, kw = tap % 3;
                const float* a = sh + (hoff + kh * S::HW + kw + t) * SM::XSTR + g;
                uint32_t au[4];
                au[0] = __float_as_uint(a[0]);
                au[1] = __float_as_uint(a[8]);
                au[2] = __float_as_uint(a[4 * SM::XSTR]);
                au[3] = __float_as_uint(a[4 * SM::XSTR + 8]);
                mma_tf32(acc[tap][0], au, bu[0]);
                mma_tf32(acc[tap][1], au, bu[1]);
            }
        }
        if (do_colsum) {
            const int c = tid % BN, sl = tid / BN;
            for (int pr = sl; pr < S::PIX; pr += 256 / BN) colsum += sg[pr * GSTR + c];
        }
        __syncthreads();
    }

    // ---------------- reduce the pixel groups, write this split's partial
    float* red = smem;                                     // [PG][9][16][BN]
#pragma unroll
    for (int tap = 0; tap < S::TAPS2; ++tap)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = g + 8 * (e >> 1), c = cg * 16 + j * 8 + 2 * t + (e & 1);
                red[((pg * S::TAPS2 + tap) * 16 + r) * BN + c] = acc[tap][j][e];
            }
    __syncthreads();
    const int K = S::TAPS * p.Cin;
    float* out = p.part + (size_t)blockIdx.z * K * p.NG;
    for (int idx = tid; idx < S::TAPS2 * 16 * BN; idx += 256) {
        const int c = idx % BN, r = (idx / BN) % 16, tap = idx / (BN * 16);
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < PG; ++q) v += red[((q * S::TAPS2 + tap) * 16 + r) * BN + c];
        const int ci = chunk * KC + r, col = n0 + c;
        if (ci < p.Cin && col < p.NG) out[((size_t)(kd * S::TAPS2 + tap) * p.Cin + ci) * p.NG + col] = v;
    }
    if (do_colsum) {
        __syncthreads();
        red[tid] = colsum;
        __syncthreads();
        if (tid < BN) {
            float v = 0.f;
            for (int sl = 0; sl < 256 / BN; ++sl) v += red[sl * BN + tid];
            if (n0 + tid < p.NG) p.part_colsum[(size_t)blockIdx.z * p.NG + n0 + tid] = v;
        }
    }
}

// =====================================================================================
// weight packing for the tile kernels: [chunks][taps][CoutP][16], TF32-rounded, zero padded
// =====================================================================================
__global__ void pack_tile_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int dgrad, int O, int I,
                                         int T, int rows, int cols, int colsP, int chunks) {
    // rows = reduction channels (I for fwd, O for dgrad), cols = produced channels (O for fwd, I for dgrad)
    const long long total = (long long)chunks * T * colsP * KC;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % KC);
        long long r = idx / KC;
        const int col = (int)(r % colsP); r /= colsP;
        const int tap = (int)(r % T);
        const int chunk = (int)(r / T);
        const int row = chunk * KC + kk;
        float v = 0.f;
        if (row < rows && col < cols) {
            v = dgrad ? w[((size_t)row * I + col) * T + (T - 1 - tap)]      // w[o = row][i = col][flipped tap]
                      : w[((size_t)col * I + row) * T + tap];               // w[o = col][i = row][tap]
            v = __uint_as_float(f2tf32(v));
        }
        out[idx] = v;
    }
}

static inline int round16(int v) { return (v + 15) / 16 * 16; }

B200_API long long b200_conv_tile_packed_floats(int dgrad, int O, int I, int T) {
    const int rows = dgrad ? O : I, cols = dgrad ? I : O;
    return (long long)((rows + KC - 1) / KC) * T * round16(cols) * KC;
}

B200_API int b200_conv_tile_pack_weights(const float* w, float* out, int dgrad, int O, int I, int T, cudaStream_t st) {
    B200_REQUIRE(w && out && O > 0 && I > 0 && T > 0, "conv_tile_pack_weights: bad arguments");
    const int rows = dgrad ? O : I, cols = dgrad ? I : O;
    const int chunks = (rows + KC - 1) / KC, colsP = round16(cols);
    const long long total = (long long)chunks * T * colsP * KC;
    int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    pack_tile_weights_kernel<<<blocks, 256, 0, st>>>(w, out, dgrad, O, I, T, rows, cols, colsP, chunks);
    B200_CHECK_LAUNCH("conv_tile_pack_weights");
    return B200_OK;
}

// =====================================================================================
// launchers
// =====================================================================================
static int tile_supported(const b200_conv_desc* d, const char* who) {
    B200_REQUIRE(d != nullptr, "%s: null descriptor", who);
    B200_REQUIRE(d->stride == 1 && d->kh == 3 && d->kw == 3 && (d->kd == 1 || d->kd == 3) && d->ph == 1 && d->pw == 1 &&
                     d->pd == (d->kd == 3 ? 1 : 0),
                 "%s: tile kernels need a 3x3 / 3x3x3 stride-1 pad-1 convolution", who);
    B200_REQUIRE(d->n > 0 && d->id > 0 && d->ih > 0 && d->iw > 0 && d->cout > 0 && d->c0 > 0 && d->c1 >= 0, "%s: bad dims", who);
    return B200_OK;
}

B200_API int b200_conv_tile_supported(const b200_conv_desc* d, int for_wgrad) {
    if (!d || d->stride != 1 || d->kh != 3 || d->kw != 3 || d->ph != 1 || d->pw != 1) return 0;
    if (!((d->kd == 1 && d->pd == 0) || (d->kd == 3 && d->pd == 1))) return 0;
    if ((d->c0 & 3) || (d->c1 & 3)) return 0;                         // 16-byte pieces must not straddle sources
    if (for_wgrad) return (d->cout & 3) == 0;
    return 1;
}

template <int DIMS>
static void fill_tiles(TileP& p) {
    using S = TileShape<DIMS>;
    p.tiles_d = (p.D + S::TD - 1) / S::TD;
    p.tiles_h = (p.H + S::TH - 1) / S::TH;
    p.tiles_w = (p.W + S::TW - 1) / S::TW;
}

template <int DIMS, int BN>
static int launch_tile_fwd(const TileP& p, cudaStream_t st) {
    using SM = FwdSmem<DIMS, BN>;
    auto kern = conv_tile_fwd_kernel<DIMS, BN>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES);
        attr_done = true;
    }
    dim3 grid(p.N * p.tiles_d * p.tiles_h * p.tiles_w, (p.Cout + BN - 1) / BN);
    // a single (chunk, kd) step never touches the second pipeline stage: ask for half the shared memory so
    // more CTAs are resident and hide each other's load latency
    const bool one_step = DIMS == 2 && p.Cin <= KC;
    kern<<<grid, 256, one_step ? SM::BYTES / 2 : SM::BYTES, st>>>(p);
    B200_CHECK_LAUNCH("conv_tile_fwd");
    return B200_OK;
}

template <int DIMS>
static int dispatch_tile_fwd(TileP& p, cudaStream_t st) {
    fill_tiles<DIMS>(p);
    // wider column tiles amortise the halo reads; narrow ones keep more CTAs resident for small channel counts --
    // and the deep, spatially tiny layers (e.g. 256 channels at 6^3) need the narrow tiles to fill the 148 SMs at all
    int bn = p.Cout <= 16 ? 16 : (p.Cout <= 32 ? 32 : 64);
    const long long tiles = (long long)p.N * p.tiles_d * p.tiles_h * p.tiles_w;
    while (bn > 16 && tiles * ((p.Cout + bn - 1) / bn) < b200_num_sms()) bn >>= 1;
    if (bn == 16) return launch_tile_fwd<DIMS, 16>(p, st);
    if (bn == 32) return launch_tile_fwd<DIMS, 32>(p, st);
    return launch_tile_fwd<DIMS, 64>(p, st);
}

// forward: weights packed with b200_conv_tile_pack_weights(dgrad = 0)
B200_API int b200_conv_tile_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wt,
                                const float* bias, float* dst, int out_nchw, cudaStream_t st) {
    if (int rc = tile_supported(d, "conv_tile_fwd")) return rc;
    B200_REQUIRE(src0 && wt && dst && (d->c1 == 0 || src1), "conv_tile_fwd: null pointer");
    B200_REQUIRE((d->c0 & 3) == 0 && (d->c1 & 3) == 0, "conv_tile_fwd: channel counts must be multiples of 4");
    TileP p;
    memset(&p, 0, sizeof(p));
    p.src0 = src0; p.src1 = src1; p.C0 = d->c0; p.C1 = d->c1; p.Cin = d->c0 + d->c1;
    p.N = d->n; p.D = d->id; p.H = d->ih; p.W = d->iw;
    p.wt = wt; p.Cout = d->cout; p.CoutP = round16(d->cout); p.bias = bias;
    p.dst0 = dst; p.D0 = d->cout; p.D1 = 0; p.out_nchw = out_nchw;
    return d->kd == 3 ? dispatch_tile_fwd<3>(p, st) : dispatch_tile_fwd<2>(p, st);
}

// data gradient: weights packed with b200_conv_tile_pack_weights(dgrad = 1); dx may be split [dx0 | dx1]
B200_API int b200_conv_tile_dgrad(const b200_conv_desc* d, const float* dy, const float* wt_dgrad, float* dx0, float* dx1,
                                  int accumulate, cudaStream_t st) {
    if (int rc = tile_supported(d, "conv_tile_dgrad")) return rc;
    B200_REQUIRE(dy && wt_dgrad && dx0 && (d->c1 == 0 || dx1), "conv_tile_dgrad: null pointer");
    B200_REQUIRE((d->cout & 3) == 0 && (d->c0 & 1) == 0 && (d->c1 & 1) == 0, "conv_tile_dgrad: unsupported channel counts");
    TileP p;
    memset(&p, 0, sizeof(p));
    p.src0 = dy; p.C0 = d->cout; p.C1 = 0; p.Cin = d->cout;
    p.N = d->n; p.D = d->id; p.H = d->ih; p.W = d->iw;
    p.wt = wt_dgrad; p.Cout = d->c0 + d->c1; p.CoutP = round16(p.Cout);
    p.dst0 = dx0; p.dst1 = dx1; p.D0 = d->c0; p.D1 = d->c1; p.accumulate = accumulate;
    return d->kd == 3 ? dispatch_tile_fwd<3>(p, st) : dispatch_tile_fwd<2>(p, st);
}

struct TileWgPlan { int BN; int splits; int tiles_per_split; size_t ws_bytes; };

static TileWgPlan plan_tile_wgrad(const b200_conv_desc* d) {
    TileWgPlan pl;
    const int Cin = d->c0 + d->c1;
    const bool is3 = d->kd == 3;
    // 3D: two stages of (4 x 10 x 10 halo + 256 x BN dy) must fit in shared memory -> BN <= 32
    pl.BN = d->cout <= 16 ? 16 : ((d->cout <= 32 || is3) ? 32 : 64);
    const int ntiles = is3 ? d->n * ((d->id + 3) / 4) * ((d->ih + 7) / 8) * ((d->iw + 7) / 8)
                           : d->n * ((d->ih + 15) / 16) * ((d->iw + 15) / 16);
    const int taps = is3 ? 27 : 9;
    const int ctas = ((Cin + KC - 1) / KC) * ((d->cout + pl.BN - 1) / pl.BN) * (is3 ? 3 : 1);
    int splits = (2 * b200_num_sms() + ctas - 1) / ctas;
    if (splits > ntiles) splits = ntiles;
    if (splits < 1) splits = 1;
    pl.tiles_per_split = (ntiles + splits - 1) / splits;
    pl.splits = (ntiles + pl.tiles_per_split - 1) / pl.tiles_per_split;
    pl.ws_bytes = ((size_t)pl.splits * taps * Cin * d->cout + (size_t)pl.splits * d->cout) * sizeof(float);
    return pl;
}

B200_API long long b200_conv_tile_wgrad_workspace_bytes(const b200_conv_desc* d) {
    if (!d) return -1;
    return (long long)plan_tile_wgrad(d).ws_bytes;
}

// defined in conv.cu
int b200_wgrad_reduce_launch(const float* part, const float* part_colsum, int splits, int K, int NG, int A, int T,
                             float* dw, float* db, int accumulate, cudaStream_t st);

template <int DIMS, int BN>
static int launch_tile_wgrad(const TileP& p, int splits, cudaStream_t st) {
    using SM = WgSmem<DIMS, BN>;
    auto kern = conv_tile_wgrad_kernel<DIMS, BN>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES);
        attr_done = true;
    }
    dim3 grid((p.Cin + KC - 1) / KC, ((p.NG + BN - 1) / BN) * (DIMS == 3 ? 3 : 1), splits);
    kern<<<grid, 256, SM::BYTES, st>>>(p);
    B200_CHECK_LAUNCH("conv_tile_wgrad");
    return B200_OK;
}

B200_API int b200_conv_tile_wgrad(const b200_conv_desc* d, const float* src0, const float* src1, const float* dy,
                                  float* workspace, long long workspace_bytes, float* dw, float* db, int accumulate,
                                  cudaStream_t st) {
    if (int rc = tile_supported(d, "conv_tile_wgrad")) return rc;
    B200_REQUIRE(src0 && dy && workspace && dw && (d->c1 == 0 || src1), "conv_tile_wgrad: null pointer");
    B200_REQUIRE((d->c0 & 3) == 0 && (d->c1 & 3) == 0 && (d->cout & 3) == 0, "conv_tile_wgrad: channels must be multiples of 4");
    TileWgPlan pl = plan_tile_wgrad(d);
    if ((size_t)workspace_bytes < pl.ws_bytes) {
        b200_set_error("conv_tile_wgrad: workspace too small (%lld < %zu bytes)", workspace_bytes, pl.ws_bytes);
        return B200_ERR_WORKSPACE;
    }
    TileP p;
    memset(&p, 0, sizeof(p));
    p.src0 = src0; p.src1 = src1; p.C0 = d->c0; p.C1 = d->c1; p.Cin = d->c0 + d->c1;
    const bool is3 = d->kd == 3;
    const int taps = is3 ? 27 : 9;
    p.N = d->n; p.D = is3 ? d->id : 1; p.H = d->ih; p.W = d->iw;
    if (is3) fill_tiles<3>(p); else fill_tiles<2>(p);
    p.g = dy; p.NG = d->cout;
    p.part = workspace;
    p.part_colsum = db ? workspace + (size_t)pl.splits * taps * p.Cin * p.NG : nullptr;
    p.tiles_per_split = pl.tiles_per_split;
    int rc;
    if (is3) rc = pl.BN == 16 ? launch_tile_wgrad<3, 16>(p, pl.splits, st) : launch_tile_wgrad<3, 32>(p, pl.splits, st);
    else rc = pl.BN == 16 ? launch_tile_wgrad<2, 16>(p, pl.splits, st)
                          : (pl.BN == 32 ? launch_tile_wgrad<2, 32>(p, pl.splits, st) : launch_tile_wgrad<2, 64>(p, pl.splits, st));
    if (rc) return rc;
    return b200_wgrad_reduce_launch(p.part, p.part_colsum, pl.splits, taps * p.Cin, p.NG, p.Cin, taps, dw, db, accumulate, st);
}
