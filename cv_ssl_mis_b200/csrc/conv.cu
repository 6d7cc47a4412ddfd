// Convolution forward / data-gradient / weight-gradient as implicit GEMMs on the tensor cores
// (TF32 inputs, fp32 accumulate; `exact` = 3xTF32 split, fp32-equivalent to ~1e-6 rel).
//
// Replaces the cuDNN calls behind nn.Conv2d / nn.Conv3d / nn.ConvTranspose3d of the reference
// (code/networks/unet.py:37,41,73,138 and code/networks/vnet.py:16,73,100,175).
#include "conv_common.cuh"
#include "../../include/b200ssl.h"

// =====================================================================================
// forward-style kernel: out[M][Ngemm] = im2col[M][K] * Wp[K][Ngemm] (+bias), several epilogues
// =====================================================================================
template <int WM, int WN, int MI, int NI, bool X3, bool VEC>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvP p) {
    constexpr int BM = WM * MI * 16, BN = WN * NI * 8, BK = 16;
    constexpr int LDA = BK + 4, LDB = BN + 8;
    constexpr int AJ = BM / 64;                       // pixel rows per thread
    constexpr int BSLOTS = BK * BN / 4;               // float4 slots of the weight tile
    constexpr int BJ = (BSLOTS + 255) / 256;
    static_assert(WM * WN == 8, "8 warps");
    static_assert(BM % 64 == 0, "BM multiple of 64");

    __shared__ __align__(16) float As[2][BM * LDA];
    __shared__ __align__(16) float Bs[2][BK * LDB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // this thread's im2col rows
    const int arow = tid >> 2, akq = tid & 3;
    PixCoord pc[AJ];
    bool pv[AJ];
#pragma unroll
    for (int j = 0; j < AJ; ++j) {
        int m = m0 + arow + 64 * j;
        pv[j] = m < p.M;
        pc[j] = conv_pix_decode(p, pv[j] ? m : 0);
    }

    float acc[MI][NI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    const int nchunks = (p.K + BK - 1) / BK;
    float4 ra[AJ], rb[BJ];

    auto gload = [&](int chunk) {
        const uint32_t k = chunk * BK + akq * 4;
        if (VEC) {
            bool kvalid = k < (uint32_t)p.K;
            TapCoord tc = conv_k_decode(p, kvalid ? k : 0);
#pragma unroll
            for (int j = 0; j < AJ; ++j) ra[j] = conv_gather4(p, pc[j], pv[j], tc, kvalid);
        } else {
#pragma unroll
            for (int j = 0; j < AJ; ++j) {
                ra[j].x = conv_gather1(p, pc[j], pv[j], k + 0);
                ra[j].y = conv_gather1(p, pc[j], pv[j], k + 1);
                ra[j].z = conv_gather1(p, pc[j], pv[j], k + 2);
                ra[j].w = conv_gather1(p, pc[j], pv[j], k + 3);
            }
        }
#pragma unroll
        for (int j = 0; j < BJ; ++j) {
            int s = tid + 256 * j;
            rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < BSLOTS) {
                int kr = s / (BN / 4), nq = s % (BN / 4);
                int kk = chunk * BK + kr, n = n0 + nq * 4;
                if (kk < p.K && n < p.ldn) rb[j] = ldg4(p.wp + (size_t)kk * p.ldn + n);
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < AJ; ++j)
            *reinterpret_cast<float4*>(&As[buf][(arow + 64 * j) * LDA + akq * 4]) = ra[j];
#pragma unroll
        for (int j = 0; j < BJ; ++j) {
            int s = tid + 256 * j;
            if (s < BSLOTS) {
                int kr = s / (BN / 4), nq = s % (BN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][kr * LDB + nq * 4]) = rb[j];
            }
        }
    };

    gload(0);
    sstore(0);
    __syncthreads();

    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < nchunks) gload(chunk + 1);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            float af[MI][4], bf[NI][2];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const float* a = &As[buf][(wm * MI * 16 + i * 16 + g) * LDA + ks * 8 + t];
                af[i][0] = a[0];
                af[i][1] = a[8 * LDA];
                af[i][2] = a[4];
                af[i][3] = a[8 * LDA + 4];
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const float* b = &Bs[buf][(ks * 8 + t) * LDB + wn * NI * 8 + j * 8 + g];
                bf[j][0] = b[0];
                bf[j][1] = b[4 * LDB];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) mma_block<X3>(acc[i][j], af[i], bf[j]);
        }
        if (chunk + 1 < nchunks) sstore(buf ^ 1);
        __syncthreads();
    }

    // ---------------- epilogue
#pragma unroll
    for (int i = 0; i < MI; ++i) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + wm * MI * 16 + i * 16 + g + 8 * h;
            if (m >= p.M) continue;
            PixCoord oc;
            size_t S = 0, sp = 0;
            if (p.epi != EPI_NHWC) {
                // decode with stride 1 / pad 0 semantics: plain (n, od, oh, ow)
                uint32_t tq, ow, oh, od, n;
                p.fd_ow.divmod((uint32_t)m, tq, ow);
                p.fd_oh.divmod(tq, tq, oh);
                p.fd_od.divmod(tq, n, od);
                oc.n = n; oc.d = od; oc.h = oh; oc.w = ow;
                S = (size_t)p.OD * p.OH * p.OW;
                sp = ((size_t)od * p.OH + oh) * p.OW + ow;
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = n0 + wn * NI * 8 + j * 8 + 2 * t + e;
                    if (n >= p.Ngemm) continue;
                    float v = acc[i][j][2 * h + e];
                    float* dst;
                    if (p.epi == EPI_NHWC) {
                        if (p.bias) v += __ldg(p.bias + n);
                        dst = (n < p.D0) ? p.dst0 + (size_t)m * p.D0 + n : p.dst1 + (size_t)m * p.D1 + (n - p.D0);
                    } else if (p.epi == EPI_NCHW) {
                        if (p.bias) v += __ldg(p.bias + n);
                        dst = p.dst0 + ((size_t)oc.n * p.Ngemm + n) * S + sp;
                    } else {   // EPI_D2S: n = tap * Cout + co, scatter to the 2x upsampled grid
                        uint32_t tap, co;
                        p.fd_cout.divmod((uint32_t)n, tap, co);
                        if (p.bias) v += __ldg(p.bias + co);
                        int kw = tap & 1, kh = (tap >> 1) & 1, kd = (p.d2s_dims == 3) ? (tap >> 2) : 0;
                        int OD2 = (p.d2s_dims == 3) ? 2 * p.OD : p.OD;
                        size_t opix = (((size_t)oc.n * OD2 + ((p.d2s_dims == 3) ? 2 * oc.d + kd : oc.d)) * (2 * p.OH) +
                                       2 * oc.h + kh) * (2 * p.OW) + 2 * oc.w + kw;
                        dst = p.dst0 + opix * p.Cout + co;
                    }
                    if (p.accumulate) v += *dst;
                    *dst = v;
                }
            }
        }
    }
}

template <int WM, int WN, int MI, int NI>
static void launch_cfg(const ConvP& p, bool exact, bool vec, cudaStream_t st) {
    constexpr int BM = WM * MI * 16, BN = WN * NI * 8;
    dim3 grid((p.M + BM - 1) / BM, (p.Ngemm + BN - 1) / BN);
    if (exact) {
        if (vec) conv_igemm_kernel<WM, WN, MI, NI, true, true><<<grid, 256, 0, st>>>(p);
        else conv_igemm_kernel<WM, WN, MI, NI, true, false><<<grid, 256, 0, st>>>(p);
    } else {
        if (vec) conv_igemm_kernel<WM, WN, MI, NI, false, true><<<grid, 256, 0, st>>>(p);
        else conv_igemm_kernel<WM, WN, MI, NI, false, false><<<grid, 256, 0, st>>>(p);
    }
}

static int launch_igemm(ConvP& p, bool exact, cudaStream_t st, const char* who) {
    p.fd_cin.init(p.Cin);
    p.fd_khw.init(p.KH * p.KW);
    p.fd_kw.init(p.KW);
    p.fd_ow.init(p.OW);
    p.fd_oh.init(p.OH);
    p.fd_od.init(p.OD);
    p.fd_cout.init(p.Cout > 0 ? p.Cout : 1);
    if (p.M <= 0 || p.Ngemm <= 0) return B200_OK;
    bool vec = !p.src_nchw && (p.C0 % 4 == 0) && (p.C1 % 4 == 0);
    const int sms = b200_num_sms();
    const int mt = (p.M + 127) / 128;
    if (p.Ngemm <= 16) launch_cfg<8, 1, 1, 2>(p, exact, vec, st);
    else if (p.Ngemm <= 32) launch_cfg<8, 1, 1, 4>(p, exact, vec, st);
    else if (p.Ngemm <= 64 || mt * ((p.Ngemm + 127) / 128) < 2 * sms) launch_cfg<4, 2, 2, 4>(p, exact, vec, st);
    else launch_cfg<2, 4, 4, 4>(p, exact, vec, st);
    B200_CHECK_LAUNCH(who);
    return B200_OK;
}

// =====================================================================================
// weight-gradient kernel: part[z][k][n] = sum_{m in split z} im2col[m][k] * G[m][n]
// =====================================================================================
struct WgradP {
    ConvP c;                // im2col operand geometry (src*, dims, K, M)
    const float* g;         // [M][NG]
    int NG;
    float* part;            // [splits][K][NG]
    float* part_colsum;     // [splits][NG] column sums of g (bias gradient) or null
    int rows_per_split;     // pixels per split (multiple of 64)
};

template <int WM, int WN, int WK, int MI, int NI>
struct WgradCfg {
    static constexpr int BMr = WM * MI * 16, BN = WN * NI * 8, BP = 64;
    static constexpr int LDP = BMr + 8, LDG = BN + 8;
    static constexpr int KQ = BMr / 4;                 // float4 columns of the im2col tile
    static constexpr int PIPE_FLOATS = 2 * BP * LDP + 2 * BP * LDG;
    static constexpr int RED_FLOATS = WK * BMr * BN;
    static constexpr int SMEM_FLOATS = (PIPE_FLOATS > RED_FLOATS ? PIPE_FLOATS : RED_FLOATS) + 4 * KQ;
    static constexpr size_t SMEM_BYTES = SMEM_FLOATS * sizeof(float);
};

template <int WM, int WN, int WK, int MI, int NI, bool X3, bool VEC>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradP q) {
    using Cfg = WgradCfg<WM, WN, WK, MI, NI>;
    constexpr int BMr = Cfg::BMr, BN = Cfg::BN, BP = Cfg::BP, LDP = Cfg::LDP, LDG = Cfg::LDG, KQ = Cfg::KQ;
    constexpr int KI = (KQ + 3) / 4;                   // im2col float4 columns per thread
    constexpr int GSLOTS = BP * BN / 4, GJ = (GSLOTS + 255) / 256;
    constexpr int KSTEPS = BP / WK / 8;
    static_assert(WM * WN * WK == 8, "8 warps");
    static_assert(BP % (WK * 8) == 0, "pixel slice");

    extern __shared__ __align__(16) float smem[];
    float* Ps = smem;                                   // [2][BP][LDP]
    float* Gs = smem + 2 * BP * LDP;                    // [2][BP][LDG]
    int* ktab = reinterpret_cast<int*>(smem + (Cfg::SMEM_FLOATS - 4 * KQ));   // [KQ][4] kd,kh,kw,ci (ci<0: invalid)

    const ConvP& p = q.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = (warp / WM) % WN, wk = warp / (WM * WN);
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * BMr, n0 = blockIdx.y * BN;
    const int mbeg = blockIdx.z * q.rows_per_split;
    const int mend = min(p.M, mbeg + q.rows_per_split);

    if (VEC && tid < KQ) {
        uint32_t k = r0 + tid * 4;
        int4 e;
        if (k < (uint32_t)p.K) {
            TapCoord tc = conv_k_decode(p, k);
            e = make_int4(tc.kd, tc.kh, tc.kw, tc.ci);
        } else e = make_int4(0, 0, 0, -1);
        reinterpret_cast<int4*>(ktab)[tid] = e;
    }
    __syncthreads();

    float acc[MI][NI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    const int prow = tid >> 2, kql = tid & 3;
    float4 rp[KI], rg[GJ];
    float colsum = 0.f;                                  // thread (c = tid % BN, slice = tid / BN)
    const bool do_colsum = q.part_colsum != nullptr && blockIdx.x == 0;

    auto gload = [&](int mbase) {
        const int m = mbase + prow;
        const bool pvalid = m < mend;
        PixCoord pc = conv_pix_decode(p, pvalid ? m : 0);
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            const int kq = kql + 4 * i;
            rp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kq < KQ) {
                if (VEC) {
                    int4 e = reinterpret_cast<const int4*>(ktab)[kq];
                    TapCoord tc; tc.kd = e.x; tc.kh = e.y; tc.kw = e.z; tc.ci = e.w;
                    rp[i] = conv_gather4(p, pc, pvalid, tc, e.w >= 0);
                } else {
                    uint32_t k = r0 + kq * 4;
                    rp[i].x = conv_gather1(p, pc, pvalid, k + 0);
                    rp[i].y = conv_gather1(p, pc, pvalid, k + 1);
                    rp[i].z = conv_gather1(p, pc, pvalid, k + 2);
                    rp[i].w = conv_gather1(p, pc, pvalid, k + 3);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < GJ; ++j) {
            const int s = tid + 256 * j;
            rg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < GSLOTS) {
                const int pr = s / (BN / 4), nq = s % (BN / 4);
                const int mm = mbase + pr, n = n0 + nq * 4;
                if (mm < mend) {
                    const float* gp = q.g + (size_t)mm * q.NG + n;
                    if ((q.NG & 3) == 0) {
                        if (n < q.NG) rg[j] = ldg4(gp);
                    } else {
                        if (n + 0 < q.NG) rg[j].x = __ldg(gp + 0);
                        if (n + 1 < q.NG) rg[j].y = __ldg(gp + 1);
                        if (n + 2 < q.NG) rg[j].z = __ldg(gp + 2);
                        if (n + 3 < q.NG) rg[j].w = __ldg(gp + 3);
                    }
                }
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            const int kq = kql + 4 * i;
            if (kq < KQ) *reinterpret_cast<float4*>(&Ps[(buf * BP + prow) * LDP + kq * 4]) = rp[i];
        }
#pragma unroll
        for (int j = 0; j < GJ; ++j) {
            const int s = tid + 256 * j;
            if (s < GSLOTS) {
                const int pr = s / (BN / 4), nq = s % (BN / 4);
                *reinterpret_cast<float4*>(&Gs[(buf * BP + pr) * LDG + nq * 4]) = rg[j];
            }
        }
    };

    const int nstages = (mend > mbeg) ? (mend - mbeg + BP - 1) / BP : 0;
    if (nstages > 0) {
        gload(mbeg);
        sstore(0);
    }
    __syncthreads();
    for (int s = 0; s < nstages; ++s) {
        const int buf = s & 1;
        if (s + 1 < nstages) gload(mbeg + (s + 1) * BP);
        const float* P = Ps + buf * BP * LDP;
        const float* G = Gs + buf * BP * LDG;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const int px = wk * (BP / WK) + ks * 8;
            float af[MI][4], bf[NI][2];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const float* a = P + (px + t) * LDP + wm * MI * 16 + i * 16 + g;
                af[i][0] = a[0];
                af[i][1] = a[8];
                af[i][2] = a[4 * LDP];
                af[i][3] = a[4 * LDP + 8];
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const float* b = G + (px + t) * LDG + wn * NI * 8 + j * 8 + g;
                bf[j][0] = b[0];
                bf[j][1] = b[4 * LDG];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) mma_block<X3>(acc[i][j], af[i], bf[j]);
        }
        if (do_colsum) {
            const int c = tid % BN, sl = tid / BN;
            for (int pr = sl; pr < BP; pr += 256 / BN) colsum += G[pr * LDG + c];
        }
        if (s + 1 < nstages) sstore(buf ^ 1);
        __syncthreads();
    }

    // ---------------- reduce the WK pixel slices through shared memory, write the split partial
    float* red = smem;                                   // [WK][BMr][BN]
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = wm * MI * 16 + i * 16 + g + 8 * (e >> 1);
                const int c = wn * NI * 8 + j * 8 + 2 * t + (e & 1);
                red[(wk * BMr + r) * BN + c] = acc[i][j][e];
            }
    __syncthreads();
    float* out = q.part + (size_t)blockIdx.z * p.K * q.NG;
    for (int idx = tid; idx < BMr * BN; idx += 256) {
        const int r = idx / BN, c = idx % BN;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < WK; ++w) v += red[(w * BMr + r) * BN + c];
        const int k = r0 + r, n = n0 + c;
        if (k < p.K && n < q.NG) out[(size_t)k * q.NG + n] = v;
    }
    if (do_colsum) {
        __syncthreads();
        red[tid] = colsum;
        __syncthreads();
        if (tid < BN) {
            float v = 0.f;
            for (int sl = 0; sl < 256 / BN; ++sl) v += red[sl * BN + tid];
            if (n0 + tid < q.NG) q.part_colsum[(size_t)blockIdx.z * q.NG + n0 + tid] = v;
        }
    }
}

// sum the split partials (fixed order: deterministic) and scatter to the framework's weight layout:
// packed row k = (tap, a), column b  ->  grad[(b * A + a) * T + tap].
// Block = 32 elements x 8 split lanes: the dependent-load depth is splits/8 instead of splits (this kernel was 20 us of
// pure latency per layer with one thread per element).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_colsum,
                                                           int splits, int K, int NG, int A, int T, float* __restrict__ dw,
                                                           float* __restrict__ db, int accumulate) {
    __shared__ float sh[8][33];
    const int total = K * NG;
    const int e = threadIdx.x & 31, zl = threadIdx.x >> 5;
    for (int base = blockIdx.x * 32; base < total + NG; base += gridDim.x * 32) {
        const int idx = base + e;
        float v = 0.f;
        if (idx < total) {
            for (int z = zl; z < splits; z += 8) v += part[(size_t)z * total + idx];
        } else if (idx < total + NG && part_colsum != nullptr) {
            for (int z = zl; z < splits; z += 8) v += part_colsum[(size_t)z * NG + (idx - total)];
        }
        sh[zl][e] = v;
        __syncthreads();
        if (zl == 0) {
            float r = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) r += sh[w][e];
            if (idx < total) {
                const int k = idx / NG, bcol = idx % NG;
                const int tap = k / A, a = k % A;
                float* d = dw + ((size_t)bcol * A + a) * T + tap;
                *d = accumulate ? *d + r : r;
            } else if (idx < total + NG && db != nullptr && part_colsum != nullptr) {
                const int bcol = idx - total;
                db[bcol] = accumulate ? db[bcol] + r : r;
            }
        }
        __syncthreads();
    }
}

// shared with conv_tile.cu
int b200_wgrad_reduce_launch(const float* part, const float* part_colsum, int splits, int K, int NG, int A, int T,
                             float* dw, float* db, int accumulate, cudaStream_t st) {
    const int total = K * NG + NG;
    int blocks = (total + 31) / 32;
    if (blocks > 8 * b200_num_sms()) blocks = 8 * b200_num_sms();
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(part, part_colsum, splits, K, NG, A, T, dw, db, accumulate);
    B200_CHECK_LAUNCH("wgrad_reduce");
    return B200_OK;
}

template <int WM, int WN, int WK, int MI, int NI>
static int launch_wgrad_cfg(WgradP& q, int splits, bool exact, bool vec, cudaStream_t st) {
    using Cfg = WgradCfg<WM, WN, WK, MI, NI>;
    dim3 grid((q.c.K + Cfg::BMr - 1) / Cfg::BMr, (q.NG + Cfg::BN - 1) / Cfg::BN, splits);
#define B200_WG_LAUNCH(X3, VEC)                                                                           \
    do {                                                                                                  \
        auto kern = conv_wgrad_kernel<WM, WN, WK, MI, NI, X3, VEC>;                                       \
        static bool attr_done = false;                                                                    \
        if (!attr_done) {                                                                                 \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES); \
            attr_done = true;                                                                             \
        }                                                                                                 \
        kern<<<grid, 256, Cfg::SMEM_BYTES, st>>>(q);                                                      \
    } while (0)
    if (exact) { if (vec) B200_WG_LAUNCH(true, true); else B200_WG_LAUNCH(true, false); }
    else       { if (vec) B200_WG_LAUNCH(false, true); else B200_WG_LAUNCH(false, false); }
#undef B200_WG_LAUNCH
    B200_CHECK_LAUNCH("conv_wgrad");
    return B200_OK;
}

struct WgradPlan { int cfg; int BMr, BN; int splits; int rows_per_split; size_t ws_bytes; };

static WgradPlan plan_wgrad(int M, int K, int NG) {
    WgradPlan pl;
    if (NG <= 16 || K <= 48) { pl.cfg = 0; pl.BMr = 48; pl.BN = 16; }
    else if (NG <= 32 || K <= 64) { pl.cfg = 1; pl.BMr = 64; pl.BN = 32; }
    else { pl.cfg = 2; pl.BMr = 128; pl.BN = 64; }
    const int tiles = ((K + pl.BMr - 1) / pl.BMr) * ((NG + pl.BN - 1) / pl.BN);
    const int sms = b200_num_sms();
    int want = (4 * sms + tiles - 1) / tiles;                 // ~4 CTAs per SM in flight
    int max_splits = (M + 255) / 256;                         // at least 4 stages per split
    int splits = want < 1 ? 1 : want;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rps = (M + splits - 1) / splits;
    rps = (rps + 63) / 64 * 64;
    splits = (M + rps - 1) / rps;
    pl.splits = splits;
    pl.rows_per_split = rps;
    pl.ws_bytes = ((size_t)splits * K * NG + (size_t)splits * NG) * sizeof(float);
    return pl;
}

static int run_wgrad(ConvP& c, const float* g, int NG, float* ws, size_t ws_bytes, float* dw, float* db,
                     int A, int T, int accumulate, bool exact, cudaStream_t st) {
    c.fd_cin.init(c.Cin);
    c.fd_khw.init(c.KH * c.KW);
    c.fd_kw.init(c.KW);
    c.fd_ow.init(c.OW);
    c.fd_oh.init(c.OH);
    c.fd_od.init(c.OD);
    c.fd_cout.init(1);
    WgradPlan pl = plan_wgrad(c.M, c.K, NG);
    if (ws_bytes < pl.ws_bytes) {
        b200_set_error("conv_wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, pl.ws_bytes);
        return B200_ERR_WORKSPACE;
    }
    WgradP q;
    q.c = c;
    q.g = g;
    q.NG = NG;
    q.part = ws;
    q.part_colsum = db ? ws + (size_t)pl.splits * c.K * NG : nullptr;
    q.rows_per_split = pl.rows_per_split;
    bool vec = !c.src_nchw && (c.C0 % 4 == 0) && (c.C1 % 4 == 0);
    int rc;
    if (pl.cfg == 0) rc = launch_wgrad_cfg<1, 1, 8, 3, 2>(q, pl.splits, exact, vec, st);
    else if (pl.cfg == 1) rc = launch_wgrad_cfg<2, 1, 4, 2, 4>(q, pl.splits, exact, vec, st);
    else rc = launch_wgrad_cfg<2, 2, 2, 4, 4>(q, pl.splits, exact, vec, st);
    if (rc) return rc;
    return b200_wgrad_reduce_launch(q.part, q.part_colsum, pl.splits, c.K, NG, A, T, dw, db, accumulate, st);
}

// =====================================================================================
// weight packing
// =====================================================================================
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int mode, int O, int I,
                                    int T, int rows, int cols, int ldn) {
    const int total = rows * ldn;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int r = idx / ldn, c = idx % ldn;
        float v = 0.f;
        if (c < cols) {
            switch (mode) {
                case B200_PACK_CONV_FWD: {          // rows (tap, i), cols o ; w[O][I][T]
                    int tap = r / I, i = r % I;
                    v = w[((size_t)c * I + i) * T + tap];
                } break;
                case B200_PACK_CONV_DGRAD: {        // rows (tap', o), cols i ; flipped taps
                    int tap = r / O, o = r % O;
                    v = w[((size_t)o * I + c) * T + (T - 1 - tap)];
                } break;
                case B200_PACK_CONV_DGRAD_D2S: {    // rows o, cols (tap, i)
                    int tap = c / I, i = c % I;
                    v = w[((size_t)r * I + i) * T + tap];
                } break;
                case B200_PACK_DECONV_FWD: {        // rows i, cols (tap, o) ; w[I][O][T]
                    int tap = c / O, o = c % O;
                    v = w[((size_t)r * O + o) * T + tap];
                } break;
                case B200_PACK_DECONV_DGRAD: {      // rows (tap, o), cols i ; w[I][O][T]
                    int tap = r / O, o = r % O;
                    v = w[((size_t)c * O + o) * T + tap];
                } break;
            }
        }
        out[idx] = v;
    }
}

static void pack_dims(int mode, int O, int I, int T, int& rows, int& cols) {
    switch (mode) {
        case B200_PACK_CONV_FWD: rows = T * I; cols = O; break;
        case B200_PACK_CONV_DGRAD: rows = T * O; cols = I; break;
        case B200_PACK_CONV_DGRAD_D2S: rows = O; cols = T * I; break;
        case B200_PACK_DECONV_FWD: rows = I; cols = T * O; break;
        default: rows = T * O; cols = I; break;
    }
}

B200_API long long b200_conv_packed_floats(int mode, int O, int I, int T) {
    int rows, cols;
    pack_dims(mode, O, I, T, rows, cols);
    return (long long)rows * ((cols + 3) / 4 * 4);
}

B200_API int b200_conv_pack_weights(const float* w, float* out, int mode, int O, int I, int T, cudaStream_t st) {
    B200_REQUIRE(w && out, "conv_pack_weights: null pointer");
    B200_REQUIRE(mode >= 0 && mode <= B200_PACK_DECONV_DGRAD, "conv_pack_weights: bad mode %d", mode);
    int rows, cols;
    pack_dims(mode, O, I, T, rows, cols);
    const int ldn = (cols + 3) / 4 * 4;
    const int total = rows * ldn;
    int blocks = (total + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    pack_weights_kernel<<<blocks, 256, 0, st>>>(w, out, mode, O, I, T, rows, cols, ldn);
    B200_CHECK_LAUNCH("conv_pack_weights");
    return B200_OK;
}

// =====================================================================================
// C-ABI entry points
// =====================================================================================
static int check_desc(const b200_conv_desc* d, const char* who) {
    B200_REQUIRE(d != nullptr, "%s: null descriptor", who);
    B200_REQUIRE(d->n > 0 && d->id > 0 && d->ih > 0 && d->iw > 0, "%s: bad input dims", who);
    B200_REQUIRE(d->c0 > 0 && d->c1 >= 0 && d->cout > 0, "%s: bad channel counts", who);
    B200_REQUIRE(d->kd > 0 && d->kh > 0 && d->kw > 0, "%s: bad kernel dims", who);
    B200_REQUIRE(d->stride == 1 || d->stride == 2, "%s: stride must be 1 or 2", who);
    return B200_OK;
}

static void out_dims(const b200_conv_desc* d, int& od, int& oh, int& ow) {
    od = (d->id + 2 * d->pd - d->kd) / d->stride + 1;
    oh = (d->ih + 2 * d->ph - d->kh) / d->stride + 1;
    ow = (d->iw + 2 * d->pw - d->kw) / d->stride + 1;
}

static ConvP base_params(const b200_conv_desc* d) {
    ConvP p;
    memset(&p, 0, sizeof(p));
    p.C0 = d->c0; p.C1 = d->c1; p.Cin = d->c0 + d->c1;
    p.N = d->n; p.ID = d->id; p.IH = d->ih; p.IW = d->iw;
    out_dims(d, p.OD, p.OH, p.OW);
    p.KD = d->kd; p.KH = d->kh; p.KW = d->kw; p.T = d->kd * d->kh * d->kw;
    p.stride = d->stride; p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
    p.M = d->n * p.OD * p.OH * p.OW;
    p.K = p.T * p.Cin;
    p.Cout = d->cout;
    p.d2s_dims = 2;
    return p;
}

B200_API int b200_conv_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wp,
                           const float* bias, float* dst, int out_nchw, int exact, cudaStream_t st) {
    if (int rc = check_desc(d, "conv_fwd")) return rc;
    B200_REQUIRE(src0 && wp && dst, "conv_fwd: null pointer");
    B200_REQUIRE(d->c1 == 0 || src1, "conv_fwd: c1 > 0 needs src1");
    ConvP p = base_params(d);
    p.src0 = src0; p.src1 = src1;
    p.wp = wp; p.Ngemm = d->cout; p.ldn = (d->cout + 3) / 4 * 4; p.bias = bias;
    p.dst0 = dst; p.D0 = d->cout; p.D1 = 0;
    p.epi = out_nchw ? EPI_NCHW : EPI_NHWC;
    return launch_igemm(p, exact != 0, st, "conv_fwd");
}

// Data gradient of a 1x1 (1x1x1) convolution with a handful of output channels (the segmentation heads: Swin-UNet 96 -> 4,
// VNet / UNETR 16 -> 2): dx[m][ci] = sum_co dy[m][co] W[co][ci] is a pure HBM stream (it writes Cin / Cout times what it
// reads), so plain FFMA with 16-byte stores instead of the implicit GEMM (316 -> ~70 us on the 224^2 x 16 Swin head).
template <int CO>
__global__ void __launch_bounds__(256) conv1x1_head_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ wp,
                                                                 float* __restrict__ dx, long long total4, int C4, int ldn,
                                                                 int accumulate, FastDiv fdC4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        uint32_t m, c4;
        fdC4.divmod((uint32_t)i, m, c4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < CO; ++k) {
            const float g = __ldg(dy + (size_t)m * CO + k);
            const float4 w = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * ldn) + c4);
            acc.x += g * w.x; acc.y += g * w.y; acc.z += g * w.z; acc.w += g * w.w;
        }
        float4* o = reinterpret_cast<float4*>(dx) + i;
        if (accumulate) { const float4 old = *o; acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w; }
        *o = acc;
    }
}

// stride-1 "same" convolution: dx = conv(dy, flipped weights); dx may be split over [dx0 | dx1]
B200_API int b200_conv_dgrad(const b200_conv_desc* d, const float* dy, const float* wp_dgrad, float* dx0, float* dx1,
                             int accumulate, int exact, cudaStream_t st) {
    if (int rc = check_desc(d, "conv_dgrad")) return rc;
    B200_REQUIRE(dy && wp_dgrad && dx0, "conv_dgrad: null pointer");
    B200_REQUIRE(d->stride == 1, "conv_dgrad: stride-1 only (use conv_k2s2_dgrad)");
    B200_REQUIRE(d->c1 == 0 || dx1, "conv_dgrad: c1 > 0 needs dx1");
    int od, oh, ow;
    out_dims(d, od, oh, ow);
    B200_REQUIRE(od == d->id && oh == d->ih && ow == d->iw, "conv_dgrad: only 'same' padding supported");
    if (!exact && d->kd * d->kh * d->kw == 1 && d->c1 == 0 && (d->c0 & 3) == 0 && (d->cout == 2 || d->cout == 4)) {
        const long long M = (long long)d->n * d->id * d->ih * d->iw, total4 = M * (d->c0 / 4);
        if (total4 < (1ll << 32)) {
            FastDiv fd;
            fd.init((uint32_t)(d->c0 / 4));
            long long blocks = (total4 + 255) / 256;
            const long long cap = (long long)b200_num_sms() * 16;
            const int grid = (int)(blocks < cap ? blocks : cap);
            if (d->cout == 2) conv1x1_head_dgrad_kernel<2><<<grid, 256, 0, st>>>(dy, wp_dgrad, dx0, total4, d->c0 / 4, d->c0, accumulate, fd);
            else conv1x1_head_dgrad_kernel<4><<<grid, 256, 0, st>>>(dy, wp_dgrad, dx0, total4, d->c0 / 4, d->c0, accumulate, fd);
            B200_CHECK_LAUNCH("conv_dgrad (1x1 head)");
            return B200_OK;
        }
    }
    ConvP p;
    memset(&p, 0, sizeof(p));
    p.src0 = dy; p.C0 = d->cout; p.C1 = 0; p.Cin = d->cout;
    p.N = d->n; p.ID = od; p.IH = oh; p.IW = ow;
    p.OD = d->id; p.OH = d->ih; p.OW = d->iw;
    p.KD = d->kd; p.KH = d->kh; p.KW = d->kw; p.T = d->kd * d->kh * d->kw;
    p.stride = 1; p.pd = d->kd - 1 - d->pd; p.ph = d->kh - 1 - d->ph; p.pw = d->kw - 1 - d->pw;
    p.M = d->n * p.OD * p.OH * p.OW;
    p.K = p.T * p.Cin;
    p.wp = wp_dgrad; p.Ngemm = d->c0 + d->c1; p.ldn = (p.Ngemm + 3) / 4 * 4; p.bias = nullptr;
    p.dst0 = dx0; p.dst1 = dx1; p.D0 = d->c0; p.D1 = d->c1;
    p.epi = EPI_NHWC; p.accumulate = accumulate; p.Cout = p.Ngemm; p.d2s_dims = 2;
    return launch_igemm(p, exact != 0, st, "conv_dgrad");
}

// kernel == stride == 2, pad 0 convolution: dx[2o + tap][ci] = sum_co dy[o][co] w[co][ci][tap]
B200_API int b200_conv_k2s2_dgrad(const b200_conv_desc* d, const float* dy, const float* wp_d2s, float* dx,
                                  int accumulate, int exact, cudaStream_t st) {
    if (int rc = check_desc(d, "conv_k2s2_dgrad")) return rc;
    B200_REQUIRE(dy && wp_d2s && dx, "conv_k2s2_dgrad: null pointer");
    B200_REQUIRE(d->stride == 2 && d->kh == 2 && d->kw == 2 && (d->kd == 2 || d->kd == 1) && d->pd == 0 && d->ph == 0 &&
                     d->pw == 0 && d->c1 == 0,
                 "conv_k2s2_dgrad: needs kernel 2 stride 2 pad 0, single source");
    B200_REQUIRE((d->kd == 1 || d->id % 2 == 0) && d->ih % 2 == 0 && d->iw % 2 == 0, "conv_k2s2_dgrad: odd input dims");
    ConvP p;
    memset(&p, 0, sizeof(p));
    int od, oh, ow;
    out_dims(d, od, oh, ow);
    p.src0 = dy; p.C0 = d->cout; p.Cin = d->cout;
    p.N = d->n; p.ID = od; p.IH = oh; p.IW = ow;
    p.OD = od; p.OH = oh; p.OW = ow;
    p.KD = p.KH = p.KW = 1; p.T = 1; p.stride = 1;
    p.M = d->n * od * oh * ow;
    p.K = d->cout;
    const int T2 = d->kd * 4;
    p.wp = wp_d2s; p.Ngemm = T2 * d->c0; p.ldn = (p.Ngemm + 3) / 4 * 4;
    p.dst0 = dx; p.D0 = p.Ngemm;
    p.epi = EPI_D2S; p.accumulate = accumulate; p.Cout = d->c0; p.d2s_dims = d->kd == 2 ? 3 : 2;
    return launch_igemm(p, exact != 0, st, "conv_k2s2_dgrad");
}

// ConvTranspose, kernel == stride == 2: y[2i + tap][co] = b[co] + sum_ci x[i][ci] w[ci][co][tap]
// (desc: n,id,ih,iw = INPUT (low-res) dims, c0 = Cin, cout = Cout, kd/kh/kw = 2 (kd = 1 for 2D))
B200_API int b200_deconv_k2s2_fwd(const b200_conv_desc* d, const float* x, const float* wp, const float* bias,
                                  float* y, int exact, cudaStream_t st) {
    if (int rc = check_desc(d, "deconv_k2s2_fwd")) return rc;
    B200_REQUIRE(x && wp && y, "deconv_k2s2_fwd: null pointer");
    B200_REQUIRE(d->kh == 2 && d->kw == 2 && (d->kd == 2 || d->kd == 1) && d->c1 == 0, "deconv_k2s2_fwd: kernel must be 2");
    ConvP p;
    memset(&p, 0, sizeof(p));
    p.src0 = x; p.C0 = d->c0; p.Cin = d->c0;
    p.N = d->n; p.ID = d->id; p.IH = d->ih; p.IW = d->iw;
    p.OD = d->id; p.OH = d->ih; p.OW = d->iw;
    p.KD = p.KH = p.KW = 1; p.T = 1; p.stride = 1;
    p.M = d->n * d->id * d->ih * d->iw;
    p.K = d->c0;
    const int T2 = d->kd * 4;
    p.wp = wp; p.Ngemm = T2 * d->cout; p.ldn = (p.Ngemm + 3) / 4 * 4; p.bias = bias;
    p.dst0 = y; p.D0 = p.Ngemm;
    p.epi = EPI_D2S; p.Cout = d->cout; p.d2s_dims = d->kd == 2 ? 3 : 2;
    return launch_igemm(p, exact != 0, st, "deconv_k2s2_fwd");
}

// dx[i][ci] = sum_{tap,co} dy[2i + tap][co] w[ci][co][tap]  (a kernel-2 stride-2 convolution over dy)
B200_API int b200_deconv_k2s2_dgrad(const b200_conv_desc* d, const float* dy, const float* wp_dgrad, float* dx,
                                    int accumulate, int exact, cudaStream_t st) {
    if (int rc = check_desc(d, "deconv_k2s2_dgrad")) return rc;
    B200_REQUIRE(dy && wp_dgrad && dx, "deconv_k2s2_dgrad: null pointer");
    B200_REQUIRE(d->kh == 2 && d->kw == 2 && (d->kd == 2 || d->kd == 1) && d->c1 == 0, "deconv_k2s2_dgrad: kernel must be 2");
    ConvP p;
    memset(&p, 0, sizeof(p));
    p.src0 = dy; p.C0 = d->cout; p.Cin = d->cout;
    p.N = d->n; p.ID = d->id * d->kd; p.IH = d->ih * 2; p.IW = d->iw * 2;
    p.OD = d->id; p.OH = d->ih; p.OW = d->iw;
    p.KD = d->kd; p.KH = 2; p.KW = 2; p.T = d->kd * 4; p.stride = 2;
    // stride applies to every dim; a 2D problem has ID == 1 so od*2 == 0 stays in range
    p.M = d->n * d->id * d->ih * d->iw;
    p.K = p.T * p.Cin;
    p.wp = wp_dgrad; p.Ngemm = d->c0; p.ldn = (d->c0 + 3) / 4 * 4;
    p.dst0 = dx; p.D0 = d->c0;
    p.epi = EPI_NHWC; p.accumulate = accumulate; p.Cout = d->c0; p.d2s_dims = 2;
    return launch_igemm(p, exact != 0, st, "deconv_k2s2_dgrad");
}

B200_API long long b200_conv_wgrad_workspace_bytes(const b200_conv_desc* d) {
    if (!d) return -1;
    int od, oh, ow;
    out_dims(d, od, oh, ow);
    const int M = d->n * od * oh * ow;
    const int K = d->kd * d->kh * d->kw * (d->c0 + d->c1);
    return (long long)plan_wgrad(M, K, d->cout).ws_bytes;
}

// dw[co][ci][tap] = sum_pixels dy[o][co] * x[o*stride + tap - pad][ci] ; db[co] = sum_pixels dy[o][co]
B200_API int b200_conv_wgrad(const b200_conv_desc* d, const float* src0, const float* src1, const float* dy,
                             float* workspace, long long workspace_bytes, float* dw, float* db, int accumulate,
                             int exact, cudaStream_t st) {
    if (int rc = check_desc(d, "conv_wgrad")) return rc;
    B200_REQUIRE(src0 && dy && workspace && dw, "conv_wgrad: null pointer");
    B200_REQUIRE(d->c1 == 0 || src1, "conv_wgrad: c1 > 0 needs src1");
    ConvP p = base_params(d);
    p.src0 = src0; p.src1 = src1;
    return run_wgrad(p, dy, d->cout, workspace, (size_t)workspace_bytes, dw, db, p.Cin, p.T, accumulate, exact != 0, st);
}

B200_API long long b200_deconv_k2s2_wgrad_workspace_bytes(const b200_conv_desc* d) {
    if (!d) return -1;
    const int M = d->n * d->id * d->ih * d->iw;
    const int K = d->kd * 4 * d->cout;
    return (long long)plan_wgrad(M, K, d->c0).ws_bytes;
}

// dw[ci][co][tap] = sum_i x[i][ci] * dy[2i + tap][co]   (db is a plain column sum of dy: b200_colsum)
B200_API int b200_deconv_k2s2_wgrad(const b200_conv_desc* d, const float* x, const float* dy, float* workspace,
                                    long long workspace_bytes, float* dw, int accumulate, int exact, cudaStream_t st) {
    if (int rc = check_desc(d, "deconv_k2s2_wgrad")) return rc;
    B200_REQUIRE(x && dy && workspace && dw, "deconv_k2s2_wgrad: null pointer");
    B200_REQUIRE(d->kh == 2 && d->kw == 2 && (d->kd == 2 || d->kd == 1) && d->c1 == 0, "deconv_k2s2_wgrad: kernel must be 2");
    ConvP p;
    memset(&p, 0, sizeof(p));
    p.src0 = dy; p.C0 = d->cout; p.Cin = d->cout;
    p.N = d->n; p.ID = d->id * d->kd; p.IH = d->ih * 2; p.IW = d->iw * 2;
    p.OD = d->id; p.OH = d->ih; p.OW = d->iw;
    p.KD = d->kd; p.KH = 2; p.KW = 2; p.T = d->kd * 4; p.stride = 2;
    p.M = d->n * d->id * d->ih * d->iw;
    p.K = p.T * p.Cin;
    // packed rows (tap, co), columns ci  ->  dw[(ci * Cout + co) * T + tap]
    return run_wgrad(p, x, d->c0, workspace, (size_t)workspace_bytes, dw, nullptr, d->cout, p.T, accumulate, exact != 0, st);
}
