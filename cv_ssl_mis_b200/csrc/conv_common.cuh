// Implicit-GEMM convolution engine: shared parameter block, im2col gather and TF32 MMA helpers.
//
// All activations are channels-last ("NHWC" / "NDHWC", a 2D image is D=1): pixel-major rows of C
// contiguous fp32 channels. A convolution is the GEMM
//     out[m][n] = sum_k  im2col(m, k) * Wp[k][n],     m = (img, od, oh, ow),  k = (tap, ci)
// and its weight gradient is the transposed product reduced over pixels. The im2col operand is
// never materialised: it is gathered tile by tile into shared memory (zero-filled at the padding).
#pragma once
#include "common.cuh"
#include <cstring>

struct ConvP {
    // ---- im2col operand (input of fwd, dy of dgrad, x (or dy) of wgrad)
    const float* src0;
    const float* src1;      // second source of a virtual channel concat [src0 | src1]; may be null
    int C0, C1, Cin;        // Cin = C0 + C1
    int src_nchw;           // 1: src0 is [N][Cin][spatial] (scalar gather path only)
    int N, ID, IH, IW;      // input spatial dims (ID = 1 for 2D)
    int OD, OH, OW;         // GEMM-row ("output pixel") dims
    int KD, KH, KW, T;      // taps
    int stride, pd, ph, pw;
    int M, K;               // GEMM rows (N*OD*OH*OW), reduction length (T*Cin)
    FastDiv fd_cin, fd_khw, fd_kw, fd_ow, fd_oh, fd_od;
    // ---- weight operand
    const float* wp;        // packed [K][ldn]
    int ldn;                // row length of wp (Ngemm rounded up to 4, zero padded)
    int Ngemm;              // GEMM columns
    const float* bias;      // [Cout] or null
    // ---- epilogue
    float* dst0;
    float* dst1;            // second destination of a channel split [dst0 | dst1]; may be null
    int D0, D1;             // channels of each destination (D0 + D1 == Ngemm for EPI_NHWC)
    int epi;                // EPI_*
    int accumulate;         // 1: dst += result
    int Cout;               // real output channels (D2S: Ngemm = T2 * Cout)
    int d2s_dims;           // 2 or 3 (depth-to-space factor 2 in each of the last d2s_dims dims)
    FastDiv fd_cout;
};

enum { EPI_NHWC = 0, EPI_NCHW = 1, EPI_D2S = 2 };

struct PixCoord { int n, d, h, w; };   // (image, od*stride-pd, oh*stride-ph, ow*stride-pw)

__device__ __forceinline__ PixCoord conv_pix_decode(const ConvP& p, uint32_t m) {
    uint32_t t, ow, oh, od, n;
    p.fd_ow.divmod(m, t, ow);
    p.fd_oh.divmod(t, t, oh);
    p.fd_od.divmod(t, n, od);
    PixCoord c;
    c.n = (int)n;
    c.d = (int)od * p.stride - p.pd;
    c.h = (int)oh * p.stride - p.ph;
    c.w = (int)ow * p.stride - p.pw;
    return c;
}

struct TapCoord { int kd, kh, kw, ci; };

__device__ __forceinline__ TapCoord conv_k_decode(const ConvP& p, uint32_t k) {
    uint32_t tap, ci, kd, r, kh, kw;
    p.fd_cin.divmod(k, tap, ci);
    p.fd_khw.divmod(tap, kd, r);
    p.fd_kw.divmod(r, kh, kw);
    TapCoord t;
    t.kd = (int)kd; t.kh = (int)kh; t.kw = (int)kw; t.ci = (int)ci;
    return t;
}

// one float of the im2col operand (any channel count / layout)
__device__ __forceinline__ float conv_gather1(const ConvP& p, const PixCoord& pc, bool pvalid, uint32_t k) {
    if (!pvalid || k >= (uint32_t)p.K) return 0.f;
    TapCoord t = conv_k_decode(p, k);
    int id = pc.d + t.kd, ih = pc.h + t.kh, iw = pc.w + t.kw;
    if ((unsigned)id >= (unsigned)p.ID || (unsigned)ih >= (unsigned)p.IH || (unsigned)iw >= (unsigned)p.IW) return 0.f;
    size_t sp = ((size_t)id * p.IH + ih) * p.IW + iw;
    if (p.src_nchw) {
        size_t S = (size_t)p.ID * p.IH * p.IW;
        return __ldg(p.src0 + ((size_t)pc.n * p.Cin + t.ci) * S + sp);
    }
    size_t pix = (size_t)pc.n * p.ID * p.IH * p.IW + sp;
    if (t.ci < p.C0) return __ldg(p.src0 + pix * p.C0 + t.ci);
    return __ldg(p.src1 + pix * p.C1 + (t.ci - p.C0));
}

// four consecutive k (same tap; requires C0 % 4 == 0 and C1 % 4 == 0, NHWC)
__device__ __forceinline__ float4 conv_gather4(const ConvP& p, const PixCoord& pc, bool pvalid, const TapCoord& t, bool kvalid) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!pvalid || !kvalid) return z;
    int id = pc.d + t.kd, ih = pc.h + t.kh, iw = pc.w + t.kw;
    if ((unsigned)id >= (unsigned)p.ID || (unsigned)ih >= (unsigned)p.IH || (unsigned)iw >= (unsigned)p.IW) return z;
    size_t pix = (((size_t)pc.n * p.ID + id) * p.IH + ih) * p.IW + iw;
    if (t.ci < p.C0) return ldg4(p.src0 + pix * p.C0 + t.ci);
    return ldg4(p.src1 + pix * p.C1 + (t.ci - p.C0));
}

// ---------------------------------------------------------------- TF32 tensor-core helpers
__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
// x ~= hi + lo with both parts exactly representable in TF32 (3xTF32 "exact" mode)
__device__ __forceinline__ void f2tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
// D(16x8) += A(16x8, row) * B(8x8, col); fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool X3>
__device__ __forceinline__ void mma_block(float (&d)[4], const float (&af)[4], const float (&bf)[2]) {
    if (X3) {
        uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) f2tf32_split(af[i], ah[i], al[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) f2tf32_split(bf[i], bh[i], bl[i]);
        // The tensor core adds into its accumulator with truncation; over thousands of k-steps that drifts
        // (~5e-5 rel at K=2304, measured).  Keep the long accumulation in round-to-nearest FADDs instead:
        // the three partial products of this k8-step go into a fresh zero tile, then one FADD per element.
        float tsum[4] = {0.f, 0.f, 0.f, 0.f};
        mma_tf32(tsum, al, bh);
        mma_tf32(tsum, ah, bl);
        mma_tf32(tsum, ah, bh);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] += tsum[i];
    } else {
        uint32_t a[4], b[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = f2tf32(af[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) b[i] = f2tf32(bf[i]);
        mma_tf32(d, a, b);
    }
}
