// Token-space kernels of the Swin-UNet path (code/networks/swin_transformer_unet_skip_expand_decoder_sys.py):
// LayerNorm, GELU, shifted-window multi-head attention with relative-position bias (forward + backward, all index
// math of roll / window_partition / window_reverse folded into the gather), DropPath residual, the space<->depth
// rearrangements of PatchMerging / PatchEmbed.  Linear layers run on the convolution GEMM engine (1x1 convs).
//
// Tokens are rows of a [B*H*W][C] fp32 matrix in natural (un-rolled, un-partitioned) order everywhere; the reference's
// roll -> partition -> attention -> reverse -> roll back (:259-282) only permutes which rows attend to each other, so the
// attention kernel computes those row indices instead of moving data.
#include "common.cuh"
#include <cstring>
#include "../../include/b200ssl.h"

static inline int ew_grid(long long work) {
    long long blocks = (work + 255) / 256;
    long long cap = (long long)b200_num_sms() * 16;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// ================================================================================================ LayerNorm
// one warp per row; stats[row] = (mean, rstd) kept for backward.  nn.LayerNorm(C), eps 1e-5, biased variance.
#define LN_MAXC 1536

template <int LN_PER_LANE>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ stats, long long M, int C, float eps) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
        const float* xr = x + row * C;
        float v[LN_PER_LANE];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LN_PER_LANE; ++i) {
            const int c = lane + 32 * i;
            v[i] = c < C ? __ldg(xr + c) : 0.f;
            s += v[i];
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_PER_LANE; ++i) {
            const int c = lane + 32 * i;
            const float d = c < C ? v[i] - mean : 0.f;
            q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
        float* yr = y + row * C;
#pragma unroll
        for (int i = 0; i < LN_PER_LANE; ++i) {
            const int c = lane + 32 * i;
            if (c < C) yr[c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        }
        if (lane == 0 && stats) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
    }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma ; per-block partial sums of dgamma / dbeta
template <int LN_PER_LANE>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ dy,
                                                            float* __restrict__ dx, float* __restrict__ part, long long M,
                                                            int C, int accumulate) {
    extern __shared__ float sm[];                       // [8 warps][2][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float dg[LN_PER_LANE], db[LN_PER_LANE];
#pragma unroll
    for (int i = 0; i < LN_PER_LANE; ++i) { dg[i] = 0.f; db[i] = 0.f; }
    for (long long row = (long long)blockIdx.x * wpb + warp; row < M; row += (long long)gridDim.x * wpb) {
        const float mean = stats[2 * row], rstd = stats[2 * row + 1];
        const float* xr = x + row * C;
        const float* gr = dy + row * C;
        float xh[LN_PER_LANE], g[LN_PER_LANE];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_PER_LANE; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float d = __ldg(gr + c);
                xh[i] = (__ldg(xr + c) - mean) * rstd;
                g[i] = d * __ldg(gamma + c);
                dg[i] += d * xh[i];
                db[i] += d;
                s1 += g[i];
                s2 += g[i] * xh[i];
            } else { xh[i] = 0.f; g[i] = 0.f; }
        }
        s1 = warp_sum(s1) / (float)C;
        s2 = warp_sum(s2) / (float)C;
        float* dr = dx + row * C;
#pragma unroll
        for (int i = 0; i < LN_PER_LANE; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float v = rstd * (g[i] - s1 - xh[i] * s2);
                dr[c] = accumulate ? dr[c] + v : v;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LN_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        if (c < C) { sm[(warp * 2 + 0) * C + c] = dg[i]; sm[(warp * 2 + 1) * C + c] = db[i]; }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
        const int which = idx / C, c = idx % C;
        float s = 0.f;
        for (int w = 0; w < wpb; ++w) s += sm[(w * 2 + which) * C + c];
        part[(size_t)blockIdx.x * 2 * C + idx] = s;
    }
}

// one warp per (dgamma | dbeta) column: lanes stride over the per-block partials, fixed-order shuffle tree
__global__ void __launch_bounds__(256) colpart_finalize_kernel(const float* __restrict__ part, int nblk, int C2,
                                                              float* __restrict__ out0, float* __restrict__ out1, int C,
                                                              int accumulate) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= C2) return;
    double s = 0;
    for (int b = lane; b < nblk; b += 32) s += (double)part[(size_t)b * C2 + idx];
    s = warp_sum_d(s);
    if (lane == 0) {
        float* o = idx < C ? out0 + idx : out1 + (idx - C);
        *o = accumulate ? *o + (float)s : (float)s;
    }
}

static inline int ln_grid(long long M) {
    long long want = (M + 7) / 8;
    const long long cap = (long long)b200_num_sms() * 8;      // one row in flight per warp: the sweep needs many resident warps to cover the load latency
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

B200_API long long b200_layernorm_workspace_bytes(long long M, int C) {
    return (long long)ln_grid(M) * 2 * C * sizeof(float);
}

B200_API int b200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, long long M,
                                int C, float eps, cudaStream_t st) {
    B200_REQUIRE(x && gamma && beta && y && M > 0 && C > 0 && C <= LN_MAXC, "layernorm_fwd: bad arguments (C <= %d)", LN_MAXC);
    long long want = (M + 7) / 8;
    const long long cap = (long long)b200_num_sms() * 16;
    const int grid = (int)(want < cap ? want : cap);
    if (C <= 96) layernorm_fwd_kernel<3><<<grid, 256, 0, st>>>(x, gamma, beta, y, stats, M, C, eps);
    else if (C <= 192) layernorm_fwd_kernel<6><<<grid, 256, 0, st>>>(x, gamma, beta, y, stats, M, C, eps);
    else if (C <= 384) layernorm_fwd_kernel<12><<<grid, 256, 0, st>>>(x, gamma, beta, y, stats, M, C, eps);
    else if (C <= 768) layernorm_fwd_kernel<24><<<grid, 256, 0, st>>>(x, gamma, beta, y, stats, M, C, eps);
    else layernorm_fwd_kernel<48><<<grid, 256, 0, st>>>(x, gamma, beta, y, stats, M, C, eps);
    B200_CHECK_LAUNCH("layernorm_fwd");
    return B200_OK;
}

B200_API int b200_layernorm_bwd(const float* x, const float* stats, const float* gamma, const float* dy, float* dx,
                                float* dgamma, float* dbeta, int accumulate_dx, long long M, int C, float* workspace,
                                long long workspace_bytes, cudaStream_t st) {
    B200_REQUIRE(x && stats && gamma && dy && dx && dgamma && dbeta && workspace && M > 0 && C > 0 && C <= LN_MAXC,
                 "layernorm_bwd: bad arguments");
    B200_REQUIRE(workspace_bytes >= b200_layernorm_workspace_bytes(M, C), "layernorm_bwd: workspace too small");
    const int grid = ln_grid(M);
    const size_t smem = (size_t)8 * 2 * C * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(layernorm_bwd_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 768 * (int)sizeof(float));
        cudaFuncSetAttribute(layernorm_bwd_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * LN_MAXC * (int)sizeof(float));
        attr_done = true;
    }
    if (C <= 96) layernorm_bwd_kernel<3><<<grid, 256, smem, st>>>(x, stats, gamma, dy, dx, workspace, M, C, accumulate_dx);
    else if (C <= 192) layernorm_bwd_kernel<6><<<grid, 256, smem, st>>>(x, stats, gamma, dy, dx, workspace, M, C, accumulate_dx);
    else if (C <= 384) layernorm_bwd_kernel<12><<<grid, 256, smem, st>>>(x, stats, gamma, dy, dx, workspace, M, C, accumulate_dx);
    else if (C <= 768) layernorm_bwd_kernel<24><<<grid, 256, smem, st>>>(x, stats, gamma, dy, dx, workspace, M, C, accumulate_dx);
    else layernorm_bwd_kernel<48><<<grid, 256, smem, st>>>(x, stats, gamma, dy, dx, workspace, M, C, accumulate_dx);
    B200_CHECK_LAUNCH("layernorm_bwd");
    colpart_finalize_kernel<<<(2 * C + 7) / 8, 256, 0, st>>>(workspace, grid, 2 * C, dgamma, dbeta, C, 0);
    B200_CHECK_LAUNCH("layernorm_bwd_finalize");
    return B200_OK;
}

// ================================================================================================ GELU (exact, erf)
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const float4 v = ldg4(x + q * 4);
        float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = 0.5f * a[i] * (1.f + erff(a[i] * 0.70710678118654752f));
        stg4(y + q * 4, make_float4(a[0], a[1], a[2], a[3]));
    }
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       float* __restrict__ dx, long long n4, int accumulate) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const float4 v = ldg4(x + q * 4), g = ldg4(dy + q * 4);
        const float a[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float cdf = 0.5f * (1.f + erff(a[i] * 0.70710678118654752f));
            const float pdf = 0.3989422804014327f * expf(-0.5f * a[i] * a[i]);
            o[i] = gg[i] * (cdf + a[i] * pdf);
        }
        if (accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(dx + q * 4);
            o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
        }
        stg4(dx + q * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
}

B200_API int b200_gelu_fwd(const float* x, float* y, long long n, cudaStream_t st) {
    B200_REQUIRE(x && y && n > 0 && (n & 3) == 0, "gelu_fwd: bad arguments (n multiple of 4)");
    gelu_fwd_kernel<<<ew_grid(n / 4), 256, 0, st>>>(x, y, n / 4);
    B200_CHECK_LAUNCH("gelu_fwd");
    return B200_OK;
}
B200_API int b200_gelu_bwd(const float* x, const float* dy, float* dx, long long n, int accumulate, cudaStream_t st) {
    B200_REQUIRE(x && dy && dx && n > 0 && (n & 3) == 0, "gelu_bwd: bad arguments (n multiple of 4)");
    gelu_bwd_kernel<<<ew_grid(n / 4), 256, 0, st>>>(x, dy, dx, n / 4, accumulate);
    B200_CHECK_LAUNCH("gelu_bwd");
    return B200_OK;
}

// ================================================================================================ window attention
// WindowAttention.forward (:115-150) inside SwinTransformerBlock.forward (:244-288).  One CTA = one (image, window, head).
// N = ws*ws tokens (<= 49), head dim 32.  Row r of the window is the token at shifted coords (wh*ws + r/ws, ww*ws + r%ws),
// i.e. original coords ((hs + shift) % H, (ws_ + shift) % W) after torch.roll(x, -shift).
#define WA_MAXN 49
#define WA_HD 32

struct WinP {
    const float* qkv;        // [B*H*W][3C]
    const float* table;      // relative_position_bias_table [(2ws-1)^2][heads]
    float* out;              // fwd: [B*H*W][C]
    const float* dout;       // bwd
    float* dqkv;             // bwd: [B*H*W][3C]
    float* dbias_part;       // bwd: [B*nW][heads][N*N] partial dS sums (per image-window), reduced by the finalize kernel
    int B, H, W, C, heads, ws, shift;
    float scale;
};

__device__ __forceinline__ int wa_token(const WinP& p, int b, int wh, int ww, int r) {
    const int hs = wh * p.ws + r / p.ws, wsx = ww * p.ws + r % p.ws;
    const int h = (hs + p.shift) % p.H, w = (wsx + p.shift) % p.W;
    return (b * p.H + h) * p.W + w;
}
// region id of the shift mask (:223-238): slices [0, H-ws), [H-ws, H-shift), [H-shift, H) of the SHIFTED coordinates
__device__ __forceinline__ int wa_region(const WinP& p, int wh, int ww, int r) {
    const int hs = wh * p.ws + r / p.ws, wsx = ww * p.ws + r % p.ws;
    const int rh = hs < p.H - p.ws ? 0 : (hs < p.H - p.shift ? 1 : 2);
    const int rw = wsx < p.W - p.ws ? 0 : (wsx < p.W - p.shift ? 1 : 2);
    return rh * 3 + rw;
}
__device__ __forceinline__ int wa_bias_index(int ws, int i, int j) {
    const int dh = i / ws - j / ws + ws - 1, dw = i % ws - j % ws + ws - 1;
    return dh * (2 * ws - 1) + dw;
}

// Register-tiled products on the shared-memory operands (one CTA = one window x head, N <= 49 tokens, head dim 32).
// The old one-output-per-thread loops issued two shared-memory loads per FMA and were bound by the LSU; a 4x7 / 4x4
// accumulator tile per thread brings that to ~0.4 loads per FMA.  Rows / columns past N are clamped on load and
// skipped on store.
#define WA_LDQ (WA_HD + 1)          // row stride of q / k / v / dO tiles
#define WA_LDP (WA_MAXN + 1)        // row stride of the N x N score tile

// acc[r][c] = sum_d A[i0 + r][d] * B[j0 + c][d]      (S = q k^T, dP = dO v^T)
__device__ __forceinline__ void wa_mm_nt_4x7(const float* __restrict__ A, const float* __restrict__ B, int N, int i0, int j0,
                                             float (&acc)[4][7]) {
    int ra[4], rb[7];
#pragma unroll
    for (int r = 0; r < 4; ++r) ra[r] = min(i0 + r, N - 1) * WA_LDQ;
#pragma unroll
    for (int c = 0; c < 7; ++c) rb[c] = min(j0 + c, N - 1) * WA_LDQ;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 7; ++c) acc[r][c] = 0.f;
#pragma unroll 4
    for (int d = 0; d < WA_HD; ++d) {
        float a[4], b[7];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = A[ra[r] + d];
#pragma unroll
        for (int c = 0; c < 7; ++c) b[c] = B[rb[c] + d];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 7; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
}
// acc[r][c] = sum_j P[i0 + r][j] * X[j][d0 + c]      (out = P v, dq = dS k)
__device__ __forceinline__ void wa_mm_nn_4x4(const float* __restrict__ P, const float* __restrict__ X, int N, int i0, int d0,
                                             float (&acc)[4][4]) {
    int rp[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) rp[r] = min(i0 + r, N - 1) * WA_LDP;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int j = 0; j < N; ++j) {
        float a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = P[rp[r] + j];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = X[j * WA_LDQ + d0 + c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
}
// acc[r][c] = sum_i P[i][j0 + r] * X[i][d0 + c]      (dv = P^T dO, dk = dS^T q)
__device__ __forceinline__ void wa_mm_tn_4x4(const float* __restrict__ P, const float* __restrict__ X, int N, int j0, int d0,
                                             float (&acc)[4][4]) {
    int cp[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) cp[r] = min(j0 + r, N - 1);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int i = 0; i < N; ++i) {
        float a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = P[i * WA_LDP + cp[r]];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = X[i * WA_LDQ + d0 + c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
}

template <bool BWD>
__global__ void __launch_bounds__(128) window_attn_kernel(const WinP p) {
    __shared__ float sq[WA_MAXN * WA_LDQ], sk[WA_MAXN * WA_LDQ], sv[WA_MAXN * WA_LDQ];
    __shared__ float sp[WA_MAXN * WA_LDP];
    __shared__ float sdo[BWD ? WA_MAXN * WA_LDQ : 1];
    __shared__ float sdot[BWD ? WA_MAXN * 8 : 1];
    __shared__ int stok[WA_MAXN], sreg[WA_MAXN];
    const int N = p.ws * p.ws;
    const int nWw = p.W / p.ws, nW = (p.H / p.ws) * nWw;
    const int head = blockIdx.x % p.heads;
    const int win = (blockIdx.x / p.heads) % nW;
    const int b = blockIdx.x / (p.heads * nW);
    const int wh = win / nWw, ww = win % nWw;
    const int tid = threadIdx.x;
    if (tid < N) {
        stok[tid] = wa_token(p, b, wh, ww, tid);
        sreg[tid] = p.shift > 0 ? wa_region(p, wh, ww, tid) : 0;
    }
    __syncthreads();
    const int C3 = 3 * p.C;
    for (int idx = tid; idx < N * WA_HD; idx += 128) {
        const int r = idx / WA_HD, d = idx % WA_HD;
        const float* row = p.qkv + (size_t)stok[r] * C3 + head * WA_HD + d;
        sq[r * WA_LDQ + d] = __ldg(row) * p.scale;
        sk[r * WA_LDQ + d] = __ldg(row + p.C);
        sv[r * WA_LDQ + d] = __ldg(row + 2 * p.C);
        if (BWD) sdo[r * WA_LDQ + d] = __ldg(p.dout + (size_t)stok[r] * p.C + head * WA_HD + d);
    }
    __syncthreads();
    // tile coordinates: 4 x 7 tiles over the N x N matrices, 4 x 4 tiles over the N x 32 ones
    const int nti = (N + 3) >> 2, ntj = (N + 6) / 7;
    const int s_ti = tid / 7, s_tj = tid % 7;                 // score tile of this thread (if s_ti < nti && s_tj < ntj)
    const bool s_on = s_ti < nti && s_tj < ntj;
    const int o_ti = tid >> 3, o_td = tid & 7;                // N x 32 tile (if o_ti < nti)
    const bool o_on = o_ti < nti;
    // scores + bias + mask
    if (s_on) {
        float acc[4][7];
        wa_mm_nt_4x7(sq, sk, N, s_ti * 4, s_tj * 7, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = s_ti * 4 + r;
            if (i >= N) break;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const int j = s_tj * 7 + c;
                if (j >= N) break;
                float s = acc[r][c] + __ldg(p.table + wa_bias_index(p.ws, i, j) * p.heads + head);
                if (p.shift > 0 && sreg[i] != sreg[j]) s += -100.0f;
                sp[i * WA_LDP + j] = s;
            }
        }
    }
    __syncthreads();
    // softmax over j (one warp per row, rows strided over the 4 warps)
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int i = warp; i < N; i += 4) {
            float* row = sp + i * WA_LDP;
            const float a0 = lane < N ? row[lane] : -INFINITY, a1 = lane + 32 < N ? row[lane + 32] : -INFINITY;
            float mx = fmaxf(a0, a1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float e0 = lane < N ? expf(a0 - mx) : 0.f, e1 = lane + 32 < N ? expf(a1 - mx) : 0.f;
            const float inv = 1.f / warp_sum(e0 + e1);
            if (lane < N) row[lane] = e0 * inv;
            if (lane + 32 < N) row[lane + 32] = e1 * inv;
        }
    }
    __syncthreads();
    if (!BWD) {
        if (o_on) {
            float acc[4][4];
            wa_mm_nn_4x4(sp, sv, N, o_ti * 4, o_td * 4, acc);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = o_ti * 4 + r;
                if (i >= N) break;
                stg4(p.out + (size_t)stok[i] * p.C + head * WA_HD + o_td * 4, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
            }
        }
        return;
    }
    // ---------------- backward: dV = P^T dO ; dP = dO V^T ; dS = P o (dP - rowsum(dP o P)) ; dq = dS k scale ; dk = dS^T q
    float* dqkv = p.dqkv;
    if (o_on) {
        float acc[4][4];
        wa_mm_tn_4x4(sp, sdo, N, o_ti * 4, o_td * 4, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = o_ti * 4 + r;
            if (j >= N) break;
            stg4(dqkv + (size_t)stok[j] * C3 + 2 * p.C + head * WA_HD + o_td * 4, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
        }
    }
    // dP tile in registers; per-row partial sums of dP o P go through shared memory (7 column tiles per row)
    float dp[4][7];
    if (s_on) {
        wa_mm_nt_4x7(sdo, sv, N, s_ti * 4, s_tj * 7, dp);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = s_ti * 4 + r;
            if (i >= N) break;
            float part = 0.f;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const int j = s_tj * 7 + c;
                if (j < N) part = fmaf(dp[r][c], sp[i * WA_LDP + j], part);
            }
            sdot[i * 8 + s_tj] = part;
        }
    }
    __syncthreads();                                            // all reads of P (dV, partial dots) are done
    if (s_on) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = s_ti * 4 + r;
            if (i >= N) break;
            float dot = 0.f;
            for (int t = 0; t < ntj; ++t) dot += sdot[i * 8 + t];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const int j = s_tj * 7 + c;
                if (j >= N) break;
                sp[i * WA_LDP + j] *= dp[r][c] - dot;              // dS in place of P
            }
        }
    }
    __syncthreads();
    if (p.dbias_part) {
        float* bp = p.dbias_part + ((size_t)(b * nW + win) * p.heads + head) * (WA_MAXN * WA_MAXN);
        for (int idx = tid; idx < N * N; idx += 128) bp[idx] = sp[(idx / N) * WA_LDP + idx % N];
    }
    if (o_on) {
        float acc[4][4];
        wa_mm_nn_4x4(sp, sk, N, o_ti * 4, o_td * 4, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = o_ti * 4 + r;
            if (i >= N) break;
            stg4(dqkv + (size_t)stok[i] * C3 + head * WA_HD + o_td * 4,
                 make_float4(acc[r][0] * p.scale, acc[r][1] * p.scale, acc[r][2] * p.scale, acc[r][3] * p.scale));
        }
        wa_mm_tn_4x4(sp, sq, N, o_ti * 4, o_td * 4, acc);          // sq already carries the scale
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = o_ti * 4 + r;
            if (j >= N) break;
            stg4(dqkv + (size_t)stok[j] * C3 + p.C + head * WA_HD + o_td * 4, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
        }
    }
}

// dtable[idx][head] = sum over (image-windows, (i,j) with bias_index(i,j) == idx) of dS, in two deterministic steps:
// (1) coalesced column sums of part[nBW][heads*49*49] over row chunks -> part2[chunk][heads*49*49];
// (2) one block per table entry: for every row i the matching column j follows from the entry's (dh, dw) directly.
#define WA_BIAS_CHUNKS 32
__global__ void __launch_bounds__(256) window_bias_colsum_kernel(const float* __restrict__ part, int nBW, int cols,
                                                                 int rows_per_chunk, float* __restrict__ part2) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    const int w0 = blockIdx.y * rows_per_chunk, w1 = min(nBW, w0 + rows_per_chunk);
    float s = 0.f;
    for (int w = w0; w < w1; ++w) s += __ldg(part + (size_t)w * cols + col);
    part2[(size_t)blockIdx.y * cols + col] = s;
}
__global__ void __launch_bounds__(64) window_bias_grad_kernel(const float* __restrict__ part2, int chunks, int heads, int ws,
                                                              float* __restrict__ dtable, int accumulate) {
    const int N = ws * ws;
    const int entry = blockIdx.x;                          // idx * heads + head
    const int idx = entry / heads, head = entry % heads;
    const int dh = idx / (2 * ws - 1) - (ws - 1), dw = idx % (2 * ws - 1) - (ws - 1);   // (row_i - row_j, col_i - col_j)
    const int cols = heads * (WA_MAXN * WA_MAXN);
    __shared__ double sh[2];
    double s = 0;
    const int i = threadIdx.x;
    if (i < N) {
        const int jr = i / ws - dh, jc = i % ws - dw;
        if (jr >= 0 && jr < ws && jc >= 0 && jc < ws) {
            const int ij = i * N + jr * ws + jc;
            for (int c = 0; c < chunks; ++c) s += (double)part2[(size_t)c * cols + head * (WA_MAXN * WA_MAXN) + ij];
        }
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double tot = sh[0] + sh[1];
        dtable[entry] = accumulate ? dtable[entry] + (float)tot : (float)tot;
    }
}

static int fill_win(WinP& p, int B, int H, int W, int C, int heads, int ws, int shift, const char* who) {
    B200_REQUIRE(B > 0 && H > 0 && W > 0 && ws > 0 && ws * ws <= WA_MAXN && H % ws == 0 && W % ws == 0, "%s: bad window geometry", who);
    B200_REQUIRE(heads > 0 && C == heads * WA_HD, "%s: head dim must be %d (C = %d, heads = %d)", who, WA_HD, C, heads);
    B200_REQUIRE(shift >= 0 && shift < ws, "%s: shift must be in [0, ws)", who);
    p.B = B; p.H = H; p.W = W; p.C = C; p.heads = heads; p.ws = ws; p.shift = shift;
    p.scale = 1.0f / sqrtf((float)WA_HD);
    return B200_OK;
}

B200_API int b200_window_attn_fwd(const float* qkv, const float* bias_table, float* out, int B, int H, int W, int C,
                                  int heads, int ws, int shift, cudaStream_t st) {
    WinP p;
    memset(&p, 0, sizeof(p));
    if (int rc = fill_win(p, B, H, W, C, heads, ws, shift, "window_attn_fwd")) return rc;
    B200_REQUIRE(qkv && bias_table && out, "window_attn_fwd: null pointer");
    p.qkv = qkv; p.table = bias_table; p.out = out;
    const int nW = (H / ws) * (W / ws);
    window_attn_kernel<false><<<B * nW * heads, 128, 0, st>>>(p);
    B200_CHECK_LAUNCH("window_attn_fwd");
    return B200_OK;
}

B200_API long long b200_window_attn_workspace_bytes(int B, int H, int W, int heads, int ws) {
    const long long nBW = (long long)B * (H / ws) * (W / ws);
    const long long chunks = nBW < WA_BIAS_CHUNKS ? nBW : WA_BIAS_CHUNKS;
    return (nBW + chunks) * heads * WA_MAXN * WA_MAXN * (long long)sizeof(float);
}

B200_API int b200_window_attn_bwd(const float* qkv, const float* bias_table, const float* dout, float* dqkv,
                                  float* dbias_table, int B, int H, int W, int C, int heads, int ws, int shift,
                                  float* workspace, long long workspace_bytes, cudaStream_t st) {
    WinP p;
    memset(&p, 0, sizeof(p));
    if (int rc = fill_win(p, B, H, W, C, heads, ws, shift, "window_attn_bwd")) return rc;
    B200_REQUIRE(qkv && bias_table && dout && dqkv && dbias_table && workspace, "window_attn_bwd: null pointer");
    B200_REQUIRE(workspace_bytes >= b200_window_attn_workspace_bytes(B, H, W, heads, ws), "window_attn_bwd: workspace too small");
    p.qkv = qkv; p.table = bias_table; p.dout = dout; p.dqkv = dqkv; p.dbias_part = workspace;
    const int nW = (H / ws) * (W / ws);
    window_attn_kernel<true><<<B * nW * heads, 128, 0, st>>>(p);
    B200_CHECK_LAUNCH("window_attn_bwd");
    const int T = (2 * ws - 1) * (2 * ws - 1);
    const int nBW = B * nW, chunks = nBW < WA_BIAS_CHUNKS ? nBW : WA_BIAS_CHUNKS;
    const int cols = heads * WA_MAXN * WA_MAXN, rows_per_chunk = (nBW + chunks - 1) / chunks;
    float* part2 = workspace + (size_t)nBW * cols;
    window_bias_colsum_kernel<<<dim3((cols + 255) / 256, chunks), 256, 0, st>>>(workspace, nBW, cols, rows_per_chunk, part2);
    B200_CHECK_LAUNCH("window_bias_colsum");
    window_bias_grad_kernel<<<T * heads, 64, 0, st>>>(part2, chunks, heads, ws, dbias_table, 0);
    B200_CHECK_LAUNCH("window_bias_grad");
    return B200_OK;
}

// ================================================================================================ DropPath residual
// out = x + branch * keep[b] / (1 - p)      (timm DropPath: per-sample Bernoulli(1 - p); identity when p == 0)
__global__ void __launch_bounds__(256) add_droppath_kernel(const float* __restrict__ x, const float* __restrict__ branch,
                                                           float* __restrict__ out, long long per_sample4, long long n4,
                                                           float p_drop, unsigned long long seed, unsigned stream,
                                                           const unsigned long long* __restrict__ seed_off) {
    if (seed_off) seed += *seed_off;
    const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        float s = 1.f;
        if (p_drop > 0.f) {
            const unsigned long long b = (unsigned long long)(q / per_sample4);
            const Philox4 r = philox4x32_10(seed, stream, b);
            s = u32_to_unit(r.x) >= p_drop ? inv_keep : 0.f;
        }
        const float4 a = x ? ldg4(x + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f), c = ldg4(branch + q * 4);
        stg4(out + q * 4, make_float4(a.x + c.x * s, a.y + c.y * s, a.z + c.z * s, a.w + c.w * s));
    }
}

// out = x (may be NULL: 0) + scale(b) * branch ; used forward (x = shortcut) and backward (x = NULL or existing grad)
B200_API int b200_add_droppath(const float* x, const float* branch, float* out, int B, long long per_sample, float p_drop,
                               unsigned long long seed, const unsigned long long* seed_offset_dev, unsigned rng_stream,
                               cudaStream_t st) {
    B200_REQUIRE(branch && out && B > 0 && per_sample > 0 && (per_sample & 3) == 0 && p_drop >= 0.f && p_drop < 1.f,
                 "add_droppath: bad arguments");
    const long long n4 = (long long)B * per_sample / 4;
    add_droppath_kernel<<<ew_grid(n4), 256, 0, st>>>(x, branch, out, per_sample / 4, n4, p_drop, seed, rng_stream, seed_offset_dev);
    B200_CHECK_LAUNCH("add_droppath");
    return B200_OK;
}

// ================================================================================================ rearrangements
// PatchMerging gather (:336-341): y[b][h2][w2][q*C + c] = x[b][2*h2 + (q & 1)][2*w2 + (q >> 1)][c]
__global__ void __launch_bounds__(256) patch_merge_gather_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H,
                                                                 int W, int C, int inverse, int accumulate, const FastDiv fdCQ,
                                                                 const FastDiv fdW2, const FastDiv fdH2) {
    const int H2 = H >> 1, W2 = W >> 1;
    const uint32_t total = (uint32_t)B * H2 * W2 * 4 * (C >> 2);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t r, cq, w2, h2, b;
        fdCQ.divmod(i, r, cq);
        const uint32_t q = r & 3; r >>= 2;
        fdW2.divmod(r, r, w2);
        fdH2.divmod(r, b, h2);
        const long long big = (((long long)b * H + 2 * h2 + (q & 1)) * W + 2 * w2 + (q >> 1)) * C + cq * 4;
        const long long small = (((long long)b * H2 + h2) * W2 + w2) * 4 * C + q * C + cq * 4;
        if (!inverse) {
            stg4(y + small, ldg4(x + big));
        } else {                                           // x = d(gathered), y = d(input)
            float4 v = ldg4(x + small);
            if (accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(y + big);
                v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
            }
            stg4(y + big, v);
        }
    }
}

B200_API int b200_patch_merge_gather(const float* x, float* y, int B, int H, int W, int C, int inverse, int accumulate,
                                     cudaStream_t st) {
    B200_REQUIRE(x && y && B > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && (C & 3) == 0, "patch_merge_gather: bad arguments");
    const long long total = (long long)B * H * W * (C / 4);
    B200_REQUIRE(total < (1ll << 32) - (1 << 24), "patch_merge_gather: tensor too large");
    FastDiv fdCQ, fdW2, fdH2;
    fdCQ.init(C / 4); fdW2.init(W / 2); fdH2.init(H / 2);
    patch_merge_gather_kernel<<<ew_grid(total), 256, 0, st>>>(x, y, B, H, W, C, inverse, accumulate, fdCQ, fdW2, fdH2);
    B200_CHECK_LAUNCH("patch_merge_gather");
    return B200_OK;
}

// PatchExpand rearrange (:378-379, :405-407): y[b][h*p + p1][w*p + p2][c] = x[b][h][w][(p1*p + p2)*C + c]
struct ShufP {
    FastDiv fdCQ, fdWP, fdHP, fdP;
};
// index math in 32 bits with precomputed reciprocals: the 64-bit divisions of the first version made this copy
// compute-bound (0.05 of the HBM peak on the final x4 expand)
__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W,
                                                            int C, int P, int inverse, const ShufP q) {
    const uint32_t total = (uint32_t)B * H * P * W * P * (C >> 2);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t r, cq, wo, ho, b, h, p1, w, p2;
        q.fdCQ.divmod(i, r, cq);
        q.fdWP.divmod(r, r, wo);
        q.fdHP.divmod(r, b, ho);
        q.fdP.divmod(ho, h, p1);
        q.fdP.divmod(wo, w, p2);
        const long long fine = (long long)i * 4;
        const long long coarse = ((((long long)b * H + h) * W + w) * P * P + p1 * P + p2) * C + cq * 4;
        if (!inverse) stg4(y + fine, ldg4(x + coarse));
        else stg4(y + coarse, ldg4(x + fine));
    }
}

B200_API int b200_pixel_shuffle(const float* x, float* y, int B, int H, int W, int C, int p, int inverse, cudaStream_t st) {
    B200_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0 && p > 0, "pixel_shuffle: bad arguments");
    const long long total = (long long)B * H * W * p * p * (C / 4);
    B200_REQUIRE(total < (1ll << 32) - (1 << 24), "pixel_shuffle: tensor too large");
    ShufP q;
    q.fdCQ.init(C / 4); q.fdWP.init(W * p); q.fdHP.init(H * p); q.fdP.init(p);
    pixel_shuffle_kernel<<<ew_grid(total), 256, 0, st>>>(x, y, B, H, W, C, p, inverse, q);
    B200_CHECK_LAUNCH("pixel_shuffle");
    return B200_OK;
}

// PatchEmbed im2col (:573,585 with vision_transformer.py:49-50 x.repeat(1,3,1,1)): rows = patches, columns (c, kh, kw)
__global__ void __launch_bounds__(256) patch_embed_gather_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H,
                                                                 int W, int P, int CH) {
    const int PH = H / P, PW = W / P, K = CH * P * P;
    const long long total = (long long)B * PH * PW * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        long long r = i / K;
        const int pw = (int)(r % PW); r /= PW;
        const int ph = (int)(r % PH);
        const int b = (int)(r / PH);
        const int kh = (k / P) % P, kw = k % P;            // the channel index (k / (P*P)) reads the same single-channel image
        y[i] = __ldg(x + ((long long)b * H + ph * P + kh) * W + pw * P + kw);
    }
}

B200_API int b200_patch_embed_gather(const float* x, float* y, int B, int H, int W, int patch, int repeat_channels,
                                     cudaStream_t st) {
    B200_REQUIRE(x && y && B > 0 && patch > 0 && H % patch == 0 && W % patch == 0 && repeat_channels > 0,
                 "patch_embed_gather: bad arguments");
    patch_embed_gather_kernel<<<ew_grid((long long)B * H * W * repeat_channels), 256, 0, st>>>(x, y, B, H, W, patch, repeat_channels);
    B200_CHECK_LAUNCH("patch_embed_gather");
    return B200_OK;
}
