// Kernels of the UNETR path (code/networks/unetr.py:215-230 over MONAI's ViT / SABlock / UnetResBlock):
// 3D patch gather of the "perceptron" patch embedding, global multi-head self-attention (forward + backward) for the
// short ViT sequences (N = 216 tokens at 96^3 / 16^3), and the residual add + LeakyReLU of UnetResBlock.
// Linear layers run on gemm_umma.cu, LayerNorm / GELU on swin.cu, the convolutions on the conv engine.
#include "common.cuh"
#include <cstring>
#include "../../include/b200ssl.h"

static inline int ew_grid(long long work) {
    long long blocks = (work + 255) / 256;
    long long cap = (long long)b200_num_sms() * 16;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// ================================================================================================ patch gather
// einops "b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)" (MONAI PatchEmbeddingBlock, pos_embed = "perceptron")
__global__ void __launch_bounds__(256) patch3d_gather_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C,
                                                             int D, int H, int W, int P) {
    const int nd = D / P, nh = H / P, nw = W / P;
    const long long K = (long long)P * P * P * C;
    const long long total = (long long)B * nd * nh * nw * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long k = i % K, r = i / K;
        const int c = (int)(k % C); k /= C;
        const int p3 = (int)(k % P); k /= P;
        const int p2 = (int)(k % P);
        const int p1 = (int)(k / P);
        const int tw = (int)(r % nw); r /= nw;
        const int th = (int)(r % nh); r /= nh;
        const int td = (int)(r % nd);
        const int b = (int)(r / nd);
        y[i] = __ldg(x + ((((long long)b * C + c) * D + td * P + p1) * H + th * P + p2) * W + tw * P + p3);
    }
}

B200_API int b200_patch3d_gather(const float* x, float* y, int B, int C, int D, int H, int W, int patch, cudaStream_t st) {
    B200_REQUIRE(x && y && B > 0 && C > 0 && patch > 0 && D % patch == 0 && H % patch == 0 && W % patch == 0,
                 "patch3d_gather: bad arguments");
    patch3d_gather_kernel<<<ew_grid((long long)B * C * D * H * W), 256, 0, st>>>(x, y, B, C, D, H, W, patch);
    B200_CHECK_LAUNCH("patch3d_gather");
    return B200_OK;
}

// ================================================================================================ residual + LeakyReLU
__global__ void __launch_bounds__(256) add_lrelu_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            float* __restrict__ out, long long n4, float slope) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const float4 u = ldg4(a + q * 4), v = ldg4(b + q * 4);
        float s[4] = {u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = s[i] > 0.f ? s[i] : s[i] * slope;
        stg4(out + q * 4, make_float4(s[0], s[1], s[2], s[3]));
    }
}
// dx = dout * (out > 0 ? 1 : slope): LeakyReLU keeps the sign, so its output decides the branch
__global__ void __launch_bounds__(256) lrelu_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout,
                                                        float* __restrict__ dx, long long n4, float slope) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const float4 o = ldg4(out + q * 4), g = ldg4(dout + q * 4);
        stg4(dx + q * 4, make_float4(o.x > 0.f ? g.x : g.x * slope, o.y > 0.f ? g.y : g.y * slope,
                                     o.z > 0.f ? g.z : g.z * slope, o.w > 0.f ? g.w : g.w * slope));
    }
}

B200_API int b200_add_lrelu_fwd(const float* a, const float* b, float* out, long long n, float slope, cudaStream_t st) {
    B200_REQUIRE(a && b && out && n > 0 && (n & 3) == 0, "add_lrelu_fwd: bad arguments (n multiple of 4)");
    add_lrelu_fwd_kernel<<<ew_grid(n / 4), 256, 0, st>>>(a, b, out, n / 4, slope);
    B200_CHECK_LAUNCH("add_lrelu_fwd");
    return B200_OK;
}
B200_API int b200_lrelu_bwd(const float* out, const float* dout, float* dx, long long n, float slope, cudaStream_t st) {
    B200_REQUIRE(out && dout && dx && n > 0 && (n & 3) == 0, "lrelu_bwd: bad arguments (n multiple of 4)");
    lrelu_bwd_kernel<<<ew_grid(n / 4), 256, 0, st>>>(out, dout, dx, n / 4, slope);
    B200_CHECK_LAUNCH("lrelu_bwd");
    return B200_OK;
}

// ================================================================================================ global MHA
// MONAI SABlock: qkv = Linear(C, 3C, bias=False)(x) with columns (qkv, head, d); att = softmax(q k^T * hd^-0.5);
// out = att v, heads concatenated.  One CTA = one (batch, head, block of MHA_QB queries); K and V of the head live in
// shared memory.  The probabilities are stored ([B*heads][N][N]) for the backward pass, which runs as two kernels:
// per query block  dP = dO V^T, dS = P o (dP - rowsum(dP o P)), dQ = dS K * scale   (dS stored next to P),
// per key block    dV = P^T dO, dK = dS^T Q * scale.
#define MHA_QB 32
#define MHA_MAXN 256
#define MHA_MAXHD 64

struct MhaP {
    const float* qkv;      // [B*N][3C]
    float* out;            // [B*N][C]
    float* probs;          // [B*heads][N][N]
    const float* dout;     // [B*N][C]
    float* dqkv;           // [B*N][3C]
    float* ds;             // [B*heads][N][N]
    int B, N, heads, hd, C;
    int QB;                // queries (bwd_kv: keys) per CTA
    float scale;
};

// acc[r][c] = sum_d A[i0 + r][d] * B[j0 + c][d]; rows / columns past the limits are clamped on load
template <int TR, int TC>
__device__ __forceinline__ void mha_mm_nt(const float* __restrict__ A, const float* __restrict__ B, int ld, int hd, int i0, int ni,
                                          int j0, int nj, float (&acc)[TR][TC]) {
    int ra[TR], rb[TC];
#pragma unroll
    for (int r = 0; r < TR; ++r) ra[r] = min(i0 + r, ni - 1) * ld;
#pragma unroll
    for (int c = 0; c < TC; ++c) rb[c] = min(j0 + c, nj - 1) * ld;
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int d = 0; d < hd; ++d) {
        float a[TR], b[TC];
#pragma unroll
        for (int r = 0; r < TR; ++r) a[r] = A[ra[r] + d];
#pragma unroll
        for (int c = 0; c < TC; ++c) b[c] = B[rb[c] + d];
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int c = 0; c < TC; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
}
// acc[r][c] = sum_j P[i0 + r][j] * X[j][d0 + c]
template <int TR, int TC>
__device__ __forceinline__ void mha_mm_nn(const float* __restrict__ P, int ldp, const float* __restrict__ X, int ldx, int N, int i0,
                                          int ni, int d0, float (&acc)[TR][TC]) {
    int rp[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) rp[r] = min(i0 + r, ni - 1) * ldp;
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int j = 0; j < N; ++j) {
        float a[TR], b[TC];
#pragma unroll
        for (int r = 0; r < TR; ++r) a[r] = P[rp[r] + j];
#pragma unroll
        for (int c = 0; c < TC; ++c) b[c] = X[j * ldx + d0 + c];
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int c = 0; c < TC; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
}

__global__ void __launch_bounds__(256) mha_fwd_kernel(const MhaP p) {
    extern __shared__ float sm[];
    const int QB = p.QB;
    const int N = p.N, hd = p.hd, LD = hd + 1, LS = N + 1;
    float* sk = sm;                       // [N][LD]
    float* sv = sk + N * LD;              // [N][LD]
    float* sq = sv + N * LD;              // [QB][LD]
    float* ss = sq + QB * LD;         // [QB][LS]
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int q0 = blockIdx.x * QB, nq = min(QB, N - q0);
    const int tid = threadIdx.x, C3 = 3 * p.C;
    const float* base = p.qkv + (size_t)b * N * C3 + h * hd;
    for (int idx = tid; idx < N * hd; idx += 256) {
        const int r = idx / hd, d = idx % hd;
        sk[r * LD + d] = __ldg(base + (size_t)r * C3 + p.C + d);
        sv[r * LD + d] = __ldg(base + (size_t)r * C3 + 2 * p.C + d);
    }
    for (int idx = tid; idx < nq * hd; idx += 256) {
        const int r = idx / hd, d = idx % hd;
        sq[r * LD + d] = __ldg(base + (size_t)(q0 + r) * C3 + d) * p.scale;
    }
    __syncthreads();
    // S = q k^T, one 4 x 8 accumulator tile per thread (0.375 shared loads per FMA instead of 2)
    for (int tile = tid; tile < (QB / 4) * ((N + 7) / 8); tile += 256) {
        const int i0 = (tile % (QB / 4)) * 4, j0 = (tile / (QB / 4)) * 8;
        float acc[4][8];
        mha_mm_nt<4, 8>(sq, sk, LD, hd, i0, nq, j0, N, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (i0 + r < nq && j0 + c < N) ss[(i0 + r) * LS + j0 + c] = acc[r][c];
    }
    __syncthreads();
    {
        const int lane = tid & 31, warp = tid >> 5;
        float* prow_g = p.probs + ((size_t)bh * N + q0) * N;
        for (int i = warp; i < nq; i += 8) {
            float mx = -INFINITY;
            for (int j = lane; j < N; j += 32) mx = fmaxf(mx, ss[i * LS + j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
            for (int j = lane; j < N; j += 32) {
                const float e = expf(ss[i * LS + j] - mx);
                ss[i * LS + j] = e;
                sum += e;
            }
            const float inv = 1.f / warp_sum(sum);
            for (int j = lane; j < N; j += 32) {
                const float v = ss[i * LS + j] * inv;
                ss[i * LS + j] = v;
                if (p.probs) prow_g[(size_t)i * N + j] = v;
            }
        }
    }
    __syncthreads();
    // out = P v, 2 x 4 tiles
    for (int tile = tid; tile < (QB / 2) * (hd / 4); tile += 256) {
        const int d0 = (tile % (hd / 4)) * 4, i0 = (tile / (hd / 4)) * 2;
        float acc[2][4];
        mha_mm_nn<2, 4>(ss, LS, sv, LD, N, i0, nq, d0, acc);
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (i0 + r < nq)
                stg4(p.out + ((size_t)b * N + q0 + i0 + r) * p.C + h * hd + d0, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
    }
}

__global__ void __launch_bounds__(256) mha_bwd_q_kernel(const MhaP p) {
    extern __shared__ float sm[];
    const int QB = p.QB;
    const int N = p.N, hd = p.hd, LD = hd + 1, LS = N + 1;
    float* sk = sm;
    float* sv = sk + N * LD;
    float* sdo = sv + N * LD;             // [QB][LD]
    float* ss = sdo + QB * LD;        // [QB][LS]  P, then dS
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int q0 = blockIdx.x * QB, nq = min(QB, N - q0);
    const int tid = threadIdx.x, C3 = 3 * p.C;
    const float* base = p.qkv + (size_t)b * N * C3 + h * hd;
    for (int idx = tid; idx < N * hd; idx += 256) {
        const int r = idx / hd, d = idx % hd;
        sk[r * LD + d] = __ldg(base + (size_t)r * C3 + p.C + d);
        sv[r * LD + d] = __ldg(base + (size_t)r * C3 + 2 * p.C + d);
    }
    for (int idx = tid; idx < nq * hd; idx += 256) {
        const int r = idx / hd, d = idx % hd;
        sdo[r * LD + d] = __ldg(p.dout + ((size_t)b * N + q0 + r) * p.C + h * hd + d);
    }
    const float* prow_g = p.probs + ((size_t)bh * N + q0) * N;
    for (int idx = tid; idx < nq * N; idx += 256) ss[(idx / N) * LS + idx % N] = __ldg(prow_g + idx);
    __syncthreads();
    {
        const int lane = tid & 31, warp = tid >> 5;
        float* ds_g = p.ds + ((size_t)bh * N + q0) * N;
        for (int i = warp; i < nq; i += 8) {
            float dot = 0.f;
            // dP kept in registers: up to MHA_MAXN / 32 columns per lane
            float dp[MHA_MAXN / 32];
#pragma unroll
            for (int u = 0; u < MHA_MAXN / 32; ++u) {
                const int j = lane + 32 * u;
                float v = 0.f;
                if (j < N) {
                    for (int d = 0; d < hd; ++d) v = fmaf(sdo[i * LD + d], sv[j * LD + d], v);
                    dot += v * ss[i * LS + j];
                }
                dp[u] = v;
            }
            dot = warp_sum(dot);
#pragma unroll
            for (int u = 0; u < MHA_MAXN / 32; ++u) {
                const int j = lane + 32 * u;
                if (j < N) {
                    const float v = ss[i * LS + j] * (dp[u] - dot);
                    ss[i * LS + j] = v;
                    ds_g[(size_t)i * N + j] = v;
                }
            }
        }
    }
    __syncthreads();
    // dq = dS k * scale, 2 x 4 tiles
    for (int tile = tid; tile < (QB / 2) * (hd / 4); tile += 256) {
        const int d0 = (tile % (hd / 4)) * 4, i0 = (tile / (hd / 4)) * 2;
        float acc[2][4];
        mha_mm_nn<2, 4>(ss, LS, sk, LD, N, i0, nq, d0, acc);
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (i0 + r < nq)
                stg4(p.dqkv + ((size_t)b * N + q0 + i0 + r) * C3 + h * hd + d0,
                     make_float4(acc[r][0] * p.scale, acc[r][1] * p.scale, acc[r][2] * p.scale, acc[r][3] * p.scale));
    }
}

__global__ void __launch_bounds__(256) mha_bwd_kv_kernel(const MhaP p) {
    extern __shared__ float sm[];
    const int QB = p.QB;
    const int N = p.N, hd = p.hd, LD = hd + 1;
    float* sq = sm;                       // [N][LD]  (scaled q)
    float* sdo = sq + N * LD;             // [N][LD]
    float* sp = sdo + N * LD;             // [N][QB + 1]  P columns of this key block
    float* sds = sp + N * (QB + 1);   // [N][QB + 1]  dS columns
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int j0 = blockIdx.x * QB, nk = min(QB, N - j0);
    const int tid = threadIdx.x, C3 = 3 * p.C;
    const float* base = p.qkv + (size_t)b * N * C3 + h * hd;
    for (int idx = tid; idx < N * hd; idx += 256) {
        const int r = idx / hd, d = idx % hd;
        sq[r * LD + d] = __ldg(base + (size_t)r * C3 + d) * p.scale;
        sdo[r * LD + d] = __ldg(p.dout + ((size_t)b * N + r) * p.C + h * hd + d);
    }
    for (int idx = tid; idx < N * nk; idx += 256) {
        const int i = idx / nk, j = idx % nk;
        sp[i * (QB + 1) + j] = __ldg(p.probs + ((size_t)bh * N + i) * N + j0 + j);
        sds[i * (QB + 1) + j] = __ldg(p.ds + ((size_t)bh * N + i) * N + j0 + j);
    }
    __syncthreads();
    // dv = P^T dO and dk = dS^T q (q carries the scale): 2 x 4 tiles, both products share the loop over the queries
    for (int tile = tid; tile < (QB / 2) * (hd / 4); tile += 256) {
        const int d0 = (tile % (hd / 4)) * 4, jj = (tile / (hd / 4)) * 2;
        const int c0 = min(jj, nk - 1), c1 = min(jj + 1, nk - 1);
        float dv[2][4], dk[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) { dv[r][c] = 0.f; dk[r][c] = 0.f; }
#pragma unroll 2
        for (int i = 0; i < N; ++i) {
            const float p0 = sp[i * (QB + 1) + c0], p1 = sp[i * (QB + 1) + c1];
            const float s0 = sds[i * (QB + 1) + c0], s1 = sds[i * (QB + 1) + c1];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float g = sdo[i * LD + d0 + c], q = sq[i * LD + d0 + c];
                dv[0][c] = fmaf(p0, g, dv[0][c]); dv[1][c] = fmaf(p1, g, dv[1][c]);
                dk[0][c] = fmaf(s0, q, dk[0][c]); dk[1][c] = fmaf(s1, q, dk[1][c]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (jj + r >= nk) break;
            float* o = p.dqkv + ((size_t)b * N + j0 + jj + r) * C3 + h * hd + d0;
            stg4(o + p.C, make_float4(dk[r][0], dk[r][1], dk[r][2], dk[r][3]));
            stg4(o + 2 * p.C, make_float4(dv[r][0], dv[r][1], dv[r][2], dv[r][3]));
        }
    }
}

static int fill_mha(MhaP& p, int B, int N, int heads, int hd, const char* who) {
    B200_REQUIRE(B > 0 && N > 0 && N <= MHA_MAXN && heads > 0 && hd > 0 && hd <= MHA_MAXHD && (hd & 3) == 0,
                 "%s: needs N <= %d, head dim <= %d and a multiple of 4", who, MHA_MAXN, MHA_MAXHD);
    memset(&p, 0, sizeof(p));
    p.B = B; p.N = N; p.heads = heads; p.hd = hd; p.C = heads * hd;
    p.scale = 1.0f / sqrtf((float)hd);
    return B200_OK;
}
static size_t mha_smem_q(int N, int hd, int QB) { return (size_t)(2 * N * (hd + 1) + QB * (hd + 1) + QB * (N + 1)) * sizeof(float); }
static size_t mha_smem_kv(int N, int hd, int QB) { return (size_t)(2 * N * (hd + 1) + 2 * N * (QB + 1)) * sizeof(float); }
// Queries per CTA: K and V of a head fill most of an SM's shared memory (one CTA per SM), so the grid should be ONE wave --
// the smallest block of 32 .. 64 queries whose CTA count fits the SMs (UNETR: 216 tokens x 24 (batch, head) pairs: 48 queries =
// 120 CTAs instead of 168 in two waves), as long as both kernels' shared memory fits.
static int mha_choose_qb(int B, int N, int heads, int hd) {
    for (int qb = MHA_QB; qb <= 64; qb += 8) {
        const size_t need = mha_smem_q(N, hd, qb) > mha_smem_kv(N, hd, qb) ? mha_smem_q(N, hd, qb) : mha_smem_kv(N, hd, qb);
        if (need > 227 * 1024) break;
        if ((long long)((N + qb - 1) / qb) * B * heads <= b200_num_sms()) return qb;
    }
    return MHA_QB;
}

B200_API long long b200_mha_probs_floats(int B, int N, int heads) { return (long long)B * heads * N * N; }

B200_API int b200_mha_fwd(const float* qkv, float* out, float* probs, int B, int N, int heads, int hd, cudaStream_t st) {
    MhaP p;
    if (int rc = fill_mha(p, B, N, heads, hd, "mha_fwd")) return rc;
    B200_REQUIRE(qkv && out, "mha_fwd: null pointer");
    p.qkv = qkv; p.out = out; p.probs = probs;
    p.QB = mha_choose_qb(B, N, heads, hd);
    const size_t smem = mha_smem_q(N, hd, p.QB);
    cudaFuncSetAttribute(mha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    mha_fwd_kernel<<<dim3((N + p.QB - 1) / p.QB, B * heads), 256, smem, st>>>(p);
    B200_CHECK_LAUNCH("mha_fwd");
    return B200_OK;
}

// workspace: B*heads*N*N floats (dS)
B200_API int b200_mha_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, float* workspace,
                          long long workspace_bytes, int B, int N, int heads, int hd, cudaStream_t st) {
    MhaP p;
    if (int rc = fill_mha(p, B, N, heads, hd, "mha_bwd")) return rc;
    B200_REQUIRE(qkv && probs && dout && dqkv && workspace, "mha_bwd: null pointer");
    B200_REQUIRE(workspace_bytes >= b200_mha_probs_floats(B, N, heads) * (long long)sizeof(float), "mha_bwd: workspace too small");
    p.qkv = qkv; p.probs = const_cast<float*>(probs); p.dout = dout; p.dqkv = dqkv; p.ds = workspace;
    p.QB = mha_choose_qb(B, N, heads, hd);
    cudaFuncSetAttribute(mha_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(mha_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const dim3 grid((N + p.QB - 1) / p.QB, B * heads);
    mha_bwd_q_kernel<<<grid, 256, mha_smem_q(N, hd, p.QB), st>>>(p);
    B200_CHECK_LAUNCH("mha_bwd_q");
    mha_bwd_kv_kernel<<<grid, 256, mha_smem_kv(N, hd, p.QB), st>>>(p);
    B200_CHECK_LAUNCH("mha_bwd_kv");
    return B200_OK;
}
