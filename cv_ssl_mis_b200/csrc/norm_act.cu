// Train-mode BatchNorm (+ LeakyReLU/ReLU + Dropout) over channels-last activations, forward and backward.
//
// Replaces nn.BatchNorm2d/3d -> nn.LeakyReLU/ReLU -> nn.Dropout/Dropout3d of the reference
// (code/networks/unet.py:38-43, code/networks/vnet.py:16-25,177,196,226). Statistics are the
// biased batch variance for normalisation and the unbiased one for running_var (momentum 0.1,
// eps 1e-5), exactly like torch.nn.BatchNorm in train mode. Dropout masks come from Philox keyed by
// (seed, stream, element) so backward regenerates them instead of storing them.
#include "common.cuh"
#include "../../include/b200ssl.h"

// bn state layout: [mean | invstd | scale | shift], each C floats
#define BN_MEAN(s, C) (s)
#define BN_INVSTD(s, C) ((s) + (C))
#define BN_SCALE(s, C) ((s) + 2 * (C))
#define BN_SHIFT(s, C) ((s) + 3 * (C))

static inline int stats_grid(long long M, int RP) {
    long long want = (M + RP - 1) / RP;
    long long cap = (long long)b200_num_sms() * 4;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

// ---------------------------------------------------------------- per-channel sums
// MODE 0: (sum y, sum y^2)            MODE 1: (sum g, sum g*xhat) with g = da * dropmask * act'(z)
template <int MODE>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ y, const float* __restrict__ da,
                                                        const float* __restrict__ state, long long M, int C,
                                                        float slope, float p_drop, int drop_mode, unsigned long long seed,
                                                        unsigned stream, long long spatial, double* __restrict__ part,
                                                        const unsigned long long* __restrict__ seed_off) {
    if (seed_off) seed += *seed_off;
    extern __shared__ double sred[];                    // [RP][2][C]
    const int CQ = C >> 2;
    const int RP = 256 / CQ;
    const int tid = threadIdx.x;
    const int cq = tid % CQ, pr = tid / CQ;
    double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
    if (pr < RP) {
        float4 sc, sh, mu, is;
        if (MODE == 1) {
            sc = ldg4(BN_SCALE(state, C) + cq * 4);
            sh = ldg4(BN_SHIFT(state, C) + cq * 4);
            mu = ldg4(BN_MEAN(state, C) + cq * 4);
            is = ldg4(BN_INVSTD(state, C) + cq * 4);
        }
        const float keep_scale = (MODE == 1 && drop_mode != 0) ? 1.f / (1.f - p_drop) : 1.f;
        float f0[4] = {0, 0, 0, 0}, f1[4] = {0, 0, 0, 0};
        int cnt = 0;
        // 4 rows per trip: all loads are issued before the first use (memory-level parallelism)
        const long long stride = (long long)gridDim.x * RP;
        for (long long mb = (long long)blockIdx.x * RP + pr; mb < M; mb += 4 * stride) {
          float4 vv[4], gvv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
              const long long m = mb + u * stride;
              if (m < M) {
                  vv[u] = ldg4_stream(y + m * C + cq * 4);
                  if (MODE == 1) gvv[u] = ldg4_stream(da + m * C + cq * 4);
              }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const long long m = mb + u * stride;
            if (m >= M) break;
            const long long e = m * C + cq * 4;
            float4 v = vv[u];
            float a[4] = {v.x, v.y, v.z, v.w};
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { f0[i] += a[i]; f1[i] += a[i] * a[i]; }
            } else {
                float4 gv = gvv[u];
                float gg[4] = {gv.x, gv.y, gv.z, gv.w};
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
                const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, isv[4] = {is.x, is.y, is.z, is.w};
                bool keep[4] = {true, true, true, true};
                if (drop_mode == 1) dropout_keep4(seed, stream, (unsigned long long)(e >> 2), p_drop, keep);
                else if (drop_mode == 2) dropout_keep4(seed, stream, (unsigned long long)(((m / spatial) * C + cq * 4) >> 2), p_drop, keep);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float z = a[i] * scv[i] + shv[i];
                    float gz = keep[i] ? gg[i] * keep_scale : 0.f;
                    gz = z > 0.f ? gz : gz * slope;
                    f0[i] += gz;
                    f1[i] += gz * ((a[i] - muv[i]) * isv[i]);
                }
            }
          }
          if (++cnt == 8) {       // flush the fp32 running sums into fp64 every 32 rows
#pragma unroll
              for (int i = 0; i < 4; ++i) { d0[i] += f0[i]; d1[i] += f1[i]; f0[i] = 0.f; f1[i] = 0.f; }
              cnt = 0;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { d0[i] += f0[i]; d1[i] += f1[i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sred[(pr * 2 + 0) * C + cq * 4 + i] = d0[i];
            sred[(pr * 2 + 1) * C + cq * 4 + i] = d1[i];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < 2 * C; idx += 256) {
        double s = 0;
        for (int r = 0; r < RP; ++r) s += sred[r * 2 * C + idx];
        part[(size_t)blockIdx.x * 2 * C + idx] = s;
    }
}

// ---------------------------------------------------------------- finalize forward statistics
// one 256-thread block per channel: every thread takes <= 3 of the per-block partials, so the dependent-load depth
// is ~3 instead of nblk/32 (these finalize kernels were 10 us each -- pure latency -- with a warp per channel);
// fixed summation order -> deterministic.  Valid in thread 0 after the call.
__device__ __forceinline__ void reduce_partials(const double* __restrict__ part, int nblk, int C, int c, double& s0,
                                                double& s1) {
    __shared__ double sh[2][8];
    double a = 0, b = 0;
    for (int k = threadIdx.x; k < nblk; k += 256) {
        a += part[(size_t)k * 2 * C + c];
        b += part[(size_t)k * 2 * C + C + c];
    }
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    s0 = s1 = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) { s0 += sh[0][w]; s1 += sh[1][w]; }
}

__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* __restrict__ part, int nblk, long long M, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ state) {
    const int c = blockIdx.x;
    double s, ss;
    reduce_partials(part, nblk, C, c, s, ss);
    if (threadIdx.x == 0) {
        const double mean = s / (double)M;
        double var = ss / (double)M - mean * mean;
        if (var < 0) var = 0;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        BN_MEAN(state, C)[c] = (float)mean;
        BN_INVSTD(state, C)[c] = invstd;
        BN_SCALE(state, C)[c] = g * invstd;
        BN_SHIFT(state, C)[c] = b - (float)mean * g * invstd;
        if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        if (running_var) {
            const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    }
}

// eval mode: scale/shift from the running statistics
__global__ void bn_eval_state_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                     const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                     float* __restrict__ state) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const float invstd = 1.f / sqrtf(running_var[c] + eps);
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        BN_MEAN(state, C)[c] = running_mean[c];
        BN_INVSTD(state, C)[c] = invstd;
        BN_SCALE(state, C)[c] = g * invstd;
        BN_SHIFT(state, C)[c] = b - running_mean[c] * g * invstd;
    }
}

// ---------------------------------------------------------------- finalize backward sums
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const double* __restrict__ part, int nblk, long long M, int C,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                       float* __restrict__ coef) {
    const int c = blockIdx.x;
    double s, sx;
    reduce_partials(part, nblk, C, c, s, sx);
    if (threadIdx.x == 0) {
        if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
        if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)sx : (float)sx;
        coef[c] = (float)(s / (double)M);
        coef[C + c] = (float)(sx / (double)M);
    }
}

// ---------------------------------------------------------------- elementwise sweeps
// Both sweeps are pure HBM streams.  The grid-stride (blocks * 256 float4) is a multiple of C / 4 whenever C / 4 divides
// 256 (every channel count of the reference networks), so a thread keeps the SAME four channels for its whole life:
// the per-channel constants are loaded once, and each trip issues U independent 16-byte loads before the first use
// (one outstanding load per thread caps a 2048-thread SM near 4 TB/s; measured 3.7 -> see profiles/).
constexpr int EW_U = 4;

__device__ __forceinline__ void drop_keep(bool (&keep)[4], int drop_mode, unsigned long long seed, unsigned stream, long long q,
                                          int CQ, int cq, int C, long long spatial, float p_drop) {
    keep[0] = keep[1] = keep[2] = keep[3] = true;
    if (drop_mode == 1) dropout_keep4(seed, stream, (unsigned long long)q, p_drop, keep);
    else if (drop_mode == 2) {
        const long long m = q / CQ;
        dropout_keep4(seed, stream, (unsigned long long)(((m / spatial) * C + cq * 4) >> 2), p_drop, keep);
    }
}

// forward: a = dropout(act(y * scale + shift))
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ state,
                                                         float* __restrict__ a, long long total4, int C, float slope,
                                                         float p_drop, int drop_mode, unsigned long long seed,
                                                         unsigned stream, long long spatial,
                                                         const unsigned long long* __restrict__ seed_off) {
    if (seed_off) seed += *seed_off;
    const float keep_scale = drop_mode != 0 ? 1.f / (1.f - p_drop) : 1.f;
    const int CQ = C >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool fixed = (stride % CQ) == 0;                  // this thread always sees the same four channels
    int cq = (int)(q0 % CQ);
    float4 sc = ldg4(BN_SCALE(state, C) + cq * 4), sh = ldg4(BN_SHIFT(state, C) + cq * 4);
    for (long long qb = q0; qb < total4; qb += EW_U * stride) {
        float4 v[EW_U];
#pragma unroll
        for (int u = 0; u < EW_U; ++u) {
            const long long q = qb + u * stride;
            if (q < total4) v[u] = ldg4_stream(y + q * 4);
        }
#pragma unroll
        for (int u = 0; u < EW_U; ++u) {
            const long long q = qb + u * stride;
            if (q >= total4) break;
            if (!fixed) {
                cq = (int)(q % CQ);
                sc = ldg4(BN_SCALE(state, C) + cq * 4); sh = ldg4(BN_SHIFT(state, C) + cq * 4);
            }
            float z[4] = {v[u].x * sc.x + sh.x, v[u].y * sc.y + sh.y, v[u].z * sc.z + sh.z, v[u].w * sc.w + sh.w};
            bool keep[4];
            drop_keep(keep, drop_mode, seed, stream, q, CQ, cq, C, spatial, p_drop);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float r = z[i] > 0.f ? z[i] : z[i] * slope;
                z[i] = keep[i] ? r * keep_scale : 0.f;
            }
            stg4(a + q * 4, make_float4(z[0], z[1], z[2], z[3]));
        }
    }
}

// backward: dy = scale * (g - mean(g) - xhat * mean(g * xhat))
__global__ void __launch_bounds__(256) bn_act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ da,
                                                         const float* __restrict__ state, const float* __restrict__ coef,
                                                         float* __restrict__ dy, long long total4, int C, float slope,
                                                         float p_drop, int drop_mode, unsigned long long seed,
                                                         unsigned stream, long long spatial,
                                                         const unsigned long long* __restrict__ seed_off) {
    if (seed_off) seed += *seed_off;
    const float keep_scale = drop_mode != 0 ? 1.f / (1.f - p_drop) : 1.f;
    const int CQ = C >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool fixed = (stride % CQ) == 0;
    int cq = (int)(q0 % CQ);
    float sc[4], sh[4], mu[4], is[4], c1[4], c2[4];
    auto load_consts = [&](int c) {
        const float4 sc4 = ldg4(BN_SCALE(state, C) + c * 4), sh4 = ldg4(BN_SHIFT(state, C) + c * 4);
        const float4 mu4 = ldg4(BN_MEAN(state, C) + c * 4), is4 = ldg4(BN_INVSTD(state, C) + c * 4);
        const float4 c14 = ldg4(coef + c * 4), c24 = ldg4(coef + C + c * 4);
        sc[0] = sc4.x; sc[1] = sc4.y; sc[2] = sc4.z; sc[3] = sc4.w;
        sh[0] = sh4.x; sh[1] = sh4.y; sh[2] = sh4.z; sh[3] = sh4.w;
        mu[0] = mu4.x; mu[1] = mu4.y; mu[2] = mu4.z; mu[3] = mu4.w;
        is[0] = is4.x; is[1] = is4.y; is[2] = is4.z; is[3] = is4.w;
        c1[0] = c14.x; c1[1] = c14.y; c1[2] = c14.z; c1[3] = c14.w;
        c2[0] = c24.x; c2[1] = c24.y; c2[2] = c24.z; c2[3] = c24.w;
    };
    load_consts(cq);
    for (long long qb = q0; qb < total4; qb += EW_U * stride) {
        float4 v[EW_U], gv[EW_U];
#pragma unroll
        for (int u = 0; u < EW_U; ++u) {
            const long long q = qb + u * stride;
            if (q < total4) { v[u] = ldg4_stream(y + q * 4); gv[u] = ldg4_stream(da + q * 4); }
        }
#pragma unroll
        for (int u = 0; u < EW_U; ++u) {
            const long long q = qb + u * stride;
            if (q >= total4) break;
            if (!fixed) { cq = (int)(q % CQ); load_consts(cq); }
            const float a[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
            bool keep[4];
            drop_keep(keep, drop_mode, seed, stream, q, CQ, cq, C, spatial, p_drop);
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float z = a[i] * sc[i] + sh[i];
                float gz = keep[i] ? gg[i] * keep_scale : 0.f;
                gz = z > 0.f ? gz : gz * slope;
                const float xhat = (a[i] - mu[i]) * is[i];
                o[i] = sc[i] * (gz - c1[i] - xhat * c2[i]);
            }
            stg4(dy + q * 4, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// keep-mask (1/0) of a dropout stream, element-wise (mode 1) or per (sample, channel) (mode 2)
__global__ void __launch_bounds__(256) dropout_mask_kernel(float* __restrict__ mask, long long total, int C, float p_drop,
                                                           int drop_mode, unsigned long long seed, unsigned stream,
                                                           long long spatial,
                                                           const unsigned long long* __restrict__ seed_off) {
    if (seed_off) seed += *seed_off;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        bool keep[4];
        long long e = i;
        if (drop_mode == 2) {
            const long long m = i / C;
            const int c = (int)(i % C);
            e = (m / spatial) * C + c;
        }
        dropout_keep4(seed, stream, (unsigned long long)(e >> 2), p_drop, keep);
        mask[i] = keep[e & 3] ? 1.f : 0.f;
    }
}

static inline int ew_grid(long long work) {
    long long blocks = (work + 255) / 256;
    long long cap = (long long)b200_num_sms() * 16;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}
// the unrolled sweeps: one resident wave (8 blocks of 256 threads per SM), EW_U float4 per thread and trip
static inline int sweep_grid(long long total4) {
    long long blocks = (total4 + 256 * EW_U - 1) / (256 * EW_U);
    long long cap = (long long)b200_num_sms() * 8;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// ================================================================ C ABI
B200_API long long b200_bn_workspace_bytes(long long M, int C) {
    if (C <= 0 || (C & 3) || C > 1024) return -1;
    const int RP = 256 / (C >> 2);
    return (long long)stats_grid(M, RP < 1 ? 1 : RP) * 2 * C * sizeof(double) + 2 * C * sizeof(float);
}

static int check_bn(long long M, int C, const char* who) {
    B200_REQUIRE(M > 0, "%s: M must be positive", who);
    B200_REQUIRE(C > 0 && (C & 3) == 0 && C <= 1024, "%s: C must be a multiple of 4 and <= 1024 (got %d)", who, C);
    return B200_OK;
}

B200_API int b200_bn_stats_fwd(const float* y, long long M, int C, const float* gamma, const float* beta, float eps,
                               float momentum, float* running_mean, float* running_var, float* state, void* workspace,
                               long long workspace_bytes, cudaStream_t st) {
    if (int rc = check_bn(M, C, "bn_stats_fwd")) return rc;
    B200_REQUIRE(y && state && workspace, "bn_stats_fwd: null pointer");
    B200_REQUIRE(workspace_bytes >= b200_bn_workspace_bytes(M, C), "bn_stats_fwd: workspace too small");
    const int RP = 256 / (C >> 2);
    const int grid = stats_grid(M, RP);
    double* part = reinterpret_cast<double*>(workspace);
    const size_t smem = (size_t)RP * 2 * C * sizeof(double);
    bn_reduce_kernel<0><<<grid, 256, smem, st>>>(y, nullptr, nullptr, M, C, 0.f, 0.f, 0, 0ull, 0u, 1, part, nullptr);
    B200_CHECK_LAUNCH("bn_stats_fwd");
    bn_finalize_kernel<<<C, 256, 0, st>>>(part, grid, M, C, gamma, beta, eps, momentum, running_mean,
                                                        running_var, state);
    B200_CHECK_LAUNCH("bn_finalize");
    return B200_OK;
}

B200_API int b200_bn_finalize(const double* partials, int nblocks, long long M, int C, const float* gamma, const float* beta, float eps,
                              float momentum, float* running_mean, float* running_var, float* state, cudaStream_t st) {
    if (int rc = check_bn(M, C, "bn_finalize")) return rc;
    B200_REQUIRE(partials && state && nblocks > 0, "bn_finalize: bad arguments");
    bn_finalize_kernel<<<C, 256, 0, st>>>(partials, nblocks, M, C, gamma, beta, eps, momentum, running_mean, running_var, state);
    B200_CHECK_LAUNCH("bn_finalize");
    return B200_OK;
}

B200_API int b200_bn_eval_state(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                                const float* running_var, float* state, cudaStream_t st) {
    B200_REQUIRE(C > 0 && running_mean && running_var && state, "bn_eval_state: bad arguments");
    bn_eval_state_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, gamma, beta, eps, running_mean, running_var, state);
    B200_CHECK_LAUNCH("bn_eval_state");
    return B200_OK;
}

B200_API int b200_bn_act_fwd(const float* y, const float* state, float* a, long long M, int C, float slope, float p_drop,
                             int drop_mode, unsigned long long seed, const unsigned long long* seed_offset_dev, unsigned stream,
                             long long spatial, cudaStream_t st) {
    if (int rc = check_bn(M, C, "bn_act_fwd")) return rc;
    B200_REQUIRE(y && state && a, "bn_act_fwd: null pointer");
    B200_REQUIRE(drop_mode >= 0 && drop_mode <= 2 && p_drop >= 0.f && p_drop < 1.f, "bn_act_fwd: bad dropout arguments");
    if (p_drop == 0.f) drop_mode = 0;
    const long long total4 = M * (C >> 2);
    bn_act_fwd_kernel<<<sweep_grid(total4), 256, 0, st>>>(y, state, a, total4, C, slope, p_drop, drop_mode, seed, stream,
                                                       spatial > 0 ? spatial : 1, seed_offset_dev);
    B200_CHECK_LAUNCH("bn_act_fwd");
    return B200_OK;
}

B200_API int b200_bn_act_bwd(const float* y, const float* da, const float* state, float* dy, float* dgamma, float* dbeta,
                             int accumulate, long long M, int C, float slope, float p_drop, int drop_mode,
                             unsigned long long seed, const unsigned long long* seed_offset_dev, unsigned stream,
                             long long spatial, void* workspace, long long workspace_bytes, cudaStream_t st) {
    if (int rc = check_bn(M, C, "bn_act_bwd")) return rc;
    B200_REQUIRE(y && da && state && dy && workspace, "bn_act_bwd: null pointer");
    B200_REQUIRE(workspace_bytes >= b200_bn_workspace_bytes(M, C), "bn_act_bwd: workspace too small");
    B200_REQUIRE(drop_mode >= 0 && drop_mode <= 2 && p_drop >= 0.f && p_drop < 1.f, "bn_act_bwd: bad dropout arguments");
    if (p_drop == 0.f) drop_mode = 0;
    if (spatial <= 0) spatial = 1;
    const int RP = 256 / (C >> 2);
    const int grid = stats_grid(M, RP);
    double* part = reinterpret_cast<double*>(workspace);
    float* coef = reinterpret_cast<float*>(part + (size_t)grid * 2 * C);
    const size_t smem = (size_t)RP * 2 * C * sizeof(double);
    bn_reduce_kernel<1><<<grid, 256, smem, st>>>(y, da, state, M, C, slope, p_drop, drop_mode, seed, stream, spatial, part, seed_offset_dev);
    B200_CHECK_LAUNCH("bn_act_bwd_reduce");
    bn_bwd_finalize_kernel<<<C, 256, 0, st>>>(part, grid, M, C, dgamma, dbeta, accumulate, coef);
    B200_CHECK_LAUNCH("bn_bwd_finalize");
    const long long total4 = M * (C >> 2);
    bn_act_bwd_kernel<<<sweep_grid(total4), 256, 0, st>>>(y, da, state, coef, dy, total4, C, slope, p_drop, drop_mode, seed,
                                                       stream, spatial, seed_offset_dev);
    B200_CHECK_LAUNCH("bn_act_bwd_apply");
    return B200_OK;
}

// materialise the keep-mask (1/0) the kernels above regenerate on the fly (parity tests inject it into the oracle)
B200_API int b200_dropout_mask(float* mask, long long M, int C, float p_drop, int drop_mode, unsigned long long seed,
                               const unsigned long long* seed_offset_dev, unsigned stream, long long spatial,
                               cudaStream_t st) {
    B200_REQUIRE(mask && M > 0 && C > 0 && (C & 3) == 0, "dropout_mask: bad arguments");
    B200_REQUIRE(drop_mode == 1 || drop_mode == 2, "dropout_mask: drop_mode must be 1 (element) or 2 (channel)");
    const long long total = M * C;
    dropout_mask_kernel<<<ew_grid(total), 256, 0, st>>>(mask, total, C, p_drop, drop_mode, seed, stream,
                                                        spatial > 0 ? spatial : 1, seed_offset_dev);
    B200_CHECK_LAUNCH("dropout_mask");
    return B200_OK;
}
