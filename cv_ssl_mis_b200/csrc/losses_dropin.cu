// Stand-alone loss entry points behind the reference's own function signatures (code/utils/losses.py:74-113,165-201):
//   DiceLoss.forward(inputs, target, weight, softmax)          -> b200_dice_fwd / b200_dice_bwd
//   softmax_mse_loss(input_logits, target_logits)              -> b200_softmax_mse_fwd / _bwd   (element-wise result)
//   softmax_kl_loss(input_logits, target_logits)               -> b200_softmax_kl_fwd / _bwd    (F.kl_div(..., 'mean'))
// A reference trainer that only swaps `from utils import losses` keeps its autograd graph: these are the forward /
// backward halves of torch.autograd.Functions (cv_ssl_mis_b200/utils/losses.py).  The fused trainer does not use them --
// it runs CE + Dice + consistency in ONE pass (losses.cu).  Tensors are [B][C][S] fp32 (NCHW / NCDHW), C <= 8; one
// thread per pixel reads its C values with stride S (coalesced across the warp), reductions go through per-block fp64
// partials summed in a fixed order (deterministic).
#include "common.cuh"
#include "../../include/b200ssl.h"

namespace {

constexpr int MAXC = 8;

__device__ __forceinline__ int label_at(const void* labels, int i64, long long idx) {
    return i64 ? (int)reinterpret_cast<const long long*>(labels)[idx] : (int)reinterpret_cast<const unsigned char*>(labels)[idx];
}

__device__ __forceinline__ void load_c(const float* __restrict__ base, long long n, long long s, long long S, int C, float (&z)[MAXC]) {
    const float* p = base + n * C * S + s;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) z[c] = c < C ? __ldg(p + (long long)c * S) : -INFINITY;
}

__device__ __forceinline__ void softmax_c(float (&z)[MAXC], int C, float& lse) {
    float mx = z[0];
#pragma unroll
    for (int c = 1; c < MAXC; ++c) if (c < C) mx = fmaxf(mx, z[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) { z[c] = c < C ? expf(z[c] - mx) : 0.f; sum += z[c]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) z[c] *= inv;
    lse = mx + logf(sum);
}

// block reduction of NV doubles per thread into part[blockIdx.x][NV]
template <int NV>
__device__ __forceinline__ void block_partials(double (&v)[NV], double* __restrict__ part) {
    __shared__ double sh[8][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum_d(v[i]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[warp][i] = v[i];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
        part[(size_t)blockIdx.x * NV + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------------ Dice
// per class: I = sum(score * t), Z = sum(score^2), Y = sum(t^2) over the whole batch (losses.py:178-186)
__global__ void __launch_bounds__(256) dice_sums_kernel(const float* __restrict__ x, int use_softmax, const void* __restrict__ labels, int i64,
                                                        int B, int C, long long S, double* __restrict__ part) {
    double acc[3 * MAXC];
#pragma unroll
    for (int i = 0; i < 3 * MAXC; ++i) acc[i] = 0.0;
    const long long total = (long long)B * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / S, s = idx - n * S;
        float z[MAXC];
        load_c(x, n, s, S, C, z);
        float lse;
        if (use_softmax) softmax_c(z, C, lse);
        const int t = label_at(labels, i64, idx);
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                const float sc = z[c], tt = t == c ? 1.f : 0.f;
                acc[c] += (double)(sc * tt);
                acc[MAXC + c] += (double)(sc * sc);
                acc[2 * MAXC + c] += (double)tt;
            }
    }
    block_partials<3 * MAXC>(acc, part);
}

// out: [0] loss, [1..1+C) class-wise dice (1 - loss_c), [1+MAXC..) I, Z, Y (3 * MAXC floats) for the backward pass
__global__ void dice_finalize_kernel(const double* __restrict__ part, int nblk, int C, const float* __restrict__ weight, float* __restrict__ out) {
    __shared__ double sums[3 * MAXC];
    if (threadIdx.x < 3 * MAXC) {
        double s = 0;
        for (int k = 0; k < nblk; ++k) s += part[(size_t)k * 3 * MAXC + threadIdx.x];
        sums[threadIdx.x] = s;
        out[1 + MAXC + threadIdx.x] = (float)s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double loss = 0;
        for (int c = 0; c < C; ++c) {
            const double d = 1.0 - (2.0 * sums[c] + 1e-5) / (sums[MAXC + c] + sums[2 * MAXC + c] + 1e-5);
            out[1 + c] = (float)(1.0 - d);
            loss += d * (weight ? (double)weight[c] : 1.0);
        }
        out[0] = (float)(loss / C);
    }
}

// d loss / d x: g_c = (w_c / C) * ( -(2 t D_c - N_c 2 score) / D_c^2 ), through the softmax when it was applied inside
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ x, int use_softmax, const void* __restrict__ labels, int i64,
                                                       int B, int C, long long S, const float* __restrict__ stats, const float* __restrict__ weight,
                                                       const float* __restrict__ gout, float* __restrict__ dx) {
    float Nn[MAXC], Dd[MAXC], wc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const float I = c < C ? stats[1 + MAXC + c] : 0.f, Z = c < C ? stats[1 + 2 * MAXC + c] : 0.f, Y = c < C ? stats[1 + 3 * MAXC + c] : 0.f;
        Nn[c] = 2.f * I + 1e-5f;
        Dd[c] = Z + Y + 1e-5f;
        wc[c] = c < C ? (weight ? weight[c] : 1.f) / (float)C : 0.f;
    }
    const float go = __ldg(gout);
    const long long total = (long long)B * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / S, s = idx - n * S;
        float z[MAXC];
        load_c(x, n, s, S, C, z);
        float lse;
        if (use_softmax) softmax_c(z, C, lse);
        const int t = label_at(labels, i64, idx);
        float g[MAXC], dot = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            g[c] = 0.f;
            if (c < C) {
                const float tt = t == c ? 1.f : 0.f;
                g[c] = -wc[c] * (2.f * tt * Dd[c] - Nn[c] * 2.f * z[c]) / (Dd[c] * Dd[c]);
                dot += z[c] * g[c];
            }
        }
        float* o = dx + n * C * S + s;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) o[(long long)c * S] = go * (use_softmax ? z[c] * (g[c] - dot) : g[c]);
    }
}

// ------------------------------------------------------------------------------------------------ softmax MSE (element-wise)
__global__ void __launch_bounds__(256) softmax_mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int B, int C, long long S,
                                                              float* __restrict__ out) {
    const long long total = (long long)B * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / S, s = idx - n * S;
        float pa[MAXC], pb[MAXC], l;
        load_c(a, n, s, S, C, pa); softmax_c(pa, C, l);
        load_c(b, n, s, S, C, pb); softmax_c(pb, C, l);
        float* o = out + n * C * S + s;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { const float d = pa[c] - pb[c]; o[(long long)c * S] = d * d; }
    }
}

// gradient w.r.t. the input logits only (the target side carries no gradient, losses.py:80)
__global__ void __launch_bounds__(256) softmax_mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gout,
                                                              int B, int C, long long S, float* __restrict__ da) {
    const long long total = (long long)B * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / S, s = idx - n * S;
        float pa[MAXC], pb[MAXC], l;
        load_c(a, n, s, S, C, pa); softmax_c(pa, C, l);
        load_c(b, n, s, S, C, pb); softmax_c(pb, C, l);
        const float* gp = gout + n * C * S + s;
        float q[MAXC], dot = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            q[c] = c < C ? 2.f * (pa[c] - pb[c]) * __ldg(gp + (long long)c * S) : 0.f;
            dot += pa[c] * q[c];
        }
        float* o = da + n * C * S + s;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) o[(long long)c * S] = pa[c] * (q[c] - dot);
    }
}

// ------------------------------------------------------------------------------------------------ softmax KL
// F.kl_div(log_softmax(a), softmax(b), reduction='mean') = mean over ALL B*C*S elements of pb (log pb - log pa)
__global__ void __launch_bounds__(256) softmax_kl_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int B, int C, long long S,
                                                             double* __restrict__ part) {
    double acc[1] = {0.0};
    const long long total = (long long)B * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / S, s = idx - n * S;
        float za[MAXC], zb[MAXC], pb[MAXC], lsa, lsb;
        load_c(a, n, s, S, C, za);
        load_c(b, n, s, S, C, zb);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) pb[c] = zb[c];
        float pa[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) pa[c] = za[c];
        softmax_c(pa, C, lsa);
        softmax_c(pb, C, lsb);
        float kl = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C && pb[c] > 0.f) kl += pb[c] * ((zb[c] - lsb) - (za[c] - lsa));      // xlogy convention: 0 log 0 = 0
        acc[0] += (double)kl;
    }
    block_partials<1>(acc, part);
}

__global__ void kl_finalize_kernel(const double* __restrict__ part, int nblk, double inv_count, float* __restrict__ out) {
    if (threadIdx.x == 0) {
        double s = 0;
        for (int k = 0; k < nblk; ++k) s += part[k];
        out[0] = (float)(s * inv_count);
    }
}

// d/da_c = (pa_c * sum_k pb_k - pb_c) / count = (pa_c - pb_c) / count
__global__ void __launch_bounds__(256) softmax_kl_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gout,
                                                             int B, int C, long long S, float inv_count, float* __restrict__ da) {
    const float go = __ldg(gout) * inv_count;
    const long long total = (long long)B * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / S, s = idx - n * S;
        float pa[MAXC], pb[MAXC], l;
        load_c(a, n, s, S, C, pa); softmax_c(pa, C, l);
        load_c(b, n, s, S, C, pb); softmax_c(pb, C, l);
        float* o = da + n * C * S + s;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) o[(long long)c * S] = go * (pa[c] - pb[c]);
    }
}

int loss_grid(long long total) {
    const long long want = (total + 255) / 256, cap = (long long)b200_num_sms() * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

B200_API long long b200_loss_dropin_workspace_bytes(int B, long long S) {
    return (long long)loss_grid((long long)B * S) * 3 * MAXC * (long long)sizeof(double);
}

B200_API int b200_dice_fwd(const float* x, int use_softmax, const void* labels, int label_dtype, int B, int C, long long S,
                           const float* weight, float* out, void* workspace, long long workspace_bytes, cudaStream_t st) {
    B200_REQUIRE(x && labels && out && workspace && B > 0 && S > 0 && C >= 1 && C <= MAXC, "dice_fwd: bad arguments (C <= 8)");
    B200_REQUIRE(workspace_bytes >= b200_loss_dropin_workspace_bytes(B, S), "dice_fwd: workspace too small");
    const int grid = loss_grid((long long)B * S);
    double* part = reinterpret_cast<double*>(workspace);
    dice_sums_kernel<<<grid, 256, 0, st>>>(x, use_softmax, labels, label_dtype == B200_LABEL_I64, B, C, S, part);
    B200_CHECK_LAUNCH("dice_fwd");
    dice_finalize_kernel<<<1, 32, 0, st>>>(part, grid, C, weight, out);
    B200_CHECK_LAUNCH("dice_finalize");
    return B200_OK;
}

B200_API int b200_dice_bwd(const float* x, int use_softmax, const void* labels, int label_dtype, int B, int C, long long S,
                           const float* weight, const float* fwd_out, const float* grad_out, float* dx, cudaStream_t st) {
    B200_REQUIRE(x && labels && fwd_out && grad_out && dx && B > 0 && S > 0 && C >= 1 && C <= MAXC, "dice_bwd: bad arguments (C <= 8)");
    dice_bwd_kernel<<<loss_grid((long long)B * S), 256, 0, st>>>(x, use_softmax, labels, label_dtype == B200_LABEL_I64, B, C, S, fwd_out,
                                                                 weight, grad_out, dx);
    B200_CHECK_LAUNCH("dice_bwd");
    return B200_OK;
}

B200_API int b200_softmax_mse_fwd(const float* input_logits, const float* target_logits, int B, int C, long long S, float* out,
                                  cudaStream_t st) {
    B200_REQUIRE(input_logits && target_logits && out && B > 0 && S > 0 && C >= 1 && C <= MAXC, "softmax_mse_fwd: bad arguments (C <= 8)");
    softmax_mse_fwd_kernel<<<loss_grid((long long)B * S), 256, 0, st>>>(input_logits, target_logits, B, C, S, out);
    B200_CHECK_LAUNCH("softmax_mse_fwd");
    return B200_OK;
}

B200_API int b200_softmax_mse_bwd(const float* input_logits, const float* target_logits, const float* grad_out, int B, int C,
                                  long long S, float* d_input, cudaStream_t st) {
    B200_REQUIRE(input_logits && target_logits && grad_out && d_input && B > 0 && S > 0 && C >= 1 && C <= MAXC,
                 "softmax_mse_bwd: bad arguments (C <= 8)");
    softmax_mse_bwd_kernel<<<loss_grid((long long)B * S), 256, 0, st>>>(input_logits, target_logits, grad_out, B, C, S, d_input);
    B200_CHECK_LAUNCH("softmax_mse_bwd");
    return B200_OK;
}

B200_API int b200_softmax_kl_fwd(const float* input_logits, const float* target_logits, int B, int C, long long S, float* out,
                                 void* workspace, long long workspace_bytes, cudaStream_t st) {
    B200_REQUIRE(input_logits && target_logits && out && workspace && B > 0 && S > 0 && C >= 1 && C <= MAXC,
                 "softmax_kl_fwd: bad arguments (C <= 8)");
    B200_REQUIRE(workspace_bytes >= b200_loss_dropin_workspace_bytes(B, S), "softmax_kl_fwd: workspace too small");
    const int grid = loss_grid((long long)B * S);
    double* part = reinterpret_cast<double*>(workspace);
    softmax_kl_fwd_kernel<<<grid, 256, 0, st>>>(input_logits, target_logits, B, C, S, part);
    B200_CHECK_LAUNCH("softmax_kl_fwd");
    kl_finalize_kernel<<<1, 32, 0, st>>>(part, grid, 1.0 / ((double)B * C * (double)S), out);
    B200_CHECK_LAUNCH("softmax_kl_finalize");
    return B200_OK;
}

B200_API int b200_softmax_kl_bwd(const float* input_logits, const float* target_logits, const float* grad_out, int B, int C,
                                 long long S, float* d_input, cudaStream_t st) {
    B200_REQUIRE(input_logits && target_logits && grad_out && d_input && B > 0 && S > 0 && C >= 1 && C <= MAXC,
                 "softmax_kl_bwd: bad arguments (C <= 8)");
    softmax_kl_bwd_kernel<<<loss_grid((long long)B * S), 256, 0, st>>>(input_logits, target_logits, grad_out, B, C, S,
                                                                       (float)(1.0 / ((double)B * C * (double)S)), d_input);
    B200_CHECK_LAUNCH("softmax_kl_bwd");
    return B200_OK;
}
