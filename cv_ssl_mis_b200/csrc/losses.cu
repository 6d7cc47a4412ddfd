// Fused semi-supervised loss: softmax + cross-entropy + batch Dice on the labeled samples and
// softmax-MSE consistency against the teacher on the unlabeled samples -- one forward kernel producing the
// reduction partials and one backward kernel producing d(loss)/d(student logits).
//
// Reference semantics (code/train_mean_teacher_2D.py:213-229, code/utils/losses.py:178-201):
//   loss_ce   = CrossEntropyLoss()(logits[:Lb], y[:Lb])                       (mean over Lb*S pixels)
//   loss_dice = mean_c [1 - (2 sum(p_c t_c) + 1e-5) / (sum(p_c^2) + sum(t_c^2) + 1e-5)]   (sums over the batch)
//   cons      = mean((softmax(student[Lb:]) - softmax(teacher))^2)            (mean over U*C*S elements)
//   loss      = 0.5 (loss_dice + loss_ce) + w * cons
#include "common.cuh"
#include "../../include/b200ssl.h"

#define SSL_MAXC 8

struct LossGeom {
    const float* logits;    // student [B][C][S] (nchw) or [B][S][C] (nhwc)
    const float* teacher;   // teacher [U][C][S] / [U][S][C] or null
    const void* labels;     // [Lb][S] uint8 or int64
    int label_i64;
    int nhwc;
    int B, Lb, C;
    long long S;
};

template <int C>
__device__ __forceinline__ void load_logits(const float* base, int nhwc, long long n, long long s, long long S, int Crt,
                                            float (&z)[C]) {
    if (nhwc) {
        const float* p = base + (n * S + s) * Crt;
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = c < Crt ? __ldg(p + c) : -INFINITY;
    } else {
        const float* p = base + n * Crt * S + s;
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = c < Crt ? __ldg(p + (long long)c * S) : -INFINITY;
    }
}

template <int C>
__device__ __forceinline__ float softmax_inplace(float (&z)[C], float& lse) {
    float mx = z[0];
#pragma unroll
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, z[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = expf(z[c] - mx); sum += z[c]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < C; ++c) z[c] *= inv;
    lse = mx + logf(sum);
    return mx;
}

__device__ __forceinline__ int load_label(const void* labels, int i64, long long idx) {
    return i64 ? (int)reinterpret_cast<const long long*>(labels)[idx] : (int)reinterpret_cast<const unsigned char*>(labels)[idx];
}

// accumulators per block: [0] ce, [1] mse, [2..2+C) I_c, [2+C..) Z_c, [2+2C..) Y_c
#define SSL_NACC(C) (2 + 3 * (C))

template <int C>
__global__ void __launch_bounds__(256) ssl_loss_fwd_kernel(const LossGeom g, double* __restrict__ part) {
    constexpr int NA = SSL_NACC(C);
    float acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.f;
    const long long total = (long long)g.B * g.S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / g.S, s = idx - n * g.S;
        float p[C];
        load_logits<C>(g.logits, g.nhwc, n, s, g.S, g.C, p);
        float raw[C];
#pragma unroll
        for (int c = 0; c < C; ++c) raw[c] = p[c];
        float lse;
        softmax_inplace<C>(p, lse);
        if (n < g.Lb) {
            const int t = load_label(g.labels, g.label_i64, n * g.S + s);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                acc[2 + C + c] += p[c] * p[c];
                if (c == t) {
                    acc[0] += lse - raw[c];
                    acc[2 + c] += p[c];
                    acc[2 + 2 * C + c] += 1.f;
                }
            }
        } else if (g.teacher) {
            float q[C];
            load_logits<C>(g.teacher, g.nhwc, n - g.Lb, s, g.S, g.C, q);
            float lse2;
            softmax_inplace<C>(q, lse2);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (c < g.C) { const float d = p[c] - q[c]; acc[1] += d * d; }
            }
        }
    }
    __shared__ float sred[8][NA];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) sred[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NA) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += sred[w][threadIdx.x];
        part[(size_t)blockIdx.x * NA + threadIdx.x] = s;
    }
}

// out: [0] ce  [1] dice  [2] cons  [3] total  [4..4+C) A_c  [4+C..4+2C) B_c   (A/B = dice-gradient coefficients)
__global__ void ssl_loss_finalize_kernel(const double* __restrict__ part, int nblk, int Cpad, int C, int Lb, int U,
                                         long long S, int has_teacher, const float* __restrict__ w_cons,
                                         float* __restrict__ out) {
    __shared__ double tot[SSL_NACC(SSL_MAXC)];
    const int NA = SSL_NACC(Cpad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;      // one warp per accumulator
    if (warp < NA) {
        double s = 0;
        for (int b = lane; b < nblk; b += 32) s += part[(size_t)b * NA + warp];
        s = warp_sum_d(s);
        if (lane == 0) tot[warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double smooth = 1e-5;
        float ce = 0.f, dice = 0.f, cons = 0.f;
        if (Lb > 0) {
            ce = (float)(tot[0] / ((double)Lb * (double)S));
            double dsum = 0;
            for (int c = 0; c < C; ++c) {
                const double I = tot[2 + c], Z = tot[2 + Cpad + c], Y = tot[2 + 2 * Cpad + c];
                const double D = Z + Y + smooth;
                dsum += 1.0 - (2.0 * I + smooth) / D;
                out[4 + c] = (float)(2.0 * (2.0 * I + smooth) / ((double)C * D * D));
                out[4 + C + c] = (float)(2.0 / ((double)C * D));
            }
            dice = (float)(dsum / (double)C);
        } else {
            for (int c = 0; c < C; ++c) { out[4 + c] = 0.f; out[4 + C + c] = 0.f; }
        }
        if (has_teacher && U > 0) cons = (float)(tot[1] / ((double)U * (double)C * (double)S));
        const float w = w_cons ? w_cons[0] : 0.f;
        out[0] = ce;
        out[1] = dice;
        out[2] = cons;
        out[3] = 0.5f * (dice + ce) + w * cons;
    }
}

template <int C>
__global__ void __launch_bounds__(256) ssl_loss_bwd_kernel(const LossGeom g, const float* __restrict__ lossbuf,
                                                           const float* __restrict__ w_cons, float gscale,
                                                           float* __restrict__ dlogits, int out_nhwc) {
    float A[C], Bc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        A[c] = c < g.C ? lossbuf[4 + c] : 0.f;
        Bc[c] = c < g.C ? lossbuf[4 + g.C + c] : 0.f;
    }
    const int U = g.B - g.Lb;
    const float w = (w_cons && g.teacher) ? w_cons[0] : 0.f;
    const float ce_scale = g.Lb > 0 ? 1.f / ((float)g.Lb * (float)g.S) : 0.f;
    const float mse_scale = U > 0 ? 2.f * w / ((float)U * (float)g.C * (float)g.S) : 0.f;
    const long long total = (long long)g.B * g.S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / g.S, s = idx - n * g.S;
        float p[C], dz[C];
        load_logits<C>(g.logits, g.nhwc, n, s, g.S, g.C, p);
        float lse;
        softmax_inplace<C>(p, lse);
        if (n < g.Lb) {
            const int t = load_label(g.labels, g.label_i64, n * g.S + s);
            float gd[C], dot = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                gd[c] = A[c] * p[c] - (c == t ? Bc[c] : 0.f);
                dot += gd[c] * p[c];
            }
#pragma unroll
            for (int c = 0; c < C; ++c)
                dz[c] = 0.5f * gscale * (p[c] * (gd[c] - dot) + (p[c] - (c == t ? 1.f : 0.f)) * ce_scale);
        } else if (g.teacher && mse_scale != 0.f) {
            float q[C];
            load_logits<C>(g.teacher, g.nhwc, n - g.Lb, s, g.S, g.C, q);
            float lse2;
            softmax_inplace<C>(q, lse2);
            float gd[C], dot = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                gd[c] = c < g.C ? mse_scale * (p[c] - q[c]) : 0.f;
                dot += gd[c] * p[c];
            }
#pragma unroll
            for (int c = 0; c < C; ++c) dz[c] = gscale * p[c] * (gd[c] - dot);
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) dz[c] = 0.f;
        }
        if (out_nhwc) {
            float* o = dlogits + (n * g.S + s) * g.C;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (c < g.C) o[c] = dz[c];
        } else {
            float* o = dlogits + n * g.C * g.S + s;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (c < g.C) o[(long long)c * g.S] = dz[c];
        }
    }
}

static inline int loss_grid(long long total) {
    long long blocks = (total + 255) / 256;
    long long cap = (long long)b200_num_sms() * 4;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}
static inline int cpad_of(int C) { return C <= 2 ? 2 : (C <= 4 ? 4 : SSL_MAXC); }

B200_API long long b200_ssl_loss_workspace_bytes(int B, long long S) {
    return (long long)loss_grid((long long)B * S) * SSL_NACC(SSL_MAXC) * sizeof(double);
}

static int fill_geom(LossGeom& g, const float* logits, const float* teacher, const void* labels, int label_dtype,
                     int layout_nhwc, int B, int Lb, int C, long long S, const char* who) {
    B200_REQUIRE(logits != nullptr, "%s: null logits", who);
    B200_REQUIRE(B > 0 && Lb >= 0 && Lb <= B && S > 0, "%s: bad batch geometry (B=%d Lb=%d)", who, B, Lb);
    B200_REQUIRE(C >= 2 && C <= SSL_MAXC, "%s: classes must be in [2,%d]", who, SSL_MAXC);
    B200_REQUIRE(Lb == 0 || labels != nullptr, "%s: labeled samples need labels", who);
    B200_REQUIRE(label_dtype == B200_LABEL_U8 || label_dtype == B200_LABEL_I64, "%s: label dtype must be u8 or i64", who);
    g.logits = logits; g.teacher = teacher; g.labels = labels;
    g.label_i64 = label_dtype == B200_LABEL_I64;
    g.nhwc = layout_nhwc; g.B = B; g.Lb = Lb; g.C = C; g.S = S;
    return B200_OK;
}

B200_API int b200_ssl_loss_fwd(const float* logits, const float* teacher_logits, const void* labels, int label_dtype,
                               int layout_nhwc, int B, int Lb, int C, long long S, const float* w_cons, float* lossbuf,
                               void* workspace, long long workspace_bytes, cudaStream_t st) {
    LossGeom g;
    if (int rc = fill_geom(g, logits, teacher_logits, labels, label_dtype, layout_nhwc, B, Lb, C, S, "ssl_loss_fwd")) return rc;
    B200_REQUIRE(lossbuf && workspace, "ssl_loss_fwd: null output/workspace");
    B200_REQUIRE(workspace_bytes >= b200_ssl_loss_workspace_bytes(B, S), "ssl_loss_fwd: workspace too small");
    const int grid = loss_grid((long long)B * S);
    double* part = reinterpret_cast<double*>(workspace);
    const int Cp = cpad_of(C);
    if (Cp == 2) ssl_loss_fwd_kernel<2><<<grid, 256, 0, st>>>(g, part);
    else if (Cp == 4) ssl_loss_fwd_kernel<4><<<grid, 256, 0, st>>>(g, part);
    else ssl_loss_fwd_kernel<SSL_MAXC><<<grid, 256, 0, st>>>(g, part);
    B200_CHECK_LAUNCH("ssl_loss_fwd");
    ssl_loss_finalize_kernel<<<1, 1024, 0, st>>>(part, grid, Cp, C, Lb, B - Lb, S, teacher_logits != nullptr, w_cons, lossbuf);
    B200_CHECK_LAUNCH("ssl_loss_finalize");
    return B200_OK;
}

B200_API int b200_ssl_loss_bwd(const float* logits, const float* teacher_logits, const void* labels, int label_dtype,
                               int layout_nhwc, int B, int Lb, int C, long long S, const float* w_cons,
                               const float* lossbuf, float grad_scale, float* dlogits, int dlogits_nhwc, cudaStream_t st) {
    LossGeom g;
    if (int rc = fill_geom(g, logits, teacher_logits, labels, label_dtype, layout_nhwc, B, Lb, C, S, "ssl_loss_bwd")) return rc;
    B200_REQUIRE(lossbuf && dlogits, "ssl_loss_bwd: null pointer");
    const int grid = loss_grid((long long)B * S);
    const int Cp = cpad_of(C);
    if (Cp == 2) ssl_loss_bwd_kernel<2><<<grid, 256, 0, st>>>(g, lossbuf, w_cons, grad_scale, dlogits, dlogits_nhwc);
    else if (Cp == 4) ssl_loss_bwd_kernel<4><<<grid, 256, 0, st>>>(g, lossbuf, w_cons, grad_scale, dlogits, dlogits_nhwc);
    else ssl_loss_bwd_kernel<SSL_MAXC><<<grid, 256, 0, st>>>(g, lossbuf, w_cons, grad_scale, dlogits, dlogits_nhwc);
    B200_CHECK_LAUNCH("ssl_loss_bwd");
    return B200_OK;
}
