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
    // uncertainty-aware variant (code/train_uncertainty_aware_mean_teacher_3D.py:161-179): sum over T stochastic
    // teacher passes of softmax probabilities, [U][C][S] in the logits' layout; null => plain mean-teacher MSE
    const float* mc_psum;
    float mc_T;
    const float* mc_thr;    // DEVICE scalar: entropy threshold
    // cross-teaching variant (code/train_cross_teaching_between_cnn_transformer_2D.py:238-246): `teacher` holds the OTHER
    // model's logits for all B samples (its own layout); the unlabeled term is Dice against their argmax
    int pseudo;
    int teacher_nhwc;
};

// mask = [ -sum_c pbar_c log(pbar_c + 1e-6) < thr ],  pbar = psum / T
template <int C>
__device__ __forceinline__ bool mc_mask(const LossGeom& g, long long u, long long s) {
    float ps[C];
    if (g.nhwc) {
        const float* p = g.mc_psum + (u * g.S + s) * g.C;
#pragma unroll
        for (int c = 0; c < C; ++c) ps[c] = c < g.C ? __ldg(p + c) : 0.f;
    } else {
        const float* p = g.mc_psum + u * g.C * g.S + s;
#pragma unroll
        for (int c = 0; c < C; ++c) ps[c] = c < g.C ? __ldg(p + (long long)c * g.S) : 0.f;
    }
    const float invT = 1.f / g.mc_T;
    float H = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (c < g.C) { const float pb = ps[c] * invT; H -= pb * logf(pb + 1e-6f); }
    return H < __ldg(g.mc_thr);
}

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

// accumulators per block: [0] ce, [1] mse, [2..2+C) I_c, [2+C..) Z_c, [2+2C..) Y_c, [2+3C] mask count,
// pseudo-label mode only: [3+3C..) I'_c, [3+4C..) Z'_c, [3+5C..) Y'_c over the unlabeled samples
#define SSL_NACC(C) (3 + 3 * (C))
#define SSL_NACC_PSEUDO(C) (3 + 6 * (C))

template <int C>
__device__ __forceinline__ int argmax_first(const float (&q)[C]) {
    int t = 0;
#pragma unroll
    for (int c = 1; c < C; ++c) t = q[c] > q[t] ? c : t;
    return t;
}

template <int C, int PSEUDO>
__global__ void __launch_bounds__(256) ssl_loss_fwd_kernel(const LossGeom g, double* __restrict__ part) {
    constexpr int NA = PSEUDO ? SSL_NACC_PSEUDO(C) : SSL_NACC(C);
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
        } else if (PSEUDO) {
            float q[C];
            load_logits<C>(g.teacher, g.teacher_nhwc, n, s, g.S, g.C, q);
            float lse2;
            softmax_inplace<C>(q, lse2);
            const int t = argmax_first<C>(q);
            if (g.pseudo == 2) {                  // cross pseudo supervision: CE against the other model's argmax
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c == t) acc[1] += lse - raw[c];
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    acc[(PSEUDO ? 3 + 4 * C : 0) + c] += p[c] * p[c];
                    if (c == t) {
                        acc[(PSEUDO ? 3 + 3 * C : 0) + c] += p[c];
                        acc[(PSEUDO ? 3 + 5 * C : 0) + c] += 1.f;
                    }
                }
            }
        } else if (g.teacher) {
            float q[C];
            load_logits<C>(g.teacher, g.nhwc, n - g.Lb, s, g.S, g.C, q);
            float lse2;
            softmax_inplace<C>(q, lse2);
            const bool on = g.mc_psum == nullptr || mc_mask<C>(g, n - g.Lb, s);
            if (on) {
                acc[2 + 3 * C] += 1.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if (c < g.C) { const float d = p[c] - q[c]; acc[1] += d * d; }
                }
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

// batch Dice over one accumulator triple; writes the gradient coefficients A_c, B_c scaled by `gw`
__device__ double dice_from_sums(const double* I, const double* Z, const double* Y, int C, double gw, float* A, float* Bc) {
    const double smooth = 1e-5;
    double dsum = 0;
    for (int c = 0; c < C; ++c) {
        const double D = Z[c] + Y[c] + smooth;
        dsum += 1.0 - (2.0 * I[c] + smooth) / D;
        A[c] = (float)(gw * 2.0 * (2.0 * I[c] + smooth) / ((double)C * D * D));
        Bc[c] = (float)(gw * 2.0 / ((double)C * D));
    }
    return dsum / (double)C;
}

// out: [0] ce  [1] dice  [2] cons  [3] total  [4..4+C) A_c  [4+C..4+2C) B_c  [4+2C] d(cons)/d(p) scale
//      (A/B = dice-gradient coefficients); pseudo-label mode: [2] = Dice against the pseudo labels and
//      [5+2C..5+3C) A'_c, [5+3C..5+4C) B'_c = its gradient coefficients, already multiplied by w
__global__ void __launch_bounds__(1024) ssl_loss_finalize_kernel(const double* __restrict__ part, int nblk, int Cpad, int C, int Lb, int U,
                                         long long S, int has_teacher, int mc_mode, int pseudo,
                                         const float* __restrict__ w_cons, float* __restrict__ out) {
    __shared__ double tot[SSL_NACC_PSEUDO(SSL_MAXC)];
    const int NA = pseudo ? SSL_NACC_PSEUDO(Cpad) : SSL_NACC(Cpad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;      // one warp per accumulator
    for (int a = warp; a < NA; a += 32) {
        double s = 0;
        for (int b = lane; b < nblk; b += 32) s += part[(size_t)b * NA + a];
        s = warp_sum_d(s);
        if (lane == 0) tot[a] = s;
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
        const float w = w_cons ? w_cons[0] : 0.f;
        float dscale = 0.f;
        if (pseudo == 2) {
            // CE over the U*S unlabeled pixels; [4+2C] carries w / (U S) for the backward kernel
            for (int c = 0; c < 2 * C; ++c) out[5 + 2 * C + c] = 0.f;
            if (U > 0) {
                const double denom = (double)U * (double)S;
                cons = (float)(tot[1] / denom);
                dscale = (float)((double)w / denom);
            }
        } else if (pseudo) {
            if (U > 0)
                cons = (float)dice_from_sums(tot + 3 + 3 * Cpad, tot + 3 + 4 * Cpad, tot + 3 + 5 * Cpad, C, (double)w,
                                             out + 5 + 2 * C, out + 5 + 3 * C);
            else
                for (int c = 0; c < 2 * C; ++c) out[5 + 2 * C + c] = 0.f;
        } else if (has_teacher && U > 0) {
            // mean over all elements (MT) or sum(mask * dist) / (2 sum(mask) + 1e-16) (UAMT: the reference divides by
            // 2 * sum(mask) whatever the class count)
            const double denom = mc_mode ? 2.0 * tot[2 + 3 * Cpad] + 1e-16 : (double)U * (double)C * (double)S;
            cons = (float)(tot[1] / denom);
            dscale = (float)(2.0 * (double)w / denom);
        }
        out[4 + 2 * C] = dscale;
        out[0] = ce;
        out[1] = dice;
        out[2] = cons;
        out[3] = 0.5f * (dice + ce) + w * cons;
    }
}

template <int C, int PSEUDO>
__global__ void __launch_bounds__(256) ssl_loss_bwd_kernel(const LossGeom g, const float* __restrict__ lossbuf,
                                                           const float* __restrict__ w_cons, float gscale,
                                                           float* __restrict__ dlogits, int out_nhwc) {
    float A[C], Bc[C], A2[C], B2[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        A[c] = c < g.C ? lossbuf[4 + c] : 0.f;
        Bc[c] = c < g.C ? lossbuf[4 + g.C + c] : 0.f;
        A2[c] = PSEUDO && c < g.C ? lossbuf[5 + 2 * g.C + c] : 0.f;
        B2[c] = PSEUDO && c < g.C ? lossbuf[5 + 3 * g.C + c] : 0.f;
    }
    const float ce_scale = g.Lb > 0 ? 1.f / ((float)g.Lb * (float)g.S) : 0.f;
    const float mse_scale = g.teacher && !PSEUDO ? lossbuf[4 + 2 * g.C] : 0.f;
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
        } else if (PSEUDO) {
            float q[C];
            load_logits<C>(g.teacher, g.teacher_nhwc, n, s, g.S, g.C, q);
            float lse2;
            softmax_inplace<C>(q, lse2);
            const int t = argmax_first<C>(q);
            if (g.pseudo == 2) {
                const float sc = gscale * lossbuf[4 + 2 * g.C];
#pragma unroll
                for (int c = 0; c < C; ++c) dz[c] = sc * (p[c] - (c == t ? 1.f : 0.f));
            } else {
                float gd[C], dot = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    gd[c] = A2[c] * p[c] - (c == t ? B2[c] : 0.f);
                    dot += gd[c] * p[c];
                }
#pragma unroll
                for (int c = 0; c < C; ++c) dz[c] = gscale * p[c] * (gd[c] - dot);
            }
        } else if (g.teacher && mse_scale != 0.f && (g.mc_psum == nullptr || mc_mask<C>(g, n - g.Lb, s))) {
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
    long long cap = (long long)b200_num_sms() * 8;      // 2048 resident threads per SM: the per-pixel exp / log chains need the warps
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}
static inline int cpad_of(int C) { return C <= 2 ? 2 : (C <= 4 ? 4 : SSL_MAXC); }

B200_API long long b200_ssl_loss_workspace_bytes(int B, long long S) {
    return (long long)loss_grid((long long)B * S) * SSL_NACC_PSEUDO(SSL_MAXC) * sizeof(double);
}

static int fill_geom(LossGeom& g, const float* logits, const float* teacher, const void* labels, int label_dtype,
                     int layout_nhwc, int B, int Lb, int C, long long S, const float* mc_psum, float mc_T,
                     const float* mc_thr, const char* who) {
    B200_REQUIRE(logits != nullptr, "%s: null logits", who);
    B200_REQUIRE(B > 0 && Lb >= 0 && Lb <= B && S > 0, "%s: bad batch geometry (B=%d Lb=%d)", who, B, Lb);
    B200_REQUIRE(C >= 2 && C <= SSL_MAXC, "%s: classes must be in [2,%d]", who, SSL_MAXC);
    B200_REQUIRE(Lb == 0 || labels != nullptr, "%s: labeled samples need labels", who);
    B200_REQUIRE(label_dtype == B200_LABEL_U8 || label_dtype == B200_LABEL_I64, "%s: label dtype must be u8 or i64", who);
    g.logits = logits; g.teacher = teacher; g.labels = labels;
    g.label_i64 = label_dtype == B200_LABEL_I64;
    g.nhwc = layout_nhwc; g.B = B; g.Lb = Lb; g.C = C; g.S = S;
    B200_REQUIRE(mc_psum == nullptr || (teacher != nullptr && mc_thr != nullptr && mc_T > 0.f),
                 "%s: the uncertainty mask needs teacher logits, a threshold and T > 0", who);
    g.mc_psum = mc_psum; g.mc_T = mc_T; g.mc_thr = mc_thr;
    g.pseudo = 0; g.teacher_nhwc = layout_nhwc;
    return B200_OK;
}

template <int PSEUDO>
static void launch_loss_fwd(const LossGeom& g, int Cp, int grid, double* part, cudaStream_t st) {
    if (Cp == 2) ssl_loss_fwd_kernel<2, PSEUDO><<<grid, 256, 0, st>>>(g, part);
    else if (Cp == 4) ssl_loss_fwd_kernel<4, PSEUDO><<<grid, 256, 0, st>>>(g, part);
    else ssl_loss_fwd_kernel<SSL_MAXC, PSEUDO><<<grid, 256, 0, st>>>(g, part);
}
template <int PSEUDO>
static void launch_loss_bwd(const LossGeom& g, int Cp, int grid, const float* lossbuf, const float* w_cons, float gs,
                            float* dlogits, int dl_nhwc, cudaStream_t st) {
    if (Cp == 2) ssl_loss_bwd_kernel<2, PSEUDO><<<grid, 256, 0, st>>>(g, lossbuf, w_cons, gs, dlogits, dl_nhwc);
    else if (Cp == 4) ssl_loss_bwd_kernel<4, PSEUDO><<<grid, 256, 0, st>>>(g, lossbuf, w_cons, gs, dlogits, dl_nhwc);
    else ssl_loss_bwd_kernel<SSL_MAXC, PSEUDO><<<grid, 256, 0, st>>>(g, lossbuf, w_cons, gs, dlogits, dl_nhwc);
}

B200_API int b200_ssl_loss_fwd(const float* logits, const float* teacher_logits, const void* labels, int label_dtype,
                               int layout_nhwc, int B, int Lb, int C, long long S, const float* w_cons,
                               const float* mc_psum, float mc_T, const float* mc_thr, float* lossbuf, void* workspace,
                               long long workspace_bytes, cudaStream_t st) {
    LossGeom g;
    if (int rc = fill_geom(g, logits, teacher_logits, labels, label_dtype, layout_nhwc, B, Lb, C, S, mc_psum, mc_T, mc_thr,
                           "ssl_loss_fwd")) return rc;
    B200_REQUIRE(lossbuf && workspace, "ssl_loss_fwd: null output/workspace");
    B200_REQUIRE(workspace_bytes >= b200_ssl_loss_workspace_bytes(B, S), "ssl_loss_fwd: workspace too small");
    const int grid = loss_grid((long long)B * S);
    double* part = reinterpret_cast<double*>(workspace);
    const int Cp = cpad_of(C);
    launch_loss_fwd<0>(g, Cp, grid, part, st);
    B200_CHECK_LAUNCH("ssl_loss_fwd");
    ssl_loss_finalize_kernel<<<1, 1024, 0, st>>>(part, grid, Cp, C, Lb, B - Lb, S, teacher_logits != nullptr, mc_psum != nullptr,
                                                 0, w_cons, lossbuf);
    B200_CHECK_LAUNCH("ssl_loss_finalize");
    return B200_OK;
}

B200_API int b200_ssl_loss_bwd(const float* logits, const float* teacher_logits, const void* labels, int label_dtype,
                               int layout_nhwc, int B, int Lb, int C, long long S, const float* w_cons,
                               const float* mc_psum, float mc_T, const float* mc_thr, const float* lossbuf,
                               float grad_scale, float* dlogits, int dlogits_nhwc, cudaStream_t st) {
    LossGeom g;
    if (int rc = fill_geom(g, logits, teacher_logits, labels, label_dtype, layout_nhwc, B, Lb, C, S, mc_psum, mc_T, mc_thr,
                           "ssl_loss_bwd")) return rc;
    B200_REQUIRE(lossbuf && dlogits, "ssl_loss_bwd: null pointer");
    const int grid = loss_grid((long long)B * S);
    const int Cp = cpad_of(C);
    launch_loss_bwd<0>(g, Cp, grid, lossbuf, w_cons, grad_scale, dlogits, dlogits_nhwc, st);
    B200_CHECK_LAUNCH("ssl_loss_bwd");
    return B200_OK;
}

// Cross-teaching loss of ONE model (code/train_cross_teaching_between_cnn_transformer_2D.py:229-247):
//   0.5 (CE + Dice)(logits[:Lb], y) + w * Dice(softmax(logits[Lb:]), argmax softmax(other[Lb:]))
// lossbuf (>= 5 + 4C floats): [0] ce [1] dice [2] pseudo-label dice [3] total, then gradient coefficients
static int pseudo_loss_fwd(int kind, const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc,
                           const void* labels, int label_dtype, int B, int Lb, int C, long long S, const float* w_cons,
                           float* lossbuf, void* workspace, long long workspace_bytes, cudaStream_t st) {
    LossGeom g;
    if (int rc = fill_geom(g, logits, other_logits, labels, label_dtype, layout_nhwc, B, Lb, C, S, nullptr, 0.f, nullptr,
                           "ct_loss_fwd")) return rc;
    B200_REQUIRE(other_logits && lossbuf && workspace && w_cons, "ct_loss_fwd: null pointer");
    B200_REQUIRE(workspace_bytes >= b200_ssl_loss_workspace_bytes(B, S), "ct_loss_fwd: workspace too small");
    g.pseudo = kind; g.teacher_nhwc = other_nhwc;
    const int grid = loss_grid((long long)B * S);
    const int Cp = cpad_of(C);
    launch_loss_fwd<1>(g, Cp, grid, reinterpret_cast<double*>(workspace), st);
    B200_CHECK_LAUNCH("ct_loss_fwd");
    ssl_loss_finalize_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<double*>(workspace), grid, Cp, C, Lb, B - Lb, S, 1, 0, kind, w_cons,
                                                 lossbuf);
    B200_CHECK_LAUNCH("ct_loss_finalize");
    return B200_OK;
}

B200_API int b200_ct_loss_fwd(const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc,
                              const void* labels, int label_dtype, int B, int Lb, int C, long long S, const float* w_cons,
                              float* lossbuf, void* workspace, long long workspace_bytes, cudaStream_t st) {
    return pseudo_loss_fwd(1, logits, layout_nhwc, other_logits, other_nhwc, labels, label_dtype, B, Lb, C, S, w_cons, lossbuf,
                           workspace, workspace_bytes, st);
}

// Cross-pseudo-supervision loss of ONE model (code/train_cross_pseudo_supervision_2D.py:187-196):
//   0.5 (CE + Dice)(logits[:Lb], y) + w * CE(logits[Lb:], argmax softmax(other[Lb:]))
// lossbuf: [0] ce [1] dice [2] pseudo-label CE [3] total, then gradient coefficients (same layout as ct_loss)
B200_API int b200_cps_loss_fwd(const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc,
                               const void* labels, int label_dtype, int B, int Lb, int C, long long S, const float* w_cons,
                               float* lossbuf, void* workspace, long long workspace_bytes, cudaStream_t st) {
    return pseudo_loss_fwd(2, logits, layout_nhwc, other_logits, other_nhwc, labels, label_dtype, B, Lb, C, S, w_cons, lossbuf,
                           workspace, workspace_bytes, st);
}

static int pseudo_loss_bwd(int kind, const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc,
                           const void* labels, int label_dtype, int B, int Lb, int C, long long S, const float* lossbuf,
                           float grad_scale, float* dlogits, int dlogits_nhwc, cudaStream_t st) {
    LossGeom g;
    if (int rc = fill_geom(g, logits, other_logits, labels, label_dtype, layout_nhwc, B, Lb, C, S, nullptr, 0.f, nullptr,
                           "ct_loss_bwd")) return rc;
    B200_REQUIRE(other_logits && lossbuf && dlogits, "ct_loss_bwd: null pointer");
    g.pseudo = kind; g.teacher_nhwc = other_nhwc;
    launch_loss_bwd<1>(g, cpad_of(C), loss_grid((long long)B * S), lossbuf, nullptr, grad_scale, dlogits, dlogits_nhwc, st);
    B200_CHECK_LAUNCH("ct_loss_bwd");
    return B200_OK;
}

B200_API int b200_ct_loss_bwd(const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc,
                              const void* labels, int label_dtype, int B, int Lb, int C, long long S, const float* lossbuf,
                              float grad_scale, float* dlogits, int dlogits_nhwc, cudaStream_t st) {
    return pseudo_loss_bwd(1, logits, layout_nhwc, other_logits, other_nhwc, labels, label_dtype, B, Lb, C, S, lossbuf, grad_scale,
                           dlogits, dlogits_nhwc, st);
}

B200_API int b200_cps_loss_bwd(const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc,
                               const void* labels, int label_dtype, int B, int Lb, int C, long long S, const float* lossbuf,
                               float grad_scale, float* dlogits, int dlogits_nhwc, cudaStream_t st) {
    return pseudo_loss_bwd(2, logits, layout_nhwc, other_logits, other_nhwc, labels, label_dtype, B, Lb, C, S, lossbuf, grad_scale,
                           dlogits, dlogits_nhwc, st);
}

// psum[u] (+)= sum_r softmax(logits[r*U + u])  -- the T stochastic teacher passes of UAMT
// (code/train_uncertainty_aware_mean_teacher_3D.py:149-163: preds.reshape(T, stride, ...).mean(0), never materialised)
template <int C>
__global__ void __launch_bounds__(256) mc_softmax_acc_kernel(const float* __restrict__ logits, float* __restrict__ psum, int R,
                                                             int U, int Crt, long long S, int nhwc, int init) {
    const long long total = (long long)U * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long u = idx / S, s = idx - u * S;
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        for (int r = 0; r < R; ++r) {
            float p[C];
            load_logits<C>(logits, nhwc, (long long)r * U + u, s, S, Crt, p);
            float lse;
            softmax_inplace<C>(p, lse);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] += p[c];
        }
        if (nhwc) {
            float* o = psum + (u * S + s) * Crt;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (c < Crt) o[c] = init ? acc[c] : o[c] + acc[c];
        } else {
            float* o = psum + u * Crt * S + s;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (c < Crt) o[(long long)c * S] = init ? acc[c] : o[(long long)c * S] + acc[c];
        }
    }
}

B200_API int b200_mc_softmax_accumulate(const float* logits, float* psum, int R, int U, int C, long long S, int layout_nhwc,
                                        int init, cudaStream_t st) {
    B200_REQUIRE(logits && psum && R > 0 && U > 0 && S > 0, "mc_softmax_accumulate: bad arguments");
    B200_REQUIRE(C >= 2 && C <= SSL_MAXC, "mc_softmax_accumulate: classes must be in [2,%d]", SSL_MAXC);
    const int grid = loss_grid((long long)U * S);
    const int Cp = cpad_of(C);
    if (Cp == 2) mc_softmax_acc_kernel<2><<<grid, 256, 0, st>>>(logits, psum, R, U, C, S, layout_nhwc, init);
    else if (Cp == 4) mc_softmax_acc_kernel<4><<<grid, 256, 0, st>>>(logits, psum, R, U, C, S, layout_nhwc, init);
    else mc_softmax_acc_kernel<SSL_MAXC><<<grid, 256, 0, st>>>(logits, psum, R, U, C, S, layout_nhwc, init);
    B200_CHECK_LAUNCH("mc_softmax_accumulate");
    return B200_OK;
}
