// Fused SGD(momentum, weight decay) + teacher EMA over flat parameter buffers, and the teacher-input noise.
//
// Reference semantics:
//   optim.SGD(lr, momentum=0.9, weight_decay=1e-4)       code/train_mean_teacher_2D.py:189-190
//       g += wd * p ; buf = mu * buf + g ; p -= lr * buf   (buf starts at 0, so the first step gives buf = g)
//   update_ema_variables                                   code/train_mean_teacher_2D.py:124-128
//       alpha = min(1 - 1/(step+1), ema_decay) ; ema = alpha * ema + (1 - alpha) * p     (p already updated)
//   noise = clamp(randn_like(x) * 0.1, -0.2, 0.2)          code/train_mean_teacher_2D.py:208-210
// Per-step scalars live in a small device array so a captured CUDA graph picks up new values each replay.
#include "common.cuh"
#include "../../include/b200ssl.h"

// hp: [0] lr  [1] momentum  [2] weight_decay  [3] ema_alpha  [4] 1 - ema_alpha  [5] grad_scale
__global__ void __launch_bounds__(256) sgd_ema_kernel(float* __restrict__ p, float* g, float* __restrict__ buf,
                                                      float* __restrict__ ema, long long n4, long long n,
                                                      const float* __restrict__ hp, int zero_grad) {
    const float lr = hp[0], mu = hp[1], wd = hp[2], alpha = hp[3], oma = hp[4], gs = hp[5];
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const long long e = q * 4;
        if (e + 4 <= n) {
            float4 pv = *reinterpret_cast<const float4*>(p + e);
            const float4 gv = *reinterpret_cast<const float4*>(g + e);
            float4 bv = *reinterpret_cast<const float4*>(buf + e);
            bv.x = mu * bv.x + (gv.x * gs + wd * pv.x);
            bv.y = mu * bv.y + (gv.y * gs + wd * pv.y);
            bv.z = mu * bv.z + (gv.z * gs + wd * pv.z);
            bv.w = mu * bv.w + (gv.w * gs + wd * pv.w);
            pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
            stg4(buf + e, bv);
            stg4(p + e, pv);
            if (ema) {
                float4 ev = *reinterpret_cast<const float4*>(ema + e);
                ev.x = ev.x * alpha + oma * pv.x;
                ev.y = ev.y * alpha + oma * pv.y;
                ev.z = ev.z * alpha + oma * pv.z;
                ev.w = ev.w * alpha + oma * pv.w;
                stg4(ema + e, ev);
            }
            if (zero_grad) stg4(g + e, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
            for (long long i = e; i < n; ++i) {
                float b = mu * buf[i] + (g[i] * gs + wd * p[i]);
                float pv = p[i] - lr * b;
                buf[i] = b;
                p[i] = pv;
                if (ema) ema[i] = ema[i] * alpha + oma * pv;
                if (zero_grad) g[i] = 0.f;
            }
        }
    }
}

// ema = alpha * ema + (1 - alpha) * p  (standalone form of update_ema_variables)
__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ ema, const float* __restrict__ p, long long n,
                                                  const float* __restrict__ hp) {
    const float alpha = hp[3], oma = hp[4];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        ema[i] = ema[i] * alpha + oma * p[i];
}

// out = x + clamp(sigma * N(0,1), -clip, clip); x may be null (pure noise)
__global__ void __launch_bounds__(256) noise_kernel(const float* __restrict__ x, float* __restrict__ out, long long n,
                                                    float sigma, float clip, unsigned long long seed, unsigned stream,
                                                    const unsigned long long* __restrict__ seed_off) {
    if (seed_off) seed += *seed_off;
    const long long n4 = (n + 3) / 4;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const Philox4 r = philox4x32_10(seed, stream, (unsigned long long)q);
        // Box-Muller on two pairs of uniforms (u in (0,1])
        const float u0 = 1.f - u32_to_unit(r.x), u1 = u32_to_unit(r.y);
        const float u2 = 1.f - u32_to_unit(r.z), u3 = u32_to_unit(r.w);
        const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
        float s0, c0, s1, c1;
        sincospif(2.f * u1, &s0, &c0);
        sincospif(2.f * u3, &s1, &c1);
        float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long e = q * 4 + i;
            if (e < n) {
                const float nz = fminf(fmaxf(z[i] * sigma, -clip), clip);
                out[e] = (x ? x[e] : 0.f) + nz;
            }
        }
    }
}

static inline int ew_grid(long long work) {
    long long blocks = (work + 255) / 256;
    long long cap = (long long)b200_num_sms() * 16;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

B200_API int b200_sgd_ema_step(float* params, float* grads, float* momentum_buf, float* ema_params, long long n,
                               const float* hparams_dev, int zero_grad, cudaStream_t st) {
    B200_REQUIRE(params && grads && momentum_buf && hparams_dev && n > 0, "sgd_ema_step: bad arguments");
    B200_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)momentum_buf | (uintptr_t)ema_params) & 15) == 0,
                 "sgd_ema_step: buffers must be 16-byte aligned");
    const long long n4 = (n + 3) / 4;
    sgd_ema_kernel<<<ew_grid(n4), 256, 0, st>>>(params, grads, momentum_buf, ema_params, n4, n, hparams_dev, zero_grad);
    B200_CHECK_LAUNCH("sgd_ema_step");
    return B200_OK;
}

B200_API int b200_ema_update(float* ema_params, const float* params, long long n, const float* hparams_dev, cudaStream_t st) {
    B200_REQUIRE(ema_params && params && hparams_dev && n > 0, "ema_update: bad arguments");
    ema_kernel<<<ew_grid(n), 256, 0, st>>>(ema_params, params, n, hparams_dev);
    B200_CHECK_LAUNCH("ema_update");
    return B200_OK;
}

B200_API int b200_noise_add(const float* x, float* out, long long n, float sigma, float clip, unsigned long long seed,
                            const unsigned long long* seed_offset_dev, unsigned stream, cudaStream_t st) {
    B200_REQUIRE(out && n > 0 && sigma >= 0.f && clip >= 0.f, "noise_add: bad arguments");
    noise_kernel<<<ew_grid((n + 3) / 4), 256, 0, st>>>(x, out, n, sigma, clip, seed, stream, seed_offset_dev);
    B200_CHECK_LAUNCH("noise_add");
    return B200_OK;
}
