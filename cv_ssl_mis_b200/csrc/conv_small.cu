// Small helpers around the convolution engines:
//   * one-launch weight packing for a whole network (a device-side job table instead of ~50 tiny launches)
//   * the first layer of the CNNs (Cin = 1, 3x3 / 3x3x3, <= 32 output channels): purely HBM-bound (reads 4 B and writes
//     4*Cout B per pixel), so plain FFMA kernels with coalesced 16-byte stores beat any tensor-core formulation
//     (code/networks/unet.py:37 with in_chns = 1; code/networks/vnet.py:152 block_one).
#include "conv_common.cuh"
#include "conv_row_pack.cuh"
#include "../../include/b200ssl.h"

// ---------------------------------------------------------------------------------------------------- batched packing
// job layout (8 x int64): [0] src ptr [1] dst ptr [2] kind (0 generic, 1 tile, 2 umma, 3 row) [3] mode (generic pack mode or
// dgrad flag) [4] O [5] I [6] T [7] total output floats
__device__ __forceinline__ float pack_elem(const float* __restrict__ w, int kind, int mode, int O, int I, int T, int idx) {
    if (kind == 0) {
        int rows, cols;
        switch (mode) {
            case B200_PACK_CONV_FWD: rows = T * I; cols = O; break;
            case B200_PACK_CONV_DGRAD: rows = T * O; cols = I; break;
            case B200_PACK_CONV_DGRAD_D2S: rows = O; cols = T * I; break;
            case B200_PACK_DECONV_FWD: rows = I; cols = T * O; break;
            default: rows = T * O; cols = I; break;
        }
        (void)rows;
        const int ldn = (cols + 3) / 4 * 4;
        const int r = idx / ldn, c = idx % ldn;
        if (c >= cols) return 0.f;
        switch (mode) {
            case B200_PACK_CONV_FWD: { int tap = r / I, i = r % I; return w[((size_t)c * I + i) * T + tap]; }
            case B200_PACK_CONV_DGRAD: { int tap = r / O, o = r % O; return w[((size_t)o * I + c) * T + (T - 1 - tap)]; }
            case B200_PACK_CONV_DGRAD_D2S: { int tap = c / I, i = c % I; return w[((size_t)r * I + i) * T + tap]; }
            case B200_PACK_DECONV_FWD: { int tap = c / O, o = c % O; return w[((size_t)r * O + o) * T + tap]; }
            default: { int tap = r / O, o = r % O; return w[((size_t)c * O + o) * T + tap]; }
        }
    }
    const int dgrad = mode & 1;
    const int rows = dgrad ? O : I, cols = dgrad ? I : O;
    if (kind == 3) return row_pack_elem(w, mode, O, I, T, idx);      // row-ring kernels (conv_row_pack.cuh)
    const int colsP = (cols + 15) / 16 * 16;
    int row, col, tap;
    if (kind == 1) {            // [chunk][tap][colsP][16]
        const int kk = idx & 15;
        int r = idx >> 4;
        col = r % colsP; r /= colsP;
        tap = r % T;
        row = (r / T) * 16 + kk;
    } else {                    // [chunk][tap][kq][colsP][4]
        const int j = idx & 3;
        int r = idx >> 2;
        col = r % colsP; r /= colsP;
        const int kq = r & 3; r >>= 2;
        tap = r % T;
        row = (r / T) * 16 + kq * 4 + j;
    }
    if (row >= rows || col >= cols) return 0.f;
    const float v = dgrad ? w[((size_t)row * I + col) * T + (T - 1 - tap)] : w[((size_t)col * I + row) * T + tap];
    return __uint_as_float(f2tf32(v));
}

__global__ void __launch_bounds__(256) pack_batch_kernel(const long long* __restrict__ jobs) {
    const long long* j = jobs + (size_t)blockIdx.y * 8;
    const float* w = reinterpret_cast<const float*>(j[0]);
    float* out = reinterpret_cast<float*>(j[1]);
    const int kind = (int)j[2], mode = (int)j[3], O = (int)j[4], I = (int)j[5], T = (int)j[6];
    const int total = (int)j[7];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
        out[idx] = pack_elem(w, kind, mode, O, I, T, idx);
}

B200_API int b200_conv_pack_batch(const long long* jobs_dev, int njobs, int blocks_per_job, cudaStream_t st) {
    B200_REQUIRE(jobs_dev && njobs > 0 && blocks_per_job > 0, "conv_pack_batch: bad arguments");
    dim3 grid(blocks_per_job, njobs);
    pack_batch_kernel<<<grid, 256, 0, st>>>(jobs_dev);
    B200_CHECK_LAUNCH("conv_pack_batch");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------------- Cin = 1 layers
struct C1P {
    const float* x;          // [N][D][H][W] (one channel)
    const float* w;          // framework layout [Cout][1][T]
    const float* bias;
    float* y;                // [pixels][Cout]
    const float* dy;         // wgrad: [pixels][Cout]
    float* part;             // wgrad: [blocks][T][Cout] (+ colsum [blocks][Cout] at part_colsum)
    float* part_colsum;
    int N, D, H, W, Cout, KD;
    long long M;
    FastDiv fd_w, fd_h, fd_d;  // pixel index -> (n, d, h, w) without 64-bit divisions (M < 2^31)
};

template <int COUT, int KD>
__global__ void __launch_bounds__(256) conv_c1_fwd_kernel(const C1P p) {
    constexpr int T = KD * 9;
    __shared__ float sw[T * COUT];
    __shared__ float sb[COUT];
    for (int i = threadIdx.x; i < T * COUT; i += 256) {
        const int tap = i / COUT, co = i % COUT;
        sw[i] = p.w[co * T + tap];
    }
    if (threadIdx.x < COUT) sb[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    __syncthreads();
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < p.M; m += (long long)gridDim.x * blockDim.x) {
        long long r = m;
        const int w = (int)(r % p.W); r /= p.W;
        const int h = (int)(r % p.H); r /= p.H;
        const int d = (int)(r % p.D);
        const long long n = r / p.D;
        float xin[T];
#pragma unroll
        for (int kd = 0; kd < KD; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int id = d + kd - (KD == 3 ? 1 : 0), ih = h + kh - 1, iw = w + kw - 1;
                    const bool ok = (unsigned)id < (unsigned)p.D && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
                    xin[(kd * 3 + kh) * 3 + kw] = ok ? __ldg(p.x + ((n * p.D + id) * p.H + ih) * p.W + iw) : 0.f;
                }
        float* out = p.y + m * COUT;
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
            float a[4] = {sb[c4 * 4], sb[c4 * 4 + 1], sb[c4 * 4 + 2], sb[c4 * 4 + 3]};
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const float4 wv = *reinterpret_cast<const float4*>(&sw[t * COUT + c4 * 4]);
                a[0] = fmaf(xin[t], wv.x, a[0]); a[1] = fmaf(xin[t], wv.y, a[1]);
                a[2] = fmaf(xin[t], wv.z, a[2]); a[3] = fmaf(xin[t], wv.w, a[3]);
            }
            stg4(out + c4 * 4, make_float4(a[0], a[1], a[2], a[3]));
        }
    }
}

// dw[co][tap] = sum_pixels dy[p][co] * x[p + tap].  Thread = (pixel lane, group of 4 output channels): one 16-byte dy
// load and T cached x loads feed 4T FMAs per pixel; deterministic two-stage reduction (block partials, then the
// generic split reducer).
template <int COUT, int KD>
__global__ void __launch_bounds__(256) conv_c1_wgrad_kernel(const C1P p) {
    constexpr int T = KD * 9, CQ = COUT / 4, PL = 256 / CQ;
    extern __shared__ float red[];                         // [PL][T + 1][COUT] would be too big: reduce tap by tap
    const int cq = threadIdx.x % CQ, pl = threadIdx.x / CQ;
    float acc[T][4], cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < T; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
#pragma unroll 2
    for (long long m = (long long)blockIdx.x * PL + pl; m < p.M; m += (long long)gridDim.x * PL) {
        uint32_t q1, q2, uw, uh, ud, un;
        p.fd_w.divmod((uint32_t)m, q1, uw);
        p.fd_h.divmod(q1, q2, uh);
        p.fd_d.divmod(q2, un, ud);
        const int w = (int)uw, h = (int)uh, d = (int)ud;
        const int n = (int)un;
        const float4 g = ldg4_stream(p.dy + m * COUT + cq * 4);
        cs[0] += g.x; cs[1] += g.y; cs[2] += g.z; cs[3] += g.w;
#pragma unroll
        for (int kd = 0; kd < KD; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int id = d + kd - (KD == 3 ? 1 : 0), ih = h + kh - 1, iw = w + kw - 1;
                    const bool ok = (unsigned)id < (unsigned)p.D && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
                    const float xv = ok ? __ldg(p.x + ((n * p.D + id) * p.H + ih) * p.W + iw) : 0.f;      // < 2^31 elements
                    const int t = (kd * 3 + kh) * 3 + kw;
                    acc[t][0] = fmaf(g.x, xv, acc[t][0]); acc[t][1] = fmaf(g.y, xv, acc[t][1]);
                    acc[t][2] = fmaf(g.z, xv, acc[t][2]); acc[t][3] = fmaf(g.w, xv, acc[t][3]);
                }
    }
    float* out = p.part + (size_t)blockIdx.x * T * COUT;
#pragma unroll
    for (int t = 0; t <= T; ++t) {                         // fully unrolled: acc[] stays in registers
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) red[pl * COUT + cq * 4 + e] = t < T ? acc[t < T ? t : 0][e] : cs[e];
        __syncthreads();
        if (threadIdx.x < COUT) {
            float v = 0.f;
            for (int k = 0; k < PL; ++k) v += red[k * COUT + threadIdx.x];
            if (t < T) out[t * COUT + threadIdx.x] = v;
            else if (p.part_colsum) p.part_colsum[(size_t)blockIdx.x * COUT + threadIdx.x] = v;
        }
    }
}

// defined in conv.cu
int b200_wgrad_reduce_launch(const float* part, const float* part_colsum, int splits, int K, int NG, int A, int T,
                             float* dw, float* db, int accumulate, cudaStream_t st);

static int c1_blocks(long long M) {
    long long b = (M + 255) / 256;
    const long long cap = (long long)b200_num_sms() * 8;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}
static int c1_wg_blocks() { return b200_num_sms() * 4; }

B200_API int b200_conv_c1_supported(const b200_conv_desc* d) {
    return d && d->c0 == 1 && d->c1 == 0 && d->stride == 1 && d->kh == 3 && d->kw == 3 && d->ph == 1 && d->pw == 1 &&
           ((d->kd == 1 && d->pd == 0) || (d->kd == 3 && d->pd == 1)) && (d->cout == 16 || d->cout == 32);
}

static void fill_c1(C1P& p, const b200_conv_desc* d) {
    memset(&p, 0, sizeof(p));
    p.N = d->n; p.D = d->id; p.H = d->ih; p.W = d->iw; p.Cout = d->cout; p.KD = d->kd;
    p.M = (long long)d->n * d->id * d->ih * d->iw;
    p.fd_w.init(d->iw); p.fd_h.init(d->ih); p.fd_d.init(d->id);
}

B200_API int b200_conv_c1_fwd(const b200_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                              cudaStream_t st) {
    B200_REQUIRE(b200_conv_c1_supported(d), "conv_c1_fwd: needs Cin = 1, 3x3(x3) stride 1 pad 1, Cout in {16, 32}");
    B200_REQUIRE(x && w && y, "conv_c1_fwd: null pointer");
    C1P p;
    fill_c1(p, d);
    p.x = x; p.w = w; p.bias = bias; p.y = y;
    const int grid = c1_blocks(p.M);
    if (d->cout == 16 && d->kd == 1) conv_c1_fwd_kernel<16, 1><<<grid, 256, 0, st>>>(p);
    else if (d->cout == 32 && d->kd == 1) conv_c1_fwd_kernel<32, 1><<<grid, 256, 0, st>>>(p);
    else if (d->cout == 16) conv_c1_fwd_kernel<16, 3><<<grid, 256, 0, st>>>(p);
    else conv_c1_fwd_kernel<32, 3><<<grid, 256, 0, st>>>(p);
    B200_CHECK_LAUNCH("conv_c1_fwd");
    return B200_OK;
}

B200_API long long b200_conv_c1_wgrad_workspace_bytes(const b200_conv_desc* d) {
    if (!d) return -1;
    return (long long)c1_wg_blocks() * (d->kd * 9 + 1) * d->cout * sizeof(float);
}

B200_API int b200_conv_c1_wgrad(const b200_conv_desc* d, const float* x, const float* dy, float* workspace,
                                long long workspace_bytes, float* dw, float* db, int accumulate, cudaStream_t st) {
    B200_REQUIRE(b200_conv_c1_supported(d), "conv_c1_wgrad: unsupported convolution");
    B200_REQUIRE(x && dy && workspace && dw, "conv_c1_wgrad: null pointer");
    B200_REQUIRE((long long)d->n * d->id * d->ih * d->iw < (1ll << 31), "conv_c1_wgrad: more than 2^31 pixels");
    B200_REQUIRE(workspace_bytes >= b200_conv_c1_wgrad_workspace_bytes(d), "conv_c1_wgrad: workspace too small");
    C1P p;
    fill_c1(p, d);
    const int T = d->kd * 9, grid = c1_wg_blocks();
    p.x = x; p.dy = dy; p.part = workspace;
    p.part_colsum = db ? workspace + (size_t)grid * T * d->cout : nullptr;
    const size_t smem = (size_t)(256 / (d->cout / 4)) * d->cout * sizeof(float);      // [PL][COUT]
    if (d->cout == 16 && d->kd == 1) conv_c1_wgrad_kernel<16, 1><<<grid, 256, smem, st>>>(p);
    else if (d->cout == 32 && d->kd == 1) conv_c1_wgrad_kernel<32, 1><<<grid, 256, smem, st>>>(p);
    else if (d->cout == 16) conv_c1_wgrad_kernel<16, 3><<<grid, 256, smem, st>>>(p);
    else conv_c1_wgrad_kernel<32, 3><<<grid, 256, smem, st>>>(p);
    B200_CHECK_LAUNCH("conv_c1_wgrad");
    // partial layout [split][k = tap (Cin = 1)][Cout] == the generic reducer's [splits][K][NG] with A = 1
    return b200_wgrad_reduce_launch(p.part, p.part_colsum, grid, T, d->cout, 1, T, dw, db, accumulate, st);
}
