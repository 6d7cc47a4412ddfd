// 2x2 max-pool, bilinear x2 upsampling (align_corners=True) and layout helpers over channels-last fp32.
//
// Replaces nn.MaxPool2d(2) (code/networks/unet.py:56) and nn.Upsample(scale_factor=2, mode='bilinear',
// align_corners=True) (code/networks/unet.py:74-75) of the reference, forward and backward.
#include "common.cuh"
#include "../../include/b200ssl.h"

static inline int ew_grid(long long work) {
    long long blocks = (work + 255) / 256;
    long long cap = (long long)b200_num_sms() * 16;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// ---------------------------------------------------------------- max pool 2x2 stride 2
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float* __restrict__ a, float* __restrict__ out, int N,
                                                           int H, int W, int C) {
    const int CQ = C >> 2, OH = H >> 1, OW = W >> 1;
    const long long total = (long long)N * OH * OW * CQ;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(q % CQ);
        long long r = q / CQ;
        const int ow = (int)(r % OW); r /= OW;
        const int oh = (int)(r % OH);
        const int n = (int)(r / OH);
        const float* base = a + (((long long)n * H + 2 * oh) * W + 2 * ow) * C + cq * 4;
        float4 v = max4(max4(ldg4(base), ldg4(base + C)), max4(ldg4(base + (long long)W * C), ldg4(base + (long long)W * C + C)));
        stg4(out + q * 4, v);
    }
}

// da[window] (+)= dp routed to the first maximum of the window (scan order (0,0),(0,1),(1,0),(1,1), strict >)
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ a, const float* __restrict__ dp,
                                                           float* __restrict__ da, int N, int H, int W, int C,
                                                           int accumulate) {
    const int CQ = C >> 2, OH = H >> 1, OW = W >> 1;
    const long long total = (long long)N * OH * OW * CQ;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(q % CQ);
        long long r = q / CQ;
        const int ow = (int)(r % OW); r /= OW;
        const int oh = (int)(r % OH);
        const int n = (int)(r / OH);
        const long long off[4] = {0, C, (long long)W * C, (long long)W * C + C};
        const long long base = (((long long)n * H + 2 * oh) * W + 2 * ow) * C + cq * 4;
        float v[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 t = ldg4(a + base + off[k]);
            v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w;
        }
        const float4 g4 = ldg4(dp + q * 4);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
        float o[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int arg = 0;
            float best = v[0][c];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (v[k][c] > best) { best = v[k][c]; arg = k; }
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][c] = (k == arg) ? g[c] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 w = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
            float* d = da + base + off[k];
            if (accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(d);
                w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
            }
            stg4(d, w);
        }
    }
}

// ---------------------------------------------------------------- bilinear x2, align_corners=True
// source coordinate exactly as ATen's area_pixel_compute_source_index: scale = (in-1)/(out-1) in fp32
__device__ __forceinline__ void bil_src(int o, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
    const float s = scale * (float)o;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}

// One block per (n, output row, segment of the row): the row's vertical source rows / weights are block-uniform, the
// horizontal ones are computed once per thread (no 64-bit divisions, no per-element index decoding); the four taps of
// neighbouring outputs overlap, so the low-resolution rows are served by L1/L2 and HBM sees x once and y once.
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N,
                                                             int H, int W, int C, float sh, float sw) {
    const int CQ = C >> 2, OH = 2 * H, OW = 2 * W;
    const int rowq = OW * CQ;                                   // float4 elements of one output row
    const int row = blockIdx.x;                                 // n * OH + oh
    const int n = row / OH, oh = row - n * OH;
    int h0, h1;
    float lh0, lh1;
    bil_src(oh, sh, H, h0, h1, lh0, lh1);
    const float* r0 = x + ((long long)n * H + h0) * W * C;
    const float* r1 = x + ((long long)n * H + h1) * W * C;
    float* yo = y + (long long)row * OW * C;
    for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < rowq; q += gridDim.y * blockDim.x) {
        const int ow = q / CQ, cq = q - ow * CQ;
        int w0, w1;
        float lw0, lw1;
        bil_src(ow, sw, W, w0, w1, lw0, lw1);
        const float4 v00 = ldg4(r0 + w0 * C + cq * 4), v01 = ldg4(r0 + w1 * C + cq * 4);
        const float4 v10 = ldg4(r1 + w0 * C + cq * 4), v11 = ldg4(r1 + w1 * C + cq * 4);
        float4 o;
        o.x = lh0 * (lw0 * v00.x + lw1 * v01.x) + lh1 * (lw0 * v10.x + lw1 * v11.x);
        o.y = lh0 * (lw0 * v00.y + lw1 * v01.y) + lh1 * (lw0 * v10.y + lw1 * v11.y);
        o.z = lh0 * (lw0 * v00.z + lw1 * v01.z) + lh1 * (lw0 * v10.z + lw1 * v11.z);
        o.w = lh0 * (lw0 * v00.w + lw1 * v01.w) + lh1 * (lw0 * v10.w + lw1 * v11.w);
        stg4(yo + q * 4, o);
    }
}

// gather form of the transpose (deterministic, no atomics): every low-res pixel sums the <= 6 x 6 high-res pixels whose
// interpolation footprint touches it, in a fixed order.  One block per (n, low-res row, segment): the vertical weights
// of the candidate output rows are block-uniform (shared memory), the horizontal ones are computed once per thread --
// not once per (row, column) pair as before (49 source-index evaluations per element).
constexpr int UPS_K = 7;      // candidate outputs per dimension: src = scale * o in (h - 1, h + 1), scale ~ 1/2, +-1 slack

__device__ __forceinline__ void ups_candidates(int i, float scale, int in, int out, int& lo, float (&wt)[UPS_K]) {
    lo = scale > 0.f ? (int)floorf((float)(i - 1) / scale) - 1 : 0;
    int hi = scale > 0.f ? (int)ceilf((float)(i + 1) / scale) + 1 : out - 1;
    lo = max(lo, 0);
    hi = min(hi, out - 1);
#pragma unroll
    for (int k = 0; k < UPS_K; ++k) {
        const int o = lo + k;
        float w = 0.f;
        if (o <= hi) {
            int i0, i1;
            float l0, l1;
            bil_src(o, scale, in, i0, i1, l0, l1);
            if (i0 == i) w += l0;
            if (i1 == i) w += l1;
        }
        wt[k] = w;
    }
}

__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N,
                                                             int H, int W, int C, float sh, float sw, int accumulate) {
    const int CQ = C >> 2, OH = 2 * H, OW = 2 * W;
    const int rowq = W * CQ;
    const int row = blockIdx.x;                                 // n * H + h
    const int n = row / H, h = row - n * H;
    int oh_lo;
    float wh[UPS_K];
    ups_candidates(h, sh, H, OH, oh_lo, wh);
    const float* b = dy + (long long)n * OH * OW * C;
    float* dxo = dx + (long long)row * W * C;
    for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < rowq; q += gridDim.y * blockDim.x) {
        const int w = q / CQ, cq = q - w * CQ;
        int ow_lo;
        float ww[UPS_K];
        ups_candidates(w, sw, W, OW, ow_lo, ww);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < UPS_K; ++i) {
            if (wh[i] == 0.f) continue;
            const float* br = b + ((long long)(oh_lo + i) * OW) * C + cq * 4;
#pragma unroll
            for (int j = 0; j < UPS_K; ++j) {
                if (ww[j] == 0.f) continue;
                const float4 g = ldg4(br + (ow_lo + j) * C);
                const float f = wh[i] * ww[j];
                acc.x += f * g.x; acc.y += f * g.y; acc.z += f * g.z; acc.w += f * g.w;
            }
        }
        float* d = dxo + q * 4;
        if (accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(d);
            acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
        }
        stg4(d, acc);
    }
}

// ---------------------------------------------------------------- layout helpers
// [N][C][S] <-> [N][S][C]
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           long long N, int C, long long S) {
    const long long total = N * S * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long r = i / C;
        const long long s = r % S, n = r / S;
        dst[i] = src[(n * C + c) * S + s];
    }
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           long long N, int C, long long S) {
    const long long total = N * S * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i % S;
        const long long r = i / S;
        const int c = (int)(r % C);
        const long long n = r / C;
        dst[i] = src[(n * S + s) * C + c];
    }
}

// column sums of a [M][C] matrix (bias gradient of a ConvTranspose): two-stage deterministic
__global__ void __launch_bounds__(256) colsum_part_kernel(const float* __restrict__ g, long long M, int C,
                                                          float* __restrict__ part) {
    __shared__ float sred[256];
    const int tid = threadIdx.x;
    const int c = tid % C, sl = tid / C, SL = 256 / C;
    float s = 0.f;
    if (sl < SL)
        for (long long m = (long long)blockIdx.x * SL + sl; m < M; m += (long long)gridDim.x * SL) s += __ldg(g + m * C + c);
    sred[tid] = s;
    __syncthreads();
    if (tid < C) {
        float v = 0.f;
        for (int k = 0; k < SL; ++k) v += sred[k * C + tid];
        part[(size_t)blockIdx.x * C + tid] = v;
    }
}
// C % 4 == 0: a thread owns one float4 column; CW column-threads x (256 / CW) row slices per block, blockIdx.y = row
// chunk; four independent 16-byte loads in flight per thread, slices reduced through shared memory in fixed order
__global__ void __launch_bounds__(256) colsum_v4_part_kernel(const float* __restrict__ g, long long M, int C4, int CW,
                                                             long long rows_per_chunk, float* __restrict__ part) {
    __shared__ float4 sred[256];
    const int tid = threadIdx.x, cw = tid % CW, sl = tid / CW, SL = 256 / CW;
    const int c4 = blockIdx.x * CW + cw;
    const long long m0 = blockIdx.y * rows_per_chunk, m1 = m0 + rows_per_chunk < M ? m0 + rows_per_chunk : M;
    float4 s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < C4) {
        const float* col = g + (size_t)c4 * 4;
        const size_t ld = (size_t)C4 * 4;
        long long m = m0 + sl;
        for (; m + 3 * SL < m1; m += 4 * SL) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 v = ldg4_stream(col + (size_t)(m + u * SL) * ld);
                s[u].x += v.x; s[u].y += v.y; s[u].z += v.z; s[u].w += v.w;
            }
        }
        for (; m < m1; m += SL) {
            const float4 v = ldg4_stream(col + (size_t)m * ld);
            s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
        }
    }
    sred[tid] = make_float4((s[0].x + s[1].x) + (s[2].x + s[3].x), (s[0].y + s[1].y) + (s[2].y + s[3].y),
                            (s[0].z + s[1].z) + (s[2].z + s[3].z), (s[0].w + s[1].w) + (s[2].w + s[3].w));
    __syncthreads();
    if (tid < CW && c4 < C4) {
        float4 v = sred[tid];
        for (int k = 1; k < SL; ++k) {
            const float4 w = sred[k * CW + tid];
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        stg4(part + (size_t)blockIdx.y * C4 * 4 + (size_t)c4 * 4, v);
    }
}
// one warp per column: lanes stride over the partial blocks, fixed-order shuffle tree (deterministic); the old
// one-thread-per-column loop spent ~35 us in dependent loads when there were ~900 partial blocks
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ out,
                                                           int accumulate) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    double s = 0;
    for (int b = lane; b < nblk; b += 32) s += (double)part[(size_t)b * C + c];
    s = warp_sum_d(s);
    if (lane == 0) out[c] = accumulate ? out[c] + (float)s : (float)s;
}

// c = a + b (VNet additive skips, code/networks/vnet.py:210,214,218,222)
// c may alias a (in-place accumulate of a gradient), so a / c carry no __restrict__
__global__ void __launch_bounds__(256) add_kernel(const float* a, const float* __restrict__ b, float* c, long long total4) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (long long)gridDim.x * blockDim.x) {
        const float4 x = *reinterpret_cast<const float4*>(a + q * 4), y = ldg4(b + q * 4);
        stg4(c + q * 4, make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w));
    }
}

// ================================================================ C ABI
B200_API int b200_maxpool2_fwd(const float* a, float* out, int N, int H, int W, int C, cudaStream_t st) {
    B200_REQUIRE(a && out && N > 0 && C > 0 && (C & 3) == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_fwd: bad arguments");
    maxpool2_fwd_kernel<<<ew_grid((long long)N * (H / 2) * (W / 2) * (C / 4)), 256, 0, st>>>(a, out, N, H, W, C);
    B200_CHECK_LAUNCH("maxpool2_fwd");
    return B200_OK;
}

B200_API int b200_maxpool2_bwd(const float* a, const float* dp, float* da, int N, int H, int W, int C, int accumulate,
                               cudaStream_t st) {
    B200_REQUIRE(a && dp && da && N > 0 && C > 0 && (C & 3) == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_bwd: bad arguments");
    maxpool2_bwd_kernel<<<ew_grid((long long)N * (H / 2) * (W / 2) * (C / 4)), 256, 0, st>>>(a, dp, da, N, H, W, C, accumulate);
    B200_CHECK_LAUNCH("maxpool2_bwd");
    return B200_OK;
}

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

B200_API int b200_upsample2x_fwd(const float* x, float* y, int N, int H, int W, int C, cudaStream_t st) {
    B200_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, "upsample2x_fwd: bad arguments");
    const int rowq = 2 * W * (C / 4);
    upsample2x_fwd_kernel<<<dim3(N * 2 * H, (rowq + 1023) / 1024), 256, 0, st>>>(x, y, N, H, W, C, ac_scale(H, 2 * H), ac_scale(W, 2 * W));
    B200_CHECK_LAUNCH("upsample2x_fwd");
    return B200_OK;
}

B200_API int b200_upsample2x_bwd(const float* dy, float* dx, int N, int H, int W, int C, int accumulate, cudaStream_t st) {
    B200_REQUIRE(dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, "upsample2x_bwd: bad arguments");
    const int rowq = W * (C / 4);
    upsample2x_bwd_kernel<<<dim3(N * H, (rowq + 511) / 512), 256, 0, st>>>(dy, dx, N, H, W, C, ac_scale(H, 2 * H), ac_scale(W, 2 * W),
                                                                           accumulate);
    B200_CHECK_LAUNCH("upsample2x_bwd");
    return B200_OK;
}

B200_API int b200_nchw_to_nhwc(const float* src, float* dst, long long N, int C, long long S, cudaStream_t st) {
    B200_REQUIRE(src && dst && N > 0 && C > 0 && S > 0, "nchw_to_nhwc: bad arguments");
    nchw_to_nhwc_kernel<<<ew_grid(N * C * S), 256, 0, st>>>(src, dst, N, C, S);
    B200_CHECK_LAUNCH("nchw_to_nhwc");
    return B200_OK;
}

B200_API int b200_nhwc_to_nchw(const float* src, float* dst, long long N, int C, long long S, cudaStream_t st) {
    B200_REQUIRE(src && dst && N > 0 && C > 0 && S > 0, "nhwc_to_nchw: bad arguments");
    nhwc_to_nchw_kernel<<<ew_grid(N * C * S), 256, 0, st>>>(src, dst, N, C, S);
    B200_CHECK_LAUNCH("nhwc_to_nchw");
    return B200_OK;
}

static inline int colsum_cw(int C4) {
    int cw = 1;
    while (cw < C4 && cw < 64) cw <<= 1;
    return cw;
}
static inline int colsum_v4_chunks(long long M, int C) {
    const int C4 = C / 4, CW = colsum_cw(C4);
    const long long colblocks = (C4 + CW - 1) / CW;
    long long chunks = ((long long)b200_num_sms() * 6 + colblocks - 1) / colblocks;
    const long long max_chunks = (M + 63) / 64;
    if (chunks > max_chunks) chunks = max_chunks;
    return (int)(chunks < 1 ? 1 : chunks);
}

B200_API long long b200_colsum_workspace_bytes(long long M, int C) {
    if ((C & 3) == 0) return (long long)colsum_v4_chunks(M, C) * C * sizeof(float);
    return (long long)b200_num_sms() * 4 * C * sizeof(float);
}

B200_API int b200_colsum(const float* g, long long M, int C, float* out, int accumulate, float* workspace,
                         long long workspace_bytes, cudaStream_t st) {
    B200_REQUIRE(g && out && workspace && M > 0 && C > 0 && ((C & 3) == 0 || C <= 256), "colsum: bad arguments (C % 4 == 0 or C <= 256)");
    B200_REQUIRE(workspace_bytes >= b200_colsum_workspace_bytes(M, C), "colsum: workspace too small");
    if ((C & 3) == 0) {
        const int C4 = C / 4, CW = colsum_cw(C4), chunks = colsum_v4_chunks(M, C);
        const long long rows_per_chunk = (M + chunks - 1) / chunks;
        colsum_v4_part_kernel<<<dim3((C4 + CW - 1) / CW, chunks), 256, 0, st>>>(g, M, C4, CW, rows_per_chunk, workspace);
        B200_CHECK_LAUNCH("colsum_v4_part");
        colsum_final_kernel<<<(C + 7) / 8, 256, 0, st>>>(workspace, chunks, C, out, accumulate);
        B200_CHECK_LAUNCH("colsum_final");
        return B200_OK;
    }
    const int SL = 256 / C;
    long long want = (M + SL - 1) / SL;
    int grid = (int)(want < (long long)b200_num_sms() * 4 ? want : (long long)b200_num_sms() * 4);
    colsum_part_kernel<<<grid, 256, 0, st>>>(g, M, C, workspace);
    B200_CHECK_LAUNCH("colsum_part");
    colsum_final_kernel<<<(C + 7) / 8, 256, 0, st>>>(workspace, grid, C, out, accumulate);
    B200_CHECK_LAUNCH("colsum_final");
    return B200_OK;
}

// ---------------------------------------------------------------- 3D pooling / trilinear x2 (code/networks/unet_3D.py)
struct Vol3P {
    int D, H, W, C4;             // INPUT dims of the pooling / of the up-sampling, channels / 4
    FastDiv fdC4, fdW, fdH, fdD; // divisors of the COARSE grid (pool output / up-sample input)
};

// nn.MaxPool3d(2): out[n][d][h][w][c] = max over the 2x2x2 window; BWD routes dp to the FIRST maximum of the window in
// scan order (kd, kh, kw) with a strict > comparison (torch's tie rule -- windows of equal zeros are common after ReLU)
template <int BWD>
__global__ void __launch_bounds__(256) maxpool3d_kernel(const float4* __restrict__ a, const float4* __restrict__ dp,
                                                        float4* __restrict__ out, long long total4, int accumulate, const Vol3P p) {
    const long long sW = p.C4, sH = (long long)p.W * p.C4, sD = (long long)p.H * p.W * p.C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        uint32_t r, c4, w, h, d, n;
        p.fdC4.divmod((uint32_t)i, r, c4);
        p.fdW.divmod(r, r, w);
        p.fdH.divmod(r, r, h);
        p.fdD.divmod(r, n, d);
        const long long base = (((long long)n * p.D + 2 * d) * p.H + 2 * h) * sH + (long long)(2 * w) * sW + c4;
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(a + base + (k >> 2) * sD + ((k >> 1) & 1) * sH + (k & 1) * sW);
        if (!BWD) {
            float4 m = v[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) m = max4(m, v[k]);
            out[i] = m;
        } else {
            const float4 g = __ldg(dp + i);
            int ax = 0, ay = 0, az = 0, aw = 0;
            float bx = v[0].x, by = v[0].y, bz = v[0].z, bw = v[0].w;
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                if (v[k].x > bx) { bx = v[k].x; ax = k; }
                if (v[k].y > by) { by = v[k].y; ay = k; }
                if (v[k].z > bz) { bz = v[k].z; az = k; }
                if (v[k].w > bw) { bw = v[k].w; aw = k; }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float4 o = make_float4(k == ax ? g.x : 0.f, k == ay ? g.y : 0.f, k == az ? g.z : 0.f, k == aw ? g.w : 0.f);
                float4* dst = out + base + (k >> 2) * sD + ((k >> 1) & 1) * sH + (k & 1) * sW;
                if (accumulate) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                *dst = o;
            }
        }
    }
}

static int pool3d_launch(int bwd, const float* a, const float* dp, float* out, int N, int D, int H, int W, int C, int accumulate,
                         cudaStream_t st, const char* who) {
    B200_REQUIRE(a && out && (!bwd || dp) && N > 0 && C > 0 && (C & 3) == 0 && D > 0 && H > 0 && W > 0 && ((D | H | W) & 1) == 0,
                 "%s: bad arguments (C %% 4 == 0, even D, H, W)", who);
    Vol3P p;
    p.D = D; p.H = H; p.W = W; p.C4 = C / 4;
    p.fdC4.init(p.C4); p.fdW.init(W / 2); p.fdH.init(H / 2); p.fdD.init(D / 2);
    const long long total4 = (long long)N * (D / 2) * (H / 2) * (W / 2) * p.C4;
    B200_REQUIRE(total4 < (1ll << 32), "%s: tensor too large", who);
    if (!bwd) maxpool3d_kernel<0><<<ew_grid(total4), 256, 0, st>>>(reinterpret_cast<const float4*>(a), nullptr, reinterpret_cast<float4*>(out), total4, 0, p);
    else maxpool3d_kernel<1><<<ew_grid(total4), 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(dp),
                                                              reinterpret_cast<float4*>(out), total4, accumulate, p);
    B200_CHECK_LAUNCH(who);
    return B200_OK;
}

B200_API int b200_maxpool3d_fwd(const float* a, float* out, int N, int D, int H, int W, int C, cudaStream_t st) {
    return pool3d_launch(0, a, nullptr, out, N, D, H, W, C, 0, st, "maxpool3d_fwd");
}

B200_API int b200_maxpool3d_bwd(const float* a, const float* dp, float* da, int N, int D, int H, int W, int C, int accumulate,
                                cudaStream_t st) {
    return pool3d_launch(1, a, dp, da, N, D, H, W, C, accumulate, st, "maxpool3d_bwd");
}

// nn.Upsample(scale_factor=2, mode='trilinear') (align_corners=False): per axis the source coordinate of output o is
// max(0, (o + 0.5) / 2 - 0.5), i.e. output 2i reads (i-1, i) with weights (0.25, 0.75) (output 0: input 0 alone) and output
// 2i+1 reads (i, min(i+1, n-1)) with weights (0.75, 0.25).  The backward is the transposed gather: input i collects from
// outputs 2i-1 (0.25), 2i (0.75, or 1 at i = 0), 2i+1 (0.75, or 1 at i = n-1), 2i+2 (0.25) -- deterministic, no atomics.
__device__ __forceinline__ void tri_src(int o, int n, int& i0, int& i1, float& w1) {
    const int i = o >> 1;
    if (o & 1) { i0 = i; i1 = i + 1 < n ? i + 1 : n - 1; w1 = 0.25f; }
    else if (i == 0) { i0 = 0; i1 = 0; w1 = 0.f; }
    else { i0 = i - 1; i1 = i; w1 = 0.75f; }
}
// weights of the (up to) four outputs 2i-1 .. 2i+2 that read input i
__device__ __forceinline__ void tri_dst(int i, int n, float (&w)[4]) {
    w[0] = i >= 1 ? 0.25f : 0.f;
    w[1] = i == 0 ? 1.f : 0.75f;
    w[2] = i == n - 1 ? 1.f : 0.75f;
    w[3] = i + 1 < n ? 0.25f : 0.f;
}

__global__ void __launch_bounds__(256) upsample3d2x_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long total4,
                                                               const Vol3P p) {
    const long long sW = p.C4, sH = (long long)p.W * p.C4, sD = (long long)p.H * p.W * p.C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        uint32_t r, c4, w, h, d, n;                       // OUTPUT coordinates (fd* hold the output extents here)
        p.fdC4.divmod((uint32_t)i, r, c4);
        p.fdW.divmod(r, r, w);
        p.fdH.divmod(r, r, h);
        p.fdD.divmod(r, n, d);
        int d0, d1, h0, h1, w0, w1;
        float ad, ah, aw;
        tri_src((int)d, p.D, d0, d1, ad);
        tri_src((int)h, p.H, h0, h1, ah);
        tri_src((int)w, p.W, w0, w1, aw);
        const float4* b = x + (long long)n * p.D * sD + c4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float wt = ((k & 4) ? ad : 1.f - ad) * ((k & 2) ? ah : 1.f - ah) * ((k & 1) ? aw : 1.f - aw);
            const float4 v = __ldg(b + ((k & 4) ? d1 : d0) * sD + ((k & 2) ? h1 : h0) * sH + ((k & 1) ? w1 : w0) * sW);
            acc.x += wt * v.x; acc.y += wt * v.y; acc.z += wt * v.z; acc.w += wt * v.w;
        }
        y[i] = acc;
    }
}

__global__ void __launch_bounds__(256) upsample3d2x_bwd_kernel(const float4* __restrict__ dy, float4* __restrict__ dx, long long total4,
                                                               int accumulate, const Vol3P p) {
    const int OD = 2 * p.D, OH = 2 * p.H, OW = 2 * p.W;
    const long long sW = p.C4, sH = (long long)OW * p.C4, sD = (long long)OH * OW * p.C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        uint32_t r, c4, w, h, d, n;                       // INPUT coordinates
        p.fdC4.divmod((uint32_t)i, r, c4);
        p.fdW.divmod(r, r, w);
        p.fdH.divmod(r, r, h);
        p.fdD.divmod(r, n, d);
        float wd[4], wh[4], ww[4];
        tri_dst((int)d, p.D, wd);
        tri_dst((int)h, p.H, wh);
        tri_dst((int)w, p.W, ww);
        const float4* b = dy + (long long)n * OD * sD + c4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int a = 0; a < 4; ++a) {
            const int od = 2 * (int)d - 1 + a;
            if (wd[a] == 0.f) continue;
            for (int e = 0; e < 4; ++e) {
                const int oh = 2 * (int)h - 1 + e;
                if (wh[e] == 0.f) continue;
                const float wde = wd[a] * wh[e];
                const float4* row = b + od * sD + oh * sH;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    if (ww[f] == 0.f) continue;
                    const float wt = wde * ww[f];
                    const float4 v = __ldg(row + (2 * (int)w - 1 + f) * sW);
                    acc.x += wt * v.x; acc.y += wt * v.y; acc.z += wt * v.z; acc.w += wt * v.w;
                }
            }
        }
        if (accumulate) { const float4 old = dx[i]; acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w; }
        dx[i] = acc;
    }
}

B200_API int b200_upsample3d2x_fwd(const float* x, float* y, int N, int D, int H, int W, int C, cudaStream_t st) {
    B200_REQUIRE(x && y && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, "upsample3d2x_fwd: bad arguments");
    Vol3P p;
    p.D = D; p.H = H; p.W = W; p.C4 = C / 4;
    p.fdC4.init(p.C4); p.fdW.init(2 * W); p.fdH.init(2 * H); p.fdD.init(2 * D);
    const long long total4 = (long long)N * 8 * D * H * W * p.C4;
    B200_REQUIRE(total4 < (1ll << 32), "upsample3d2x_fwd: tensor too large");
    upsample3d2x_fwd_kernel<<<ew_grid(total4), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), total4, p);
    B200_CHECK_LAUNCH("upsample3d2x_fwd");
    return B200_OK;
}

B200_API int b200_upsample3d2x_bwd(const float* dy, float* dx, int N, int D, int H, int W, int C, int accumulate, cudaStream_t st) {
    B200_REQUIRE(dy && dx && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, "upsample3d2x_bwd: bad arguments");
    Vol3P p;
    p.D = D; p.H = H; p.W = W; p.C4 = C / 4;
    p.fdC4.init(p.C4); p.fdW.init(W); p.fdH.init(H); p.fdD.init(D);
    const long long total4 = (long long)N * D * H * W * p.C4;
    B200_REQUIRE(total4 < (1ll << 32), "upsample3d2x_bwd: tensor too large");
    upsample3d2x_bwd_kernel<<<ew_grid(total4), 256, 0, st>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), total4,
                                                             accumulate, p);
    B200_CHECK_LAUNCH("upsample3d2x_bwd");
    return B200_OK;
}

// ---------------------------------------------------------------- 2x2x2 stride-2 (transposed) convolutions as GEMMs
// A kernel-2 stride-2 convolution touches every input voxel exactly once, so it is the GEMM
//     y[(n,do,ho,wo)][co] = sum over (kd,kh,kw,ci) of  xs[(n,do,ho,wo)][(kd,kh,kw,ci)] * W2[co][(kd,kh,kw,ci)]
// over the space-to-depth view xs of x (code/networks/vnet.py:73), and its transpose (vnet.py:100, the UNETR up-blocks)
// is ys[(n,d,h,w)][(kd,kh,kw,co)] = x W2d^T scattered depth-to-space.  These two kernels are the views (pure 16-byte
// copies; the (kw, c) pairs are contiguous on both sides); the products run on gemm_umma_kernel (b200_linear_fwd).
struct S2dP {
    int D, H, W, C4;             // INPUT dims of the strided conv / of the transposed conv, channels / 4 of the moved tensor
    FastDiv fdC4, fdW2, fdH2, fdD2;
};

// gather: DIR = 0: xs[m][(tap, c)] = x[fine voxel]     scatter: DIR = 1: y[fine voxel] = ys[m][(tap, c)] + bias[c]
template <int DIR>
__global__ void __launch_bounds__(256) s2d_kernel(const float4* __restrict__ src, float4* __restrict__ dst, const float* __restrict__ bias,
                                                  long long total4, int accumulate, const S2dP p) {
    const int Wc = DIR == 0 ? p.W / 2 : p.W, Hc = DIR == 0 ? p.H / 2 : p.H, Dc = DIR == 0 ? p.D / 2 : p.D;   // coarse grid
    const int Wf = 2 * Wc, Hf = 2 * Hc, Df = 2 * Dc;                                                          // fine grid
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        // coarse-major index: i = ((m * 8 + tap) * C4 + c4)
        uint32_t r, c4, tap, wo, ho, dz, n;
        p.fdC4.divmod((uint32_t)i, r, c4);
        tap = r & 7; r >>= 3;
        p.fdW2.divmod(r, r, wo);
        p.fdH2.divmod(r, r, ho);
        p.fdD2.divmod(r, n, dz);
        const int kd = tap >> 2, kh = (tap >> 1) & 1, kw = tap & 1;
        const long long fine = ((((long long)n * Df + 2 * dz + kd) * Hf + 2 * ho + kh) * Wf + 2 * wo + kw) * p.C4 + c4;
        if (DIR == 0) {
            dst[i] = __ldg(src + fine);
        } else {
            float4 v = __ldg(src + i);
            if (bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (accumulate) { const float4 old = dst[fine]; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
            dst[fine] = v;
        }
    }
    (void)Wc; (void)Hc; (void)Dc;
}

static int s2d_launch(int dir, const float* src, float* dst, const float* bias, int N, int D, int H, int W, int C, int accumulate,
                      cudaStream_t st, const char* who) {
    B200_REQUIRE(src && dst && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, "%s: bad arguments (C %% 4 == 0)", who);
    B200_REQUIRE(dir == 1 || ((D | H | W) & 1) == 0, "%s: D, H, W must be even", who);
    S2dP p;
    p.D = D; p.H = H; p.W = W; p.C4 = C / 4;
    const int Wc = dir == 0 ? W / 2 : W, Hc = dir == 0 ? H / 2 : H, Dc = dir == 0 ? D / 2 : D;
    p.fdC4.init(p.C4); p.fdW2.init(Wc); p.fdH2.init(Hc); p.fdD2.init(Dc);
    const long long total4 = (long long)N * Dc * Hc * Wc * 8 * p.C4;
    B200_REQUIRE(total4 < (1ll << 32), "%s: tensor too large", who);
    if (dir == 0) s2d_kernel<0><<<ew_grid(total4), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), nullptr, total4, 0, p);
    else s2d_kernel<1><<<ew_grid(total4), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), bias, total4, accumulate, p);
    B200_CHECK_LAUNCH(who);
    return B200_OK;
}

B200_API int b200_s2d_gather3d(const float* x, float* xs, int N, int D, int H, int W, int C, cudaStream_t st) {
    return s2d_launch(0, x, xs, nullptr, N, D, H, W, C, 0, st, "s2d_gather3d");
}

B200_API int b200_d2s_scatter3d(const float* ys, const float* bias, float* y, int N, int D, int H, int W, int C, int accumulate,
                                cudaStream_t st) {
    return s2d_launch(1, ys, y, bias, N, D, H, W, C, accumulate, st, "d2s_scatter3d");
}

B200_API int b200_add(const float* a, const float* b, float* c, long long n, cudaStream_t st) {
    B200_REQUIRE(a && b && c && n > 0 && (n & 3) == 0, "add: bad arguments (n multiple of 4)");
    add_kernel<<<ew_grid(n / 4), 256, 0, st>>>(a, b, c, n / 4);
    B200_CHECK_LAUNCH("add");
    return B200_OK;
}
