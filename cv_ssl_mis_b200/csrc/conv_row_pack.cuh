// Weight packing of the row-ring convolution kernels (conv_row.cu), shared by the stand-alone packer and the
// one-launch batch packer (conv_small.cu).  mode: bit 0 = data gradient (reduce over cout, taps flipped),
// bit 1 = 16-channel planes (64-byte rows), bit 2 = pixel-pair mode.
//
// Plain modes:  out[tap][plane][col][k in cpp]        forward: col = cout, k = cin;  data gradient: col = cin, k = cout.
// Pair mode (16-channel tensors viewed as rows of 2 pixels x 16 channels = 128 bytes): the GEMM runs on pixel PAIRS,
//   out[kh][j][plane][col = (dest, pa, c_out)][k = (pb, c_red)]  with the horizontal tap kw = 2 (s - 1) + pb - pa + 1
//   (pair shift s - 1, pixel parities pa of the output and pb of the input) and zeros where that kw is not a tap.
//   j = 0 is the centre shift s = 1.  The shifts s = 0 and s = 2 only meet the odd (pb = 1) and the even (pb = 0) input
//   pixel, so they SHARE block j = 1: its pb = 0 half holds s = 2, its pb = 1 half s = 0 (the kernel skips the other k-steps).
// Values are rounded to TF32 (round to nearest; the tensor core truncates its operands).
#pragma once
#include "conv_common.cuh"

__host__ __device__ __forceinline__ long long row_pack_total(int mode, int O, int I) {
    return (mode & 4) ? 6ll * O * I * 4 : 9ll * O * I;
}

// T: taps of the filter (9, or 27 for the 3D halo-block kernel; the pair mode is 2D only)
__device__ __forceinline__ float row_pack_elem(const float* __restrict__ w, int mode, int O, int I, int T, int idx) {
    const int dgrad = mode & 1;
    const int rows = dgrad ? O : I, cols = dgrad ? I : O;
    float v;
    if (!(mode & 4)) {
        const int cpp = (mode & 2) ? 16 : 32, np = rows / cpp;
        const int k = idx % cpp;
        int r = idx / cpp;
        const int col = r % cols; r /= cols;
        const int pl = r % np, tap = r / np;
        const int row = pl * cpp + k;
        v = dgrad ? w[((size_t)row * I + col) * T + (T - 1 - tap)] : w[((size_t)col * I + row) * T + tap];
    } else {
        const int np = rows / 16;
        const int k = idx & 31;
        int r = idx >> 5;
        const int col = r % (2 * cols); r /= (2 * cols);
        const int pl = r % np, t = r / np;
        const int kh = t >> 1, j = t & 1;
        const int pb = k >> 4, arow = pl * 16 + (k & 15);
        const int s = j == 0 ? 1 : (pb ? 0 : 2);
        const int dst = col >> 5, pa = (col >> 4) & 1, ncol = dst * 16 + (col & 15);
        const int kw = 2 * (s - 1) + pb - pa + 1;
        if (kw < 0 || kw > 2) return 0.f;
        v = dgrad ? w[((size_t)arow * I + ncol) * 9 + (2 - kh) * 3 + (2 - kw)] : w[((size_t)ncol * I + arow) * 9 + kh * 3 + kw];
    }
    return __uint_as_float(f2tf32(v));
}
