/* b200ssl -- C ABI of the B200 (sm_100a) kernels behind the semi-supervised training step.
 *
 * The reference (ziyangwang007/CV-SSL-MIS) has no FFI layer: its hot path is Python calling ATen/cuDNN.
 * Each entry point below therefore cites the reference Python call it replaces (paths relative to the
 * reference root). Conventions:
 *   - plain pointers and sizes, no framework types; all tensor pointers are DEVICE pointers (fp32 unless
 *     stated), activations are channels-last: [N][D][H][W][C] with C contiguous (a 2D image has D = 1);
 *   - every call is asynchronous on `stream`, allocates nothing, and keeps no global state; scratch memory
 *     comes from the caller (query the size with the matching *_workspace_bytes function);
 *   - return value: 0 on success, negative on error (B200_ERR_*), message via b200_last_error();
 *   - `exact` != 0 selects 3xTF32 (fp32-equivalent) tensor-core math, 0 selects TF32 inputs with fp32
 *     accumulation (what cuDNN runs for the reference by default, torch.backends.cudnn.allow_tf32 = True).
 */
#ifndef B200SSL_H
#define B200SSL_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define B200_ABI_VERSION 1

enum { B200_LABEL_U8 = 0, B200_LABEL_I64 = 1 };

/* weight packing modes (b200_conv_pack_weights); T = taps = kd*kh*kw */
enum {
    B200_PACK_CONV_FWD = 0,       /* w[O][I][T] -> rows (tap, i), cols o             : b200_conv_fwd          */
    B200_PACK_CONV_DGRAD = 1,     /* w[O][I][T] -> rows (flipped tap, o), cols i     : b200_conv_dgrad        */
    B200_PACK_CONV_DGRAD_D2S = 2, /* w[O][I][T] -> rows o, cols (tap, i)             : b200_conv_k2s2_dgrad   */
    B200_PACK_DECONV_FWD = 3,     /* w[I][O][T] -> rows i, cols (tap, o)             : b200_deconv_k2s2_fwd   */
    B200_PACK_DECONV_DGRAD = 4    /* w[I][O][T] -> rows (tap, o), cols i             : b200_deconv_k2s2_dgrad */
};

typedef struct b200_conv_desc {
    int n, id, ih, iw;   /* input batch and spatial dims (id = 1 for 2D)                      */
    int c0, c1;          /* input channels; c1 > 0: the input is the channel concat [src0|src1] */
    int cout;            /* output channels                                                   */
    int kd, kh, kw;      /* kernel (kd = 1 for 2D)                                            */
    int stride;          /* 1 or 2, all dims                                                  */
    int pd, ph, pw;      /* zero padding                                                      */
} b200_conv_desc;

const char* b200_last_error(void);
int b200_abi_version(void);
/* compute capability major*10+minor of the current device, or <0 */
int b200_device_sm(void);

/* ------------------------------------------------------------------ convolutions
 * nn.Conv2d / nn.Conv3d / nn.ConvTranspose3d: code/networks/unet.py:37,41,73,138 ; code/networks/vnet.py:16,73,100,175 */
long long b200_conv_packed_floats(int mode, int O, int I, int T);
int b200_conv_pack_weights(const float* w, float* out, int mode, int O, int I, int T, cudaStream_t stream);
/* dst[n][od][oh][ow][cout] (out_nchw = 0) or dst[n][cout][spatial] (out_nchw = 1) */
int b200_conv_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wp_fwd, const float* bias,
                  float* dst, int out_nchw, int exact, cudaStream_t stream);
/* stride-1 'same' convolution; the input gradient may be split over the two concat sources [dx0|dx1] */
int b200_conv_dgrad(const b200_conv_desc* d, const float* dy, const float* wp_dgrad, float* dx0, float* dx1,
                    int accumulate, int exact, cudaStream_t stream);
/* kernel 2 / stride 2 / pad 0 convolution (code/networks/vnet.py:73) */
int b200_conv_k2s2_dgrad(const b200_conv_desc* d, const float* dy, const float* wp_d2s, float* dx, int accumulate,
                         int exact, cudaStream_t stream);
long long b200_conv_wgrad_workspace_bytes(const b200_conv_desc* d);
/* dw in the framework layout [cout][c0+c1][kd][kh][kw]; db[cout] may be NULL */
int b200_conv_wgrad(const b200_conv_desc* d, const float* src0, const float* src1, const float* dy, float* workspace,
                    long long workspace_bytes, float* dw, float* db, int accumulate, int exact, cudaStream_t stream);
/* ConvTranspose kernel 2 / stride 2 (code/networks/vnet.py:100); desc dims are the LOW-resolution input,
 * c0 = in channels, cout = out channels, weight layout [c0][cout][kd][kh][kw] */
int b200_deconv_k2s2_fwd(const b200_conv_desc* d, const float* x, const float* wp, const float* bias, float* y,
                         int exact, cudaStream_t stream);
int b200_deconv_k2s2_dgrad(const b200_conv_desc* d, const float* dy, const float* wp_dgrad, float* dx, int accumulate,
                           int exact, cudaStream_t stream);
long long b200_deconv_k2s2_wgrad_workspace_bytes(const b200_conv_desc* d);
int b200_deconv_k2s2_wgrad(const b200_conv_desc* d, const float* x, const float* dy, float* workspace,
                           long long workspace_bytes, float* dw, int accumulate, int exact, cudaStream_t stream);

/* Tile kernels for 3x3 (2D) / 3x3x3 (3D) stride-1 pad-1 convolutions -- the production path of the same
 * nn.Conv2d/nn.Conv3d calls (TF32 only): a spatial halo tile is staged once in shared memory and all taps read
 * it through ldmatrix.  Weights use their own packing ([16-channel chunk][tap][cout][16], TF32-rounded):
 * dgrad = 0 for b200_conv_tile_fwd, 1 for b200_conv_tile_dgrad.  b200_conv_tile_supported tells whether a
 * descriptor qualifies (otherwise use the generic b200_conv_* entry points above). */
int b200_conv_tile_supported(const b200_conv_desc* d, int for_wgrad);
long long b200_conv_tile_packed_floats(int dgrad, int O, int I, int T);
int b200_conv_tile_pack_weights(const float* w, float* out, int dgrad, int O, int I, int T, cudaStream_t stream);
int b200_conv_tile_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wt, const float* bias,
                       float* dst, int out_nchw, cudaStream_t stream);
int b200_conv_tile_dgrad(const b200_conv_desc* d, const float* dy, const float* wt_dgrad, float* dx0, float* dx1,
                         int accumulate, cudaStream_t stream);
long long b200_conv_tile_wgrad_workspace_bytes(const b200_conv_desc* d);
int b200_conv_tile_wgrad(const b200_conv_desc* d, const float* src0, const float* src1, const float* dy, float* workspace,
                         long long workspace_bytes, float* dw, float* db, int accumulate, cudaStream_t stream);

/* tcgen05 (5th-gen tensor core) variant of the 2D 3x3 stride-1 pad-1 forward / data-gradient: tcgen05.mma
 * kind::tf32 over a shared-memory halo tile, accumulators in TMEM (csrc/conv_umma.cu).  Own weight packing
 * ([16-channel chunk][tap][4-channel group][cout][4], TF32-rounded). */
int b200_conv_umma_supported(const b200_conv_desc* d, int for_dgrad);
long long b200_conv_umma_packed_floats(int dgrad, int O, int I, int T);
int b200_conv_umma_pack_weights(const float* w, float* out, int dgrad, int O, int I, int T, cudaStream_t stream);
/* persistent and warp-specialised: cp.async producer warps + cp.async.bulk weights -> multi-stage mbarrier ring -> one
 * MMA-issuing lane -> double-buffered TMEM accumulators -> epilogue warps */
int b200_conv_umma2_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wt, const float* bias,
                        float* dst, int out_nchw, cudaStream_t stream);
int b200_conv_umma2_dgrad(const b200_conv_desc* d, const float* dy, const float* wt_dgrad, float* dx0, float* dx1,
                          int accumulate, cudaStream_t stream);

/* Row-ring tcgen05 weight gradient of the 2D 3x3 stride-1 pad-1 convolutions (backward of code/networks/unet.py:37,41;
 * csrc/conv_row_wgrad.cu): image rows staged by TMA, the nine taps expressed as descriptor shifts, TF32 products with fp32
 * accumulation in TMEM, row-range partials reduced in fixed order.  Channel counts: c0, c1, cout multiples of 32
 * (cout <= 128 or a multiple of 128), or the 16-channel family (c0 = 16, c1 in {0, 16}, cout = 16, even width).
 * dw in the framework layout [cout][c0+c1][3][3].  The bias gradient is not summed here (b200_colsum); db_zero, if not
 * NULL, is the bias gradient of a convolution that feeds a TRAIN-mode BatchNorm: it is identically zero there (the
 * BatchNorm backward removes the per-channel mean of its gradient) and is written as such (left alone when accumulating). */
int b200_conv_row_wgrad_supported(const b200_conv_desc* d);
long long b200_conv_row_wgrad_workspace_bytes(const b200_conv_desc* d);
int b200_conv_row_wgrad(const b200_conv_desc* d, const float* src0, const float* src1, const float* dy, float* workspace,
                        long long workspace_bytes, float* dw, float* db_zero, int accumulate, cudaStream_t stream);

/* Row-ring tcgen05 forward / data gradient of the wide-image 2D 3x3 stride-1 pad-1 convolutions (csrc/conv_row.cu;
 * code/networks/unet.py:37,41 at the 256^2 / 128^2 levels): width a multiple of 128, input channels a multiple of 32 or
 * 16 (+16), 16 / 32 / 64 GEMM columns; weights resident in shared memory, image rows staged once by TMA.
 * b200_conv_row_fwd can emit the per-channel (sum, sum of squares) of its output as b200_conv_row_stats_blocks(d) fp64
 * partials [block][2][cout] for b200_bn_finalize -- the statistics pass of the following train-mode BatchNorm
 * (code/networks/unet.py:38,42) without re-reading the output.  Packed weights: [tap][plane][column][k], TF32-rounded. */
/* 0 when the convolution is not served; otherwise 8 + the weight-packing mode (bit 0 data gradient, bit 1 16-channel
 * planes, bit 2 pixel-pair mode for 16-channel tensors) -- the mode is what b200_conv_pack_batch jobs of kind 3 carry */
int b200_conv_row_supported(const b200_conv_desc* d, int dgrad);
long long b200_conv_row_packed_floats(const b200_conv_desc* d, int dgrad);
int b200_conv_row_pack_weights(const b200_conv_desc* d, int dgrad, const float* w, float* out, cudaStream_t stream);
long long b200_conv_row_stats_blocks(const b200_conv_desc* d);
int b200_conv_row_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wpk, const float* bias,
                      float* dst, double* stats_partials, cudaStream_t stream);
int b200_conv_row_dgrad(const b200_conv_desc* d, const float* dy, const float* wpk_dgrad, float* dx0, float* dx1,
                        int accumulate, cudaStream_t stream);

/* nn.MaxPool3d(2) and nn.Upsample(scale_factor=2, mode='trilinear') (align_corners=False) on channels-last volumes
 * [N][D][H][W][C] (code/networks/unet_3D.py:35-48, code/networks/utils.py:264); D, H, W are the INPUT extents of the op.  The
 * pooling backward recomputes the arg-max (first maximum in (kd, kh, kw) scan order, torch's tie rule). */
int b200_maxpool3d_fwd(const float* a, float* out, int N, int D, int H, int W, int C, cudaStream_t stream);
int b200_maxpool3d_bwd(const float* a, const float* dp, float* da, int N, int D, int H, int W, int C, int accumulate,
                       cudaStream_t stream);
int b200_upsample3d2x_fwd(const float* x, float* y, int N, int D, int H, int W, int C, cudaStream_t stream);
int b200_upsample3d2x_bwd(const float* dy, float* dx, int N, int D, int H, int W, int C, int accumulate, cudaStream_t stream);

/* 2x2x2 stride-2 convolutions and transposed convolutions as GEMMs (code/networks/vnet.py:73,100; the UNETR up-blocks):
 * b200_s2d_gather3d writes the space-to-depth view xs[(n,do,ho,wo)][(kd,kh,kw,c)] of x[N][D][H][W][C] (D, H, W even), so that
 * the strided convolution is b200_linear_fwd(xs, W2[cout][8 cin]) (weights packed with B200_PACK_CONV_DGRAD_D2S);
 * b200_d2s_scatter3d writes y[N][2D][2H][2W][C] from ys[(n,d,h,w)][(kd,kh,kw,c)] (+ bias[c]), ys = b200_linear_fwd(x,
 * W2d[8 cout][cin]) (weights packed with B200_PACK_DECONV_DGRAD); accumulate adds to y (data gradient of the strided conv). */
int b200_s2d_gather3d(const float* x, float* xs, int N, int D, int H, int W, int C, cudaStream_t stream);
int b200_d2s_scatter3d(const float* ys, const float* bias, float* y, int N, int D, int H, int W, int C, int accumulate,
                       cudaStream_t stream);

/* Halo-block tcgen05 forward / data gradient of the narrow-image 2D 3x3 stride-1 pad-1 convolutions (csrc/conv_blk.cu;
 * code/networks/unet.py:37,41 at the 64^2 / 32^2 / 16^2 levels): width <= 96, input channels a multiple of 32, GEMM
 * columns a multiple of 32.  Packed weights: the row-kernel layout [tap][plane][column][32] (b200_conv_pack_batch kind 3
 * with mode = data-gradient flag, or b200_conv_blk_pack_weights).  b200_conv_blk_fwd can emit the BatchNorm (sum, sum of
 * squares) of its output as b200_conv_blk_stats_blocks(d) fp64 partials [block][2][cout] for b200_bn_finalize.
 * The same entry points serve the 3D 3x3x3 stride-1 pad-1 convolutions of the VNet (code/networks/vnet.py:28; taps = 27,
 * framework weights [O][I][3][3][3]): the halo block gains a depth axis; 16-channel tensors are staged as 64-byte rows.
 * b200_conv_blk_supported returns 0 or 8 + the weight-pack mode (bit 0 data gradient, bit 1 16-channel planes). */
int b200_conv_blk_supported(const b200_conv_desc* d, int dgrad);
long long b200_conv_blk_stats_blocks(const b200_conv_desc* d);
int b200_conv_blk_pack_weights(const float* w, float* out, int mode, int O, int I, int taps, cudaStream_t stream);
int b200_conv_blk_fwd(const b200_conv_desc* d, const float* src0, const float* src1, const float* wpk, const float* bias,
                      float* dst, double* stats_partials, cudaStream_t stream);
int b200_conv_blk_dgrad(const b200_conv_desc* d, const float* dy, const float* wpk_dgrad, float* dx0, float* dx1,
                        int accumulate, cudaStream_t stream);

/* One-launch weight packing for a whole network.  jobs_dev: DEVICE array of njobs x 8 int64:
 * [src ptr, dst ptr, kind (0 generic / 1 tile / 2 umma / 3 row), mode (generic: B200_PACK_*; tile/umma: dgrad flag; row: b200_conv_row_supported() - 8), O, I, T, total]. */
int b200_conv_pack_batch(const long long* jobs_dev, int njobs, int blocks_per_job, cudaStream_t stream);
/* First layer of the CNNs (Cin = 1, 3x3 / 3x3x3 stride 1 pad 1, Cout in {16, 32}): HBM-bound FFMA kernels working on the
 * framework weight layout directly (code/networks/unet.py:37 with in_chns = 1, code/networks/vnet.py:152). */
int b200_conv_c1_supported(const b200_conv_desc* d);
int b200_conv_c1_fwd(const b200_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t stream);
long long b200_conv_c1_wgrad_workspace_bytes(const b200_conv_desc* d);
int b200_conv_c1_wgrad(const b200_conv_desc* d, const float* x, const float* dy, float* workspace, long long workspace_bytes,
                       float* dw, float* db, int accumulate, cudaStream_t stream);

/* ------------------------------------------------------------------ BatchNorm(train) + activation + dropout
 * nn.BatchNorm2d/3d + nn.LeakyReLU/ReLU + nn.Dropout/Dropout3d: code/networks/unet.py:38-43 ; code/networks/vnet.py:16-25,177
 * state = [mean | invstd | scale | shift] (4*C floats).  drop_mode: 0 none, 1 per element, 2 per (sample, channel).
 * The Philox key is seed + *seed_offset_dev (DEVICE scalar, may be NULL): bump it per step so a captured CUDA
 * graph draws fresh masks/noise at every replay. */
long long b200_bn_workspace_bytes(long long M, int C);
int b200_bn_stats_fwd(const float* y, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                      float* running_mean, float* running_var, float* state, void* workspace, long long workspace_bytes,
                      cudaStream_t stream);
/* second half of b200_bn_stats_fwd: [nblocks][2][C] fp64 partial (sum, sum of squares) -> state + running statistics */
int b200_bn_finalize(const double* partials, int nblocks, long long M, int C, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, float* state, cudaStream_t stream);
int b200_bn_eval_state(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                       const float* running_var, float* state, cudaStream_t stream);
int b200_bn_act_fwd(const float* y, const float* state, float* a, long long M, int C, float slope, float p_drop,
                    int drop_mode, unsigned long long seed, const unsigned long long* seed_offset_dev, unsigned rng_stream,
                    long long spatial, cudaStream_t stream);
int b200_bn_act_bwd(const float* y, const float* da, const float* state, float* dy, float* dgamma, float* dbeta,
                    int accumulate, long long M, int C, float slope, float p_drop, int drop_mode, unsigned long long seed,
                    const unsigned long long* seed_offset_dev, unsigned rng_stream, long long spatial, void* workspace,
                    long long workspace_bytes, cudaStream_t stream);
int b200_dropout_mask(float* mask, long long M, int C, float p_drop, int drop_mode, unsigned long long seed,
                      const unsigned long long* seed_offset_dev, unsigned rng_stream, long long spatial, cudaStream_t stream);

/* ------------------------------------------------------------------ resampling / layout
 * nn.MaxPool2d(2): code/networks/unet.py:56 ; nn.Upsample(bilinear, align_corners=True): code/networks/unet.py:74-75 */
int b200_maxpool2_fwd(const float* a, float* out, int N, int H, int W, int C, cudaStream_t stream);
int b200_maxpool2_bwd(const float* a, const float* dp, float* da, int N, int H, int W, int C, int accumulate,
                      cudaStream_t stream);
int b200_upsample2x_fwd(const float* x, float* y, int N, int H, int W, int C, cudaStream_t stream);
int b200_upsample2x_bwd(const float* dy, float* dx, int N, int H, int W, int C, int accumulate, cudaStream_t stream);
int b200_nchw_to_nhwc(const float* src, float* dst, long long N, int C, long long S, cudaStream_t stream);
int b200_nhwc_to_nchw(const float* src, float* dst, long long N, int C, long long S, cudaStream_t stream);
long long b200_colsum_workspace_bytes(long long M, int C);
int b200_colsum(const float* g, long long M, int C, float* out, int accumulate, float* workspace,
                long long workspace_bytes, cudaStream_t stream);
int b200_add(const float* a, const float* b, float* c, long long n, cudaStream_t stream);

/* ------------------------------------------------------------------ semi-supervised loss
 * CE + DiceLoss + softmax-MSE consistency: code/train_mean_teacher_2D.py:213-229 ; code/utils/losses.py:74-91,165-201
 * logits [B][C][S] (layout_nhwc = 0) or [B][S][C] (1); the first Lb samples are labeled, the remaining B-Lb are
 * compared with teacher_logits [B-Lb][..] (NULL: no consistency term).  w_cons: DEVICE scalar (consistency weight).
 * lossbuf (>= 4 + 2*C floats): [0] ce [1] dice [2] consistency [3] total, then backward coefficients. */
long long b200_ssl_loss_workspace_bytes(int B, long long S);
/* Uncertainty-aware variant (code/train_uncertainty_aware_mean_teacher_3D.py:149-179): pass mc_psum = sum over the T
 * stochastic teacher passes of softmax probabilities ([B-Lb][..], logits' layout; see b200_mc_softmax_accumulate), mc_T = T
 * and mc_thr = DEVICE scalar threshold; the consistency term becomes sum(mask*dist) / (2 sum(mask) + 1e-16) with
 * mask = [entropy(mc_psum / T) < thr].  mc_psum = NULL selects the plain mean over all elements.
 * lossbuf needs 5 + 2*C floats. */
int b200_ssl_loss_fwd(const float* logits, const float* teacher_logits, const void* labels, int label_dtype,
                      int layout_nhwc, int B, int Lb, int C, long long S, const float* w_cons, const float* mc_psum,
                      float mc_T, const float* mc_thr, float* lossbuf, void* workspace, long long workspace_bytes,
                      cudaStream_t stream);
int b200_ssl_loss_bwd(const float* logits, const float* teacher_logits, const void* labels, int label_dtype,
                      int layout_nhwc, int B, int Lb, int C, long long S, const float* w_cons, const float* mc_psum,
                      float mc_T, const float* mc_thr, const float* lossbuf, float grad_scale, float* dlogits,
                      int dlogits_nhwc, cudaStream_t stream);
/* psum[u] (+)= sum_{r<R} softmax(logits[r*U + u]) : accumulates the MC-dropout teacher passes without materialising them */
int b200_mc_softmax_accumulate(const float* logits, float* psum, int R, int U, int C, long long S, int layout_nhwc,
                               int init, cudaStream_t stream);

/* ------------------------------------------------------------------ nn.Linear on tcgen05 (TMA-fed, TF32, TMEM accumulators)
 * Replaces torch.nn.functional.linear and its autograd for the token matrices of the Swin-UNet
 * (swin_transformer_unet_skip_expand_decoder_sys.py: Mlp :19-25, qkv/proj :115-150, PatchMerging.reduction :346,
 * PatchExpand.expand :378, concat_back_dim :771) and any 1x1 convolution on channels-last rows.
 * x = [x0 | x1] is a virtual concat along the feature dim (c1 = 0: single source; else c0 % 32 == 0); w is the
 * module's own row-major [O][c0+c1] weight (no packing); rows M, features multiples of 4, O >= 16.
 *   fwd   : y[M][O]  = x w^T (+ bias)
 *   dgrad : dx0[M][c0] | dx1[M][c1] (+)= dy w
 *   wgrad : dw[O][c0+c1] (+)= dy^T x   (split over row ranges into `workspace`, reduced in fixed order) */
int b200_linear_supported(long long M, int O, int c0, int c1);
int b200_linear_fwd(const float* x0, const float* x1, int c0, int c1, const float* w, const float* bias, float* y,
                    long long M, int O, cudaStream_t stream);
int b200_linear_dgrad(const float* dy, const float* w, float* dx0, float* dx1, int c0, int c1, int accumulate,
                      long long M, int O, cudaStream_t stream);
long long b200_linear_wgrad_workspace_bytes(long long M, int O, int I);
int b200_linear_wgrad(const float* x0, const float* x1, int c0, int c1, const float* dy, float* dw, int accumulate,
                      float* workspace, long long workspace_bytes, long long M, int O, cudaStream_t stream);

/* ------------------------------------------------------------------ UNETR (code/networks/unetr.py:215-230 over MONAI blocks)
 * patch3d_gather: einops "b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)" of PatchEmbeddingBlock("perceptron").
 * mha: SABlock -- qkv [B*N][3C] (q | k | v, head-major inside each, no bias), softmax(q k^T / sqrt(hd)) v; N <= 256,
 *      hd <= 64; `probs` [B*heads][N][N] is written by fwd and read by bwd (workspace: as many floats again).
 * add_lrelu / lrelu_bwd: `out += residual; out = lrelu(out)` of UnetResBlock and its gradient. */
int b200_patch3d_gather(const float* x, float* y, int B, int C, int D, int H, int W, int patch, cudaStream_t stream);
long long b200_mha_probs_floats(int B, int N, int heads);
int b200_mha_fwd(const float* qkv, float* out, float* probs, int B, int N, int heads, int hd, cudaStream_t stream);
int b200_mha_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, float* workspace,
                 long long workspace_bytes, int B, int N, int heads, int hd, cudaStream_t stream);
int b200_add_lrelu_fwd(const float* a, const float* b, float* out, long long n, float slope, cudaStream_t stream);
int b200_lrelu_bwd(const float* out, const float* dout, float* dx, long long n, float slope, cudaStream_t stream);

/* ------------------------------------------------------------------ Swin-UNet token ops
 * code/networks/swin_transformer_unet_skip_expand_decoder_sys.py: nn.LayerNorm (:204,211,...), nn.GELU in Mlp (:19-25),
 * WindowAttention + roll/partition/reverse of SwinTransformerBlock (:115-150,244-288), DropPath residual (:284-286),
 * PatchMerging gather (:336-341), PatchEmbed im2col (:573).  Tokens are rows of [B*H*W][C] in natural order. */
long long b200_layernorm_workspace_bytes(long long M, int C);
int b200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, long long M, int C,
                       float eps, cudaStream_t stream);
int b200_layernorm_bwd(const float* x, const float* stats, const float* gamma, const float* dy, float* dx, float* dgamma,
                       float* dbeta, int accumulate_dx, long long M, int C, float* workspace, long long workspace_bytes,
                       cudaStream_t stream);
int b200_gelu_fwd(const float* x, float* y, long long n, cudaStream_t stream);
int b200_gelu_bwd(const float* x, const float* dy, float* dx, long long n, int accumulate, cudaStream_t stream);
/* qkv [B*H*W][3C] (q | k | v, head-major inside each), head dim 32, window ws (ws*ws <= 49), cyclic shift `shift` */
int b200_window_attn_fwd(const float* qkv, const float* bias_table, float* out, int B, int H, int W, int C, int heads,
                         int ws, int shift, cudaStream_t stream);
long long b200_window_attn_workspace_bytes(int B, int H, int W, int heads, int ws);
int b200_window_attn_bwd(const float* qkv, const float* bias_table, const float* dout, float* dqkv, float* dbias_table,
                         int B, int H, int W, int C, int heads, int ws, int shift, float* workspace,
                         long long workspace_bytes, cudaStream_t stream);
/* out = x + branch * keep[b] / (1 - p), keep ~ Bernoulli(1 - p) per sample (Philox); x may be NULL (treated as 0) */
int b200_add_droppath(const float* x, const float* branch, float* out, int B, long long per_sample, float p_drop,
                      unsigned long long seed, const unsigned long long* seed_offset_dev, unsigned rng_stream,
                      cudaStream_t stream);
/* inverse = 0: y[B][H/2][W/2][4C] gathered from x[B][H][W][C]; inverse = 1: x is the gathered gradient, y the input gradient */
int b200_patch_merge_gather(const float* x, float* y, int B, int H, int W, int C, int inverse, int accumulate,
                            cudaStream_t stream);
/* PatchExpand rearrange 'b h w (p1 p2 c) -> b (h p1) (w p2) c' (…_sys.py:378-379,405-407): x [B][H][W][p*p*C] ->
 * y [B][H*p][W*p][C]; inverse = 1 maps a gradient laid out like y back to x's layout */
int b200_pixel_shuffle(const float* x, float* y, int B, int H, int W, int C, int p, int inverse, cudaStream_t stream);
int b200_patch_embed_gather(const float* x, float* y, int B, int H, int W, int patch, int repeat_channels, cudaStream_t stream);

/* Cross-teaching loss of ONE model (code/train_cross_teaching_between_cnn_transformer_2D.py:229-247):
 *   0.5 (CE + Dice)(logits[:Lb], labels) + w * Dice(softmax(logits[Lb:]), argmax softmax(other_logits[Lb:]))
 * other_logits holds all B samples of the other model in its own layout.  lossbuf (>= 5 + 4C floats):
 * [0] ce [1] dice [2] pseudo-label dice [3] total, then the gradient coefficients read by b200_ct_loss_bwd. */
int b200_ct_loss_fwd(const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc, const void* labels,
                     int label_dtype, int B, int Lb, int C, long long S, const float* w_cons_dev, float* lossbuf,
                     void* workspace, long long workspace_bytes, cudaStream_t stream);
int b200_ct_loss_bwd(const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc, const void* labels,
                     int label_dtype, int B, int Lb, int C, long long S, const float* lossbuf, float grad_scale,
                     float* dlogits, int dlogits_nhwc, cudaStream_t stream);

/* Cross-pseudo-supervision loss of ONE model (code/train_cross_pseudo_supervision_2D.py:187-196): as above with
 *   w * CrossEntropy(logits[Lb:], argmax softmax(other_logits[Lb:]))  as the pseudo-label term ([2] of lossbuf). */
int b200_cps_loss_fwd(const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc, const void* labels,
                      int label_dtype, int B, int Lb, int C, long long S, const float* w_cons_dev, float* lossbuf,
                      void* workspace, long long workspace_bytes, cudaStream_t stream);
int b200_cps_loss_bwd(const float* logits, int layout_nhwc, const float* other_logits, int other_nhwc, const void* labels,
                      int label_dtype, int B, int Lb, int C, long long S, const float* lossbuf, float grad_scale,
                      float* dlogits, int dlogits_nhwc, cudaStream_t stream);

/* ------------------------------------------------------------------ optimizer / EMA / noise
 * optim.SGD + update_ema_variables + input noise: code/train_mean_teacher_2D.py:124-128,189-190,208-210,230-233
 * hparams_dev (DEVICE, 6 floats): [0] lr [1] momentum [2] weight_decay [3] ema_alpha [4] 1-ema_alpha [5] grad_scale */
int b200_sgd_ema_step(float* params, float* grads, float* momentum_buf, float* ema_params, long long n,
                      const float* hparams_dev, int zero_grad, cudaStream_t stream);
/* Stand-alone losses behind the reference's own signatures (code/utils/losses.py:74-113,165-201), forward / backward
 * halves for torch.autograd.Functions (cv_ssl_mis_b200/utils/losses.py).  Tensors [B][C][S] fp32, C <= 8.
 *   dice:  out[0] = loss, out[1..1+C) = class-wise dice, out[9..33) = per-class (I, Z, Y) sums kept for the backward;
 *          x = probabilities, or logits with use_softmax = 1; labels [B][S] uint8 / int64; weight[C] or NULL;
 *   softmax_mse: element-wise (softmax(input) - softmax(target))^2, gradient to the input logits only;
 *   softmax_kl:  F.kl_div(log_softmax(input), softmax(target), reduction='mean') (mean over all B*C*S elements).
 * grad_out of dice / kl is a DEVICE scalar (the upstream gradient). */
long long b200_loss_dropin_workspace_bytes(int B, long long S);
int b200_dice_fwd(const float* x, int use_softmax, const void* labels, int label_dtype, int B, int C, long long S,
                  const float* weight, float* out, void* workspace, long long workspace_bytes, cudaStream_t stream);
int b200_dice_bwd(const float* x, int use_softmax, const void* labels, int label_dtype, int B, int C, long long S,
                  const float* weight, const float* fwd_out, const float* grad_out, float* dx, cudaStream_t stream);
int b200_softmax_mse_fwd(const float* input_logits, const float* target_logits, int B, int C, long long S, float* out,
                         cudaStream_t stream);
int b200_softmax_mse_bwd(const float* input_logits, const float* target_logits, const float* grad_out, int B, int C,
                         long long S, float* d_input, cudaStream_t stream);
int b200_softmax_kl_fwd(const float* input_logits, const float* target_logits, int B, int C, long long S, float* out,
                        void* workspace, long long workspace_bytes, cudaStream_t stream);
int b200_softmax_kl_bwd(const float* input_logits, const float* target_logits, const float* grad_out, int B, int C,
                        long long S, float* d_input, cudaStream_t stream);
/* teacher = alpha * teacher + (1 - alpha) * student with (alpha, 1 - alpha) read from hparams_dev[3], [4]
 * (code/train_mean_teacher_2D.py:124-128 update_ema_variables) */
int b200_ema_update(float* ema_params, const float* params, long long n, const float* hparams_dev, cudaStream_t stream);
int b200_noise_add(const float* x, float* out, long long n, float sigma, float clip, unsigned long long seed,
                   const unsigned long long* seed_offset_dev, unsigned rng_stream, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SSL_H */
