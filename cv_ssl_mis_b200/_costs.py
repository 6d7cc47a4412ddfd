"""Algorithmic bytes / flops of one C-ABI call, computed from its raw arguments -- used only by the profiling mode of
`_lib.call` (bench.py's roofline tables).  "Algorithmic" = every operand read once and every result written once in
fp32: convolutions 4 (M_in Cin + M_out Cout + T Cin Cout) bytes and 2 M_out T Cin Cout flops, etc.
Returns (bytes, flops) or None when the entry point has no model here."""


def _conv(d, out_scale=1.0):
    m_in = d.n * d.id * d.ih * d.iw
    T, cin = d.kd * d.kh * d.kw, d.c0 + d.c1
    if d.stride == 2:                       # kernel-2 stride-2 convolution: one tap per input pixel
        m_out = m_in // T
        return 4.0 * (m_in * cin + m_out * d.cout + T * cin * d.cout), 2.0 * m_out * T * cin * d.cout
    m_out = m_in
    return 4.0 * (m_in * cin + m_out * d.cout + T * cin * d.cout), 2.0 * m_out * T * cin * d.cout


def _deconv(d):                             # ConvTranspose k2 s2: desc holds the LOW-resolution input
    m_in = d.n * d.id * d.ih * d.iw
    T = d.kd * d.kh * d.kw
    return 4.0 * (m_in * d.c0 + m_in * T * d.cout + T * d.c0 * d.cout), 2.0 * m_in * T * d.c0 * d.cout


def _linear(M, O, I):
    return 4.0 * (M * I + O * I + M * O), 2.0 * M * O * I


def cost_of(name, a):
    try:
        if name.startswith("b200_deconv_k2s2"):
            return _deconv(a[0]._obj)
        if name.startswith("b200_conv") and hasattr(a[0], "_obj"):
            return _conv(a[0]._obj)
        if name == "b200_bn_stats_fwd":
            return 4.0 * a[1] * a[2], 0.0
        if name == "b200_bn_act_fwd":
            return 8.0 * a[3] * a[4], 0.0
        if name == "b200_bn_act_bwd":
            return 20.0 * a[7] * a[8], 0.0
        if name == "b200_linear_fwd":
            return _linear(a[7], a[8], a[2] + a[3])
        if name == "b200_linear_dgrad":
            return _linear(a[7], a[8], a[4] + a[5])
        if name == "b200_linear_wgrad":
            return _linear(a[9], a[10], a[2] + a[3])
        if name == "b200_window_attn_fwd":
            t, c, n = a[3] * a[4] * a[5], a[6], a[8] * a[8]
            return 16.0 * t * c, 4.0 * n * t * c
        if name == "b200_window_attn_bwd":
            t, c, n = a[5] * a[6] * a[7], a[8], a[10] * a[10]
            return 28.0 * t * c, 10.0 * n * t * c
        if name == "b200_mha_fwd":
            B, N, heads, hd = a[3], a[4], a[5], a[6]
            return 16.0 * B * N * heads * hd + 4.0 * B * heads * N * N, 4.0 * N * B * N * heads * hd
        if name == "b200_mha_bwd":
            B, N, heads, hd = a[6], a[7], a[8], a[9]
            return 28.0 * B * N * heads * hd + 4.0 * B * heads * N * N, 8.0 * N * B * N * heads * hd
        if name == "b200_layernorm_fwd":
            return 8.0 * a[5] * a[6], 0.0
        if name == "b200_layernorm_bwd":
            return 12.0 * a[8] * a[9], 0.0
        if name == "b200_gelu_fwd":
            return 8.0 * a[2], 0.0
        if name == "b200_gelu_bwd":
            return 12.0 * a[3], 0.0
        if name == "b200_add_droppath":
            return 12.0 * a[3] * a[4], 0.0
        if name in ("b200_add_lrelu_fwd", "b200_lrelu_bwd", "b200_add"):
            return 12.0 * a[3], 0.0
        if name == "b200_maxpool3d_fwd":
            return 4.0 * a[2] * a[3] * a[4] * a[5] * a[6] * 1.125, 0.0
        if name == "b200_maxpool3d_bwd":
            return 4.0 * a[3] * a[4] * a[5] * a[6] * a[7] * (2.125 + (1.0 if a[8] else 0.0)), 0.0
        if name == "b200_upsample3d2x_fwd":
            return 4.0 * a[2] * a[3] * a[4] * a[5] * a[6] * 9.0, 0.0
        if name == "b200_upsample3d2x_bwd":
            return 4.0 * a[2] * a[3] * a[4] * a[5] * a[6] * 9.0, 0.0
        if name == "b200_s2d_gather3d":
            return 8.0 * a[2] * a[3] * a[4] * a[5] * a[6], 0.0
        if name == "b200_d2s_scatter3d":
            return (12.0 if a[8] else 8.0) * a[3] * a[4] * a[5] * a[6] * 8 * a[7], 0.0
        if name == "b200_maxpool2_fwd":
            return 4.0 * a[2] * a[3] * a[4] * a[5] * 1.25, 0.0
        if name == "b200_maxpool2_bwd":
            return 4.0 * a[3] * a[4] * a[5] * a[6] * (2.25 + (1.0 if a[7] else 0.0)), 0.0
        if name == "b200_upsample2x_fwd":
            return 4.0 * a[2] * a[3] * a[4] * a[5] * 5.0, 0.0
        if name == "b200_upsample2x_bwd":
            return 4.0 * a[2] * a[3] * a[4] * a[5] * 5.0, 0.0
        if name == "b200_colsum":
            return 4.0 * a[1] * a[2], 0.0
        if name == "b200_ssl_loss_fwd" or name == "b200_ssl_loss_bwd":
            lab, B, Lb, C, S = (8 if a[3] else 1), a[5], a[6], a[7], a[8]
            by = 4.0 * C * S * (B + (B - Lb)) + Lb * S * lab
            return by + (4.0 * C * S * B if name.endswith("bwd") else 0.0), 0.0
        if name in ("b200_ct_loss_fwd", "b200_cps_loss_fwd", "b200_ct_loss_bwd", "b200_cps_loss_bwd"):
            B, Lb, C, S = a[6], a[7], a[8], a[9]
            return 4.0 * C * S * 2 * B + Lb * S + (4.0 * C * S * B if name.endswith("bwd") else 0.0), 0.0
        if name == "b200_mc_softmax_accumulate":
            R, U, C, S = a[2], a[3], a[4], a[5]
            return 4.0 * C * S * U * (R + 2), 0.0
        if name == "b200_sgd_ema_step":
            return (28.0 if a[3] else 20.0) * a[4], 0.0
        if name == "b200_ema_update":
            return 12.0 * a[2], 0.0
        if name == "b200_noise_add":
            return 8.0 * a[2], 0.0
        if name in ("b200_nchw_to_nhwc", "b200_nhwc_to_nchw"):
            return 8.0 * a[2] * a[3] * a[4], 0.0
        if name == "b200_patch_merge_gather":
            return 8.0 * a[2] * a[3] * a[4] * a[5], 0.0
        if name == "b200_pixel_shuffle":                  # C is the channel count AFTER the expand: B H W p^2 C elements
            return 8.0 * a[2] * a[3] * a[4] * a[5] * a[6] * a[6], 0.0
    except (AttributeError, IndexError, TypeError):
        return None
    return None
