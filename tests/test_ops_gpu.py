"""Every C-ABI kernel against a plain PyTorch fp32 reference of the same op (tests/fake_ops.py, CPU), on seeded
inputs.  `exact` (3xTF32) must match to fp32 round-off; the default TF32 mode to TF32 tolerance."""
import numpy as np
import pytest
import torch

from cv_ssl_mis_b200 import ops, _lib
from oracle import philox
from tests import fake_ops as ref

pytestmark = pytest.mark.gpu
DEV = "cuda"

TOL = {True: dict(rtol=1e-3, atol=1e-4), False: dict(rtol=2e-2, atol=5e-3)}


def rnd(g, *shape, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).float()


def cu(t):
    return None if t is None else t.to(DEV).contiguous()


def packed(w_cpu, mode, O, I, T):
    out = torch.empty(ops.conv_packed_floats(mode, O, I, T), device=DEV)
    ops.conv_pack_weights(cu(w_cpu), out, mode, O, I, T)
    out_ref = torch.empty(ref.conv_packed_floats(mode, O, I, T))
    ref.conv_pack_weights(w_cpu, out_ref, mode, O, I, T)
    assert torch.equal(out.cpu(), out_ref), "weight packing differs"
    return out, out_ref


CONV_CASES = [
    # n, d, h, w, c0, c1, cout, k, stride, dims
    (2, 1, 32, 32, 1, 0, 16, 3, 1, 2),
    (2, 1, 20, 24, 16, 0, 16, 3, 1, 2),
    (1, 1, 16, 16, 16, 16, 32, 3, 1, 2),
    (2, 1, 8, 8, 64, 0, 128, 3, 1, 2),
    (1, 1, 8, 8, 128, 128, 160, 3, 1, 2),
    (2, 1, 16, 16, 32, 0, 16, 1, 1, 2),
    (2, 1, 16, 16, 16, 0, 4, 3, 1, 2),
    (1, 8, 12, 16, 1, 0, 16, 3, 1, 3),
    (1, 6, 6, 10, 16, 0, 32, 3, 1, 3),
    (2, 8, 8, 8, 16, 0, 32, 2, 2, 3),
    (1, 4, 4, 4, 16, 0, 2, 1, 1, 3),
    (2, 1, 16, 16, 16, 0, 32, 2, 2, 2),
    (2, 1, 14, 14, 96, 0, 4, 1, 1, 2),       # Swin-UNet head (1x1, 96 -> 4): FFMA data gradient
]


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(case, exact):
    n, d, h, w, c0, c1, cout, k, stride, dims = case
    g = torch.Generator().manual_seed(hash(case) % 10000)
    desc = ops.conv_desc(n, d, h, w, c0, c1, cout, k, stride, None, dims)
    od, oh, ow = ops.desc_out_dims(desc)
    T = k ** dims
    cin = c0 + c1
    M_in, M_out = n * d * h * w, n * od * oh * ow
    wgt = rnd(g, cout, cin, *([k] * dims), scale=(cin * T) ** -0.5)
    bias = rnd(g, cout)
    x0, x1 = rnd(g, M_in, c0), (rnd(g, M_in, c1) if c1 else None)
    dy = rnd(g, M_out, cout)
    tol = TOL[exact]
    # forward
    wp, wp_ref = packed(wgt, _lib.PACK_CONV_FWD, cout, cin, T)
    y = torch.empty(M_out, cout, device=DEV)
    ops.conv_fwd(desc, cu(x0), cu(x1), wp, cu(bias), y, False, exact)
    y_ref = torch.empty(M_out, cout)
    ref.conv_fwd(desc, x0, x1, wp_ref, bias, y_ref)
    torch.testing.assert_close(y.cpu(), y_ref, **tol)
    y2 = torch.empty(n, cout, od * oh * ow, device=DEV)
    ops.conv_fwd(desc, cu(x0), cu(x1), wp, cu(bias), y2, True, exact)
    torch.testing.assert_close(y2.cpu(), y_ref.view(n, -1, cout).permute(0, 2, 1), **tol)
    # weight gradient
    ws = torch.empty(ops.conv_wgrad_workspace_bytes(desc) // 4 + 4, device=DEV)
    dw, db = torch.empty_like(cu(wgt)), torch.empty(cout, device=DEV)
    ops.conv_wgrad(desc, cu(x0), cu(x1), cu(dy), ws, dw, db, False, exact)
    dw_ref, db_ref = torch.empty_like(wgt), torch.empty(cout)
    ref.conv_wgrad(desc, x0, x1, dy, None, dw_ref, db_ref)
    scale = float(dw_ref.abs().max())
    torch.testing.assert_close(dw.cpu(), dw_ref, rtol=tol["rtol"], atol=tol["atol"] * max(scale, 1.0))
    torch.testing.assert_close(db.cpu(), db_ref, rtol=1e-4, atol=1e-4 * max(float(db_ref.abs().max()), 1.0))
    # data gradient
    if stride == 1:
        wpd, wpd_ref = packed(wgt, _lib.PACK_CONV_DGRAD, cout, cin, T)
        dx0 = torch.full((M_in, c0), 7.0, device=DEV)
        dx1 = torch.full((M_in, c1), 7.0, device=DEV) if c1 else None
        ops.conv_dgrad(desc, cu(dy), wpd, dx0, dx1, False, exact)
        r0, r1 = torch.empty(M_in, c0), (torch.empty(M_in, c1) if c1 else None)
        ref.conv_dgrad(desc, dy, wpd_ref, r0, r1)
        torch.testing.assert_close(dx0.cpu(), r0, **tol)
        if c1:
            torch.testing.assert_close(dx1.cpu(), r1, **tol)
        ops.conv_dgrad(desc, cu(dy), wpd, dx0, dx1, True, exact)        # accumulate
        torch.testing.assert_close(dx0.cpu(), 2 * r0, rtol=tol["rtol"], atol=2 * tol["atol"])
    else:
        wpd, wpd_ref = packed(wgt, _lib.PACK_CONV_DGRAD_D2S, cout, cin, T)
        dx = torch.zeros(M_in, c0, device=DEV)
        ops.conv_k2s2_dgrad(desc, cu(dy), wpd, dx, False, exact)
        r = torch.empty(M_in, c0)
        ref.conv_k2s2_dgrad(desc, dy, wpd_ref, r)
        torch.testing.assert_close(dx.cpu(), r, **tol)


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("case", [(2, 4, 6, 8, 32, 16, 3), (1, 3, 3, 3, 256, 128, 3), (2, 1, 8, 8, 32, 16, 2)])
def test_deconv_k2s2(case, exact):
    n, d, h, w, cin, cout, dims = case
    g = torch.Generator().manual_seed(sum(case))
    desc = ops.conv_desc(n, d, h, w, cin, 0, cout, 2, 2, 0, dims)
    T = 2 ** dims
    M_in, M_out = n * d * h * w, n * d * h * w * T
    wgt = rnd(g, cin, cout, *([2] * dims), scale=cin ** -0.5)
    bias, x, dy = rnd(g, cout), rnd(g, M_in, cin), rnd(g, M_out, cout)
    tol = TOL[exact]
    wp, wp_ref = packed(wgt, _lib.PACK_DECONV_FWD, cout, cin, T)
    y, y_ref = torch.empty(M_out, cout, device=DEV), torch.empty(M_out, cout)
    ops.deconv_k2s2_fwd(desc, cu(x), wp, cu(bias), y, exact)
    ref.deconv_k2s2_fwd(desc, x, wp_ref, bias, y_ref)
    torch.testing.assert_close(y.cpu(), y_ref, **tol)
    wpd, wpd_ref = packed(wgt, _lib.PACK_DECONV_DGRAD, cout, cin, T)
    dx, dx_ref = torch.empty(M_in, cin, device=DEV), torch.empty(M_in, cin)
    ops.deconv_k2s2_dgrad(desc, cu(dy), wpd, dx, False, exact)
    ref.deconv_k2s2_dgrad(desc, dy, wpd_ref, dx_ref)
    torch.testing.assert_close(dx.cpu(), dx_ref, **tol)
    ws = torch.empty(ops.deconv_k2s2_wgrad_workspace_bytes(desc) // 4 + 4, device=DEV)
    dw, dw_ref = torch.empty_like(cu(wgt)), torch.empty_like(wgt)
    ops.deconv_k2s2_wgrad(desc, cu(x), cu(dy), ws, dw, False, exact)
    ref.deconv_k2s2_wgrad(desc, x, dy, None, dw_ref)
    torch.testing.assert_close(dw.cpu(), dw_ref, rtol=tol["rtol"], atol=tol["atol"] * max(float(dw_ref.abs().max()), 1.0))
    ws2 = torch.empty(ops.colsum_workspace_bytes(M_out, cout) // 4 + 4, device=DEV)
    db = torch.empty(cout, device=DEV)
    ops.colsum(cu(dy), M_out, cout, db, ws2)
    torch.testing.assert_close(db.cpu(), dy.sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("drop", [(0.0, 0, 1), (0.3, 1, 1), (0.5, 2, 48)])
@pytest.mark.parametrize("MC", [(2 * 24 * 24, 16), (96, 256), (1000, 96)])
@pytest.mark.parametrize("slope", [0.01, 0.0])
def test_bn_act_fwd_bwd(MC, drop, slope):
    M, C = MC
    p, mode, spatial = drop
    if mode == 2:
        spatial = M // 2 if M % 2 == 0 else M
    g = torch.Generator().manual_seed(M + C)
    y = rnd(g, M, C) * 2 + 0.5
    gamma, beta = rnd(g, C).abs() + 0.5, rnd(g, C)
    da = rnd(g, M, C)
    rm, rv = rnd(g, C), rnd(g, C).abs() + 0.1
    seed, off, stream = 4242, 3, 5
    ws = torch.empty(ops.bn_workspace_bytes(M, C) // 4 + 4, device=DEV)
    state, rm_d, rv_d = torch.empty(4 * C, device=DEV), cu(rm.clone()), cu(rv.clone())
    off_d = torch.tensor([off], dtype=torch.int64, device=DEV)
    ops.bn_stats_fwd(cu(y), M, C, cu(gamma), cu(beta), 1e-5, 0.1, rm_d, rv_d, state, ws)
    state_ref, rm_r, rv_r = torch.empty(4 * C), rm.clone(), rv.clone()
    ref.bn_stats_fwd(y, M, C, gamma, beta, 1e-5, 0.1, rm_r, rv_r, state_ref, None)
    torch.testing.assert_close(state.cpu(), state_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rm_d.cpu(), rm_r, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv_d.cpu(), rv_r, rtol=1e-5, atol=1e-6)
    off_c = torch.tensor([off])
    a, a_ref = torch.empty(M, C, device=DEV), torch.empty(M, C)
    ops.bn_act_fwd(cu(y), state, a, M, C, slope, p, mode, seed, off_d, stream, spatial)
    ref.bn_act_fwd(y, state_ref, a_ref, M, C, slope, p, mode, seed, off_c, stream, spatial)
    torch.testing.assert_close(a.cpu(), a_ref, rtol=1e-5, atol=1e-5)
    dyv, dg, db = torch.empty(M, C, device=DEV), torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    ops.bn_act_bwd(cu(y), cu(da), state, dyv, dg, db, M, C, slope, ws, p, mode, seed, off_d, stream, spatial)
    dy_r, dg_r, db_r = torch.empty(M, C), torch.empty(C), torch.empty(C)
    ref.bn_act_bwd(y, da, state_ref, dy_r, dg_r, db_r, M, C, slope, None, p, mode, seed, off_c, stream, spatial)
    torch.testing.assert_close(dyv.cpu(), dy_r, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dg.cpu(), dg_r, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db.cpu(), db_r, rtol=1e-4, atol=1e-3)
    # in-place form used by the engine (dy written over da)
    da_d = cu(da.clone())
    ops.bn_act_bwd(cu(y), da_d, state, da_d, dg, db, M, C, slope, ws, p, mode, seed, off_d, stream, spatial)
    torch.testing.assert_close(da_d.cpu(), dy_r, rtol=1e-4, atol=1e-5)
    # eval-mode coefficients
    ops.bn_eval_state(C, cu(gamma), cu(beta), 1e-5, cu(rm), cu(rv), state)
    ref.bn_eval_state(C, gamma, beta, 1e-5, rm, rv, state_ref)
    torch.testing.assert_close(state.cpu(), state_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mode,spatial", [(1, 1), (2, 64)])
def test_dropout_mask_is_the_philox_stream(mode, spatial):
    M, C, p = 256, 32, 0.3
    mask = torch.empty(M, C, device=DEV)
    off = torch.tensor([11], dtype=torch.int64, device=DEV)
    ops.dropout_mask(mask, M, C, p, mode, 99, off, 7, spatial)
    expect = philox.keep_mask(99 + 11, 7, M, C, p, mode, spatial)
    assert np.array_equal(mask.cpu().numpy(), expect)              # bit-exact integer RNG
    assert 0.6 < float(mask.mean()) < 0.8


def test_pool_upsample_layout():
    g = torch.Generator().manual_seed(5)
    N, H, W, C = 2, 12, 20, 16
    a = rnd(g, N * H * W, C)
    out, out_r = torch.empty(N * H * W // 4, C, device=DEV), torch.empty(N * H * W // 4, C)
    ops.maxpool2_fwd(cu(a), out, N, H, W, C)
    ref.maxpool2_fwd(a, out_r, N, H, W, C)
    assert torch.equal(out.cpu(), out_r)
    dp = rnd(g, N * H * W // 4, C)
    da, da_r = torch.ones(N * H * W, C, device=DEV), torch.ones(N * H * W, C)
    ops.maxpool2_bwd(cu(a), cu(dp), da, N, H, W, C, True)
    ref.maxpool2_bwd(a, dp, da_r, N, H, W, C, True)
    torch.testing.assert_close(da.cpu(), da_r, rtol=0, atol=0)
    up, up_r = torch.empty(N * 4 * H * W, C, device=DEV), torch.empty(N * 4 * H * W, C)
    ops.upsample2x_fwd(cu(a), up, N, H, W, C)
    ref.upsample2x_fwd(a, up_r, N, H, W, C)
    torch.testing.assert_close(up.cpu(), up_r, rtol=1e-5, atol=1e-5)
    dyu = rnd(g, N * 4 * H * W, C)
    dx, dx_r = torch.empty(N * H * W, C, device=DEV), torch.empty(N * H * W, C)
    ops.upsample2x_bwd(cu(dyu), dx, N, H, W, C)
    ref.upsample2x_bwd(dyu, dx_r, N, H, W, C)
    torch.testing.assert_close(dx.cpu(), dx_r, rtol=1e-4, atol=1e-5)
    t = rnd(g, 3, 5, 77)
    o = torch.empty(3, 77, 5, device=DEV)
    ops.nchw_to_nhwc(cu(t), o, 3, 5, 77)
    assert torch.equal(o.cpu(), t.permute(0, 2, 1).contiguous())
    o2 = torch.empty(3, 5, 77, device=DEV)
    ops.nhwc_to_nchw(o, o2, 3, 5, 77)
    assert torch.equal(o2.cpu(), t)
    s = torch.empty(N * H * W, C, device=DEV)
    ops.add(cu(a), cu(a), s)
    assert torch.equal(s.cpu(), a + a)


@pytest.mark.parametrize("name", ["2d", "3d"])
@pytest.mark.parametrize("nhwc", [False, True])
def test_ssl_loss_against_reference_fixture(golden, name, nhwc):
    """Loss values and d(loss)/d(logits) against numbers produced by the reference's DiceLoss/CrossEntropy/softmax_mse."""
    gold = golden("losses.pt")[name]
    logits, teacher, y = gold["logits"], gold["teacher"], gold["y"]
    B, C = logits.shape[:2]
    S = logits[0, 0].numel()
    # the fixture applies Dice+CE and the consistency term to ALL samples; reproduce that with two calls
    def lay(t):
        t = t.reshape(t.shape[0], C, S)
        return cu(t.permute(0, 2, 1) if nhwc else t)
    ws = torch.empty(ops.ssl_loss_workspace_bytes(B, S) // 4 + 4, device=DEV)
    w = torch.tensor([gold["w"]], device=DEV)
    sup, con = torch.zeros(20, device=DEV), torch.zeros(20, device=DEV)
    ops.ssl_loss_fwd(lay(logits), None, cu(y), nhwc, B, B, C, S, w, sup, ws)
    ops.ssl_loss_fwd(lay(logits), lay(teacher), None, nhwc, B, 0, C, S, w, con, ws)
    torch.testing.assert_close(sup[0].cpu(), gold["ce"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sup[1].cpu(), gold["dice"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(con[2].cpu(), gold["mse"].mean(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close((sup[3] + con[3]).cpu(), gold["total"], rtol=1e-5, atol=1e-6)
    g1, g2 = torch.empty(B, C, S, device=DEV), torch.empty(B, C, S, device=DEV)
    ops.ssl_loss_bwd(lay(logits), None, cu(y), nhwc, B, B, C, S, w, sup, 1.0, g1, False)
    ops.ssl_loss_bwd(lay(logits), lay(teacher), None, nhwc, B, 0, C, S, w, con, 1.0, g2, False)
    torch.testing.assert_close((g1 + g2).cpu().reshape(gold["grad"].shape), gold["grad"], rtol=1e-4, atol=1e-8)
    g3 = torch.empty(B, S, C, device=DEV)
    ops.ssl_loss_bwd(lay(logits), None, cu(y), nhwc, B, B, C, S, w, sup, 2.0, g3, True)
    torch.testing.assert_close(g3.permute(0, 2, 1).cpu(), 2 * g1.cpu(), rtol=1e-6, atol=1e-9)


def test_ssl_loss_mixed_batch_matches_reference_op():
    g = torch.Generator().manual_seed(17)
    B, Lb, C, S = 6, 2, 4, 24 * 24
    logits, teacher = rnd(g, B, C, S) * 2, rnd(g, B - Lb, C, S) * 2
    y = torch.randint(0, C, (B, S), generator=g).to(torch.uint8)
    w = torch.tensor([0.07])
    ws = torch.empty(ops.ssl_loss_workspace_bytes(B, S) // 4 + 4, device=DEV)
    lb, lb_r = torch.zeros(20, device=DEV), torch.zeros(20)
    ops.ssl_loss_fwd(cu(logits), cu(teacher), cu(y), False, B, Lb, C, S, cu(w), lb, ws)
    ref.ssl_loss_fwd(logits, teacher, y, False, B, Lb, C, S, w, lb_r, None)
    torch.testing.assert_close(lb[:4].cpu(), lb_r[:4], rtol=1e-5, atol=1e-6)
    dl, dl_r = torch.empty(B, S, C, device=DEV), torch.empty(B, S, C)
    ops.ssl_loss_bwd(cu(logits), cu(teacher), cu(y), False, B, Lb, C, S, cu(w), lb, 1.0, dl, True)
    ref.ssl_loss_bwd(logits, teacher, y, False, B, Lb, C, S, w, lb_r, 1.0, dl_r, True)
    torch.testing.assert_close(dl.cpu(), dl_r, rtol=1e-4, atol=1e-9)


def test_sgd_ema_and_noise():
    g = torch.Generator().manual_seed(23)
    n = 4 * 1000 + 4
    p, gr, buf, ema = rnd(g, n), rnd(g, n), rnd(g, n), rnd(g, n)
    hp = torch.tensor([0.01, 0.9, 1e-4, 0.99, 0.01, 0.5, 0.0, 0.0])
    pd, gd, bd, ed = cu(p.clone()), cu(gr.clone()), cu(buf.clone()), cu(ema.clone())
    ops.sgd_ema_step(pd, gd, bd, ed, cu(hp), zero_grad=True)
    pr, br, er = p.clone(), buf.clone(), ema.clone()
    ref.sgd_ema_step(pr, gr.clone(), br, er, hp)
    torch.testing.assert_close(pd.cpu(), pr, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(bd.cpu(), br, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(ed.cpu(), er, rtol=1e-6, atol=1e-7)
    assert float(gd.abs().sum()) == 0.0
    ops.ema_update(ed, pd, cu(hp))
    ref.ema_update(er, pr, hp)
    torch.testing.assert_close(ed.cpu(), er, rtol=1e-6, atol=1e-7)
    x = rnd(g, 3, 1, 16, 16)
    out = torch.empty(3, 1, 16, 16, device=DEV)
    off = torch.tensor([2], dtype=torch.int64, device=DEV)
    ops.noise_add(cu(x), out, 0.1, 0.2, 31, off, 1000)
    nz = out.cpu() - x
    expect = torch.from_numpy(philox.clamp_noise(33, 1000, x.numel())).reshape(x.shape)
    torch.testing.assert_close(nz, expect, rtol=0, atol=2e-6)
    assert float(nz.abs().max()) <= 0.2 + 1e-6 and 0.05 < float(nz.std()) < 0.12


def test_errors_are_loud():
    with pytest.raises(ops.B200Error):
        ops.add(torch.zeros(8), torch.zeros(8), torch.zeros(8))            # CPU tensors: no fallback
    with pytest.raises(ops.B200Error):
        ops.bn_act_fwd(torch.zeros(4, 6, device=DEV), torch.zeros(24, device=DEV), torch.zeros(4, 6, device=DEV), 4, 6, 0.01)
    assert _lib.query("b200_device_sm") == 100


TILE_CASES = [
    # n, d, h, w, c0, c1, cout, dims
    (2, 1, 32, 32, 16, 0, 16, 2),
    (2, 1, 20, 24, 16, 0, 32, 2),          # partial tiles
    (1, 1, 16, 48, 16, 16, 16, 2),         # virtual concat
    (2, 1, 16, 16, 64, 0, 64, 2),
    (1, 1, 16, 16, 128, 128, 128, 2),
    (3, 1, 16, 16, 256, 0, 256, 2),
    (2, 1, 32, 16, 16, 0, 4, 2),           # head: cout = 4
    (1, 1, 8, 8, 32, 0, 48, 2),            # image smaller than the tile, cout not a multiple of 16
    (1, 8, 16, 16, 16, 0, 16, 3),
    (1, 6, 6, 10, 32, 0, 64, 3),           # ragged 3D volume
    (2, 4, 8, 8, 64, 0, 32, 3),
]


@pytest.mark.parametrize("case", TILE_CASES)
def test_conv_tile_kernels(case):
    n, d, h, w, c0, c1, cout, dims = case
    g = torch.Generator().manual_seed(sum(case) * 7)
    desc = ops.conv_desc(n, d, h, w, c0, c1, cout, 3, 1, 1, dims)
    assert ops.conv_tile_supported(desc)
    T, cin, M = 3 ** dims, c0 + c1, n * d * h * w
    wgt = rnd(g, cout, cin, *([3] * dims), scale=(cin * T) ** -0.5)
    bias, x0, dy = rnd(g, cout), rnd(g, M, c0), rnd(g, M, cout)
    x1 = rnd(g, M, c1) if c1 else None
    tol = TOL[False]
    # packing (the device rounds to TF32)
    for dg in (False, True):
        pk = torch.empty(ops.conv_tile_packed_floats(dg, cout, cin, T), device=DEV)
        ops.conv_tile_pack_weights(cu(wgt), pk, dg, cout, cin, T)
        pk_ref = torch.empty(ref.conv_tile_packed_floats(dg, cout, cin, T))
        ref.conv_tile_pack_weights(wgt, pk_ref, dg, cout, cin, T)
        torch.testing.assert_close(pk.cpu(), pk_ref, rtol=1e-3, atol=0)
        if dg:
            wt_d, wt_d_ref = pk, pk_ref
        else:
            wt_f, wt_f_ref = pk, pk_ref
    y, y_ref = torch.empty(M, cout, device=DEV), torch.empty(M, cout)
    ops.conv_tile_fwd(desc, cu(x0), cu(x1), wt_f, cu(bias), y)
    ref.conv_tile_fwd(desc, x0, x1, wt_f_ref, bias, y_ref)
    torch.testing.assert_close(y.cpu(), y_ref, **tol)
    y2 = torch.empty(n, cout, d * h * w, device=DEV)
    ops.conv_tile_fwd(desc, cu(x0), cu(x1), wt_f, cu(bias), y2, True)
    torch.testing.assert_close(y2.cpu(), y_ref.view(n, -1, cout).permute(0, 2, 1), **tol)
    if cout % 4 == 0:
        dx0 = torch.full((M, c0), 3.0, device=DEV)
        dx1 = torch.full((M, c1), 3.0, device=DEV) if c1 else None
        ops.conv_tile_dgrad(desc, cu(dy), wt_d, dx0, dx1, False)
        r0, r1 = torch.empty(M, c0), (torch.empty(M, c1) if c1 else None)
        ref.conv_tile_dgrad(desc, dy, wt_d_ref, r0, r1)
        torch.testing.assert_close(dx0.cpu(), r0, **tol)
        if c1:
            torch.testing.assert_close(dx1.cpu(), r1, **tol)
        ops.conv_tile_dgrad(desc, cu(dy), wt_d, dx0, dx1, True)
        torch.testing.assert_close(dx0.cpu(), 2 * r0, rtol=tol["rtol"], atol=2 * tol["atol"])
    if ops.conv_tile_supported(desc, True):
        ws = torch.empty(ops.conv_tile_wgrad_workspace_bytes(desc) // 4 + 4, device=DEV)
        dw, db = torch.empty_like(cu(wgt)), torch.empty(cout, device=DEV)
        ops.conv_tile_wgrad(desc, cu(x0), cu(x1), cu(dy), ws, dw, db)
        dw_ref, db_ref = torch.empty_like(wgt), torch.empty(cout)
        ref.conv_wgrad(desc, x0, x1, dy, None, dw_ref, db_ref)
        torch.testing.assert_close(dw.cpu(), dw_ref, rtol=tol["rtol"], atol=tol["atol"] * max(float(dw_ref.abs().max()), 1.0))
        torch.testing.assert_close(db.cpu(), db_ref, rtol=1e-4, atol=1e-4 * max(float(db_ref.abs().max()), 1.0))


UMMA_CASES = [
    # n, h, w, c0, c1, cout
    (2, 32, 32, 16, 0, 16),
    (2, 20, 24, 16, 0, 32),           # ragged tile edges
    (1, 16, 48, 16, 16, 16),          # virtual concat
    (2, 16, 16, 64, 0, 64),
    (1, 16, 16, 128, 128, 128),
    (3, 16, 16, 256, 0, 256),         # column tiles (blockIdx.y), 16 chunks
    (2, 32, 16, 16, 0, 4),            # head: cout = 4, NCHW output
    (1, 8, 8, 32, 0, 48),
    (1, 64, 256, 16, 0, 16),          # wide image: TW = 30 tiles
    (2, 40, 72, 32, 0, 32),
]


@pytest.mark.parametrize("case", UMMA_CASES)
def test_conv_umma_kernels(case):
    """tcgen05.mma / TMEM convolution (persistent warp-specialised conv_umma2 kernel) against the fp32 reference op."""
    n, h, w, c0, c1, cout = case
    g = torch.Generator().manual_seed(sum(case) * 3)
    desc = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    assert ops.conv_umma_supported(desc)
    T, cin, M = 9, c0 + c1, n * h * w
    wgt = rnd(g, cout, cin, 3, 3, scale=(cin * T) ** -0.5)
    bias, x0, dy = rnd(g, cout), rnd(g, M, c0), rnd(g, M, cout)
    x1 = rnd(g, M, c1) if c1 else None
    tol = TOL[False]
    packs = {}
    for dg in (False, True):
        pk = torch.empty(ops.conv_umma_packed_floats(dg, cout, cin, T), device=DEV)
        ops.conv_umma_pack_weights(cu(wgt), pk, dg, cout, cin, T)
        pk_ref = torch.empty(ref.conv_umma_packed_floats(dg, cout, cin, T))
        ref.conv_umma_pack_weights(wgt, pk_ref, dg, cout, cin, T)
        torch.testing.assert_close(pk.cpu(), pk_ref, rtol=1e-3, atol=0)
        packs[dg] = (pk, pk_ref)
    y_ref = torch.empty(M, cout)
    ref.conv_umma_fwd(desc, x0, x1, packs[False][1], bias, y_ref)
    if cout % 4 == 0:
        y = torch.full((M, cout), 5.0, device=DEV)
        ops.conv_umma_fwd(desc, cu(x0), cu(x1), packs[False][0], cu(bias), y)
        torch.testing.assert_close(y.cpu(), y_ref, **tol)
    y2 = torch.empty(n, cout, h * w, device=DEV)
    ops.conv_umma_fwd(desc, cu(x0), cu(x1), packs[False][0], cu(bias), y2, True)
    torch.testing.assert_close(y2.cpu(), y_ref.view(n, -1, cout).permute(0, 2, 1), **tol)
    if ops.conv_umma_supported(desc, True):
        dx0 = torch.full((M, c0), 3.0, device=DEV)
        dx1 = torch.full((M, c1), 3.0, device=DEV) if c1 else None
        ops.conv_umma_dgrad(desc, cu(dy), packs[True][0], dx0, dx1, False)
        r0, r1 = torch.empty(M, c0), (torch.empty(M, c1) if c1 else None)
        ref.conv_umma_dgrad(desc, dy, packs[True][1], r0, r1)
        torch.testing.assert_close(dx0.cpu(), r0, **tol)
        if c1:
            torch.testing.assert_close(dx1.cpu(), r1, **tol)
        ops.conv_umma_dgrad(desc, cu(dy), packs[True][0], dx0, dx1, True)
        torch.testing.assert_close(dx0.cpu(), 2 * r0, rtol=tol["rtol"], atol=2 * tol["atol"])


@pytest.mark.parametrize("case", [(2, 1, 20, 24, 16), (3, 1, 32, 32, 32), (1, 6, 10, 12, 16), (2, 4, 8, 8, 32)])
def test_first_layer_kernels(case):
    n, d, h, w, cout = case
    dims = 3 if d > 1 else 2
    g = torch.Generator().manual_seed(sum(case))
    desc = ops.conv_desc(n, d, h, w, 1, 0, cout, 3, 1, 1, dims)
    assert ops.conv_c1_supported(desc)
    T, M = 3 ** dims, n * d * h * w
    wgt, bias = rnd(g, cout, 1, *([3] * dims), scale=0.3), rnd(g, cout)
    x, dy = rnd(g, M, 1), rnd(g, M, cout)
    y, y_ref = torch.empty(M, cout, device=DEV), torch.empty(M, cout)
    ops.conv_c1_fwd(desc, cu(x), cu(wgt), cu(bias), y)
    ref.conv_c1_fwd(desc, x, wgt, bias, y_ref)
    torch.testing.assert_close(y.cpu(), y_ref, rtol=1e-5, atol=1e-5)            # FFMA: fp32-exact
    ws = torch.empty(ops.conv_c1_wgrad_workspace_bytes(desc) // 4 + 4, device=DEV)
    dw, db = torch.empty_like(cu(wgt)), torch.empty(cout, device=DEV)
    ops.conv_c1_wgrad(desc, cu(x), cu(dy), ws, dw, db)
    dw_ref, db_ref = torch.empty_like(wgt), torch.empty(cout)
    ref.conv_c1_wgrad(desc, x, dy, None, dw_ref, db_ref)
    torch.testing.assert_close(dw.cpu(), dw_ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(db.cpu(), db_ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", [(2, 4, 6, 8, 16), (1, 2, 2, 4, 32), (3, 6, 4, 2, 8)])
def test_k2s2_views_and_gemm_match_torch(case):
    """2x2x2 stride-2 conv = GEMM over the space-to-depth view, transposed conv = GEMM + depth-to-space scatter
    (code/networks/vnet.py:73,100): the two copy kernels bit-exact against the fake-ops restatement, the composed
    products against torch's conv3d / conv_transpose3d (TF32 tolerance)."""
    import torch.nn.functional as F
    from tests import fake_ops as ref
    from cv_ssl_mis_b200._lib import PACK_CONV_DGRAD_D2S, PACK_DECONV_DGRAD
    n, dd, h, w, c = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(n * dd * h * w, c, generator=g)
    xs, xs_r = torch.empty(n * dd * h * w // 8, 8 * c, device="cuda"), torch.empty(n * dd * h * w // 8, 8 * c)
    ops.s2d_gather3d(x.cuda(), xs, n, dd, h, w, c)
    ref.s2d_gather3d(x, xs_r, n, dd, h, w, c)
    assert torch.equal(xs.cpu(), xs_r)
    cout = 2 * c
    wgt, bias = torch.randn(cout, c, 2, 2, 2, generator=g) * (8 * c) ** -0.5, torch.randn(cout, generator=g)
    wp = torch.empty(ops.conv_packed_floats(PACK_CONV_DGRAD_D2S, cout, c, 8), device="cuda")
    ops.conv_pack_weights(wgt.cuda(), wp, PACK_CONV_DGRAD_D2S, cout, c, 8)
    y = torch.empty(xs.shape[0], cout, device="cuda")
    ops.linear_fwd(xs, None, wp.view(cout, 8 * c), bias.cuda(), y, xs.shape[0], cout)
    xn = x.view(n, dd, h, w, c).permute(0, 4, 1, 2, 3).double()
    want = F.conv3d(xn, wgt.double(), bias.double(), stride=2).permute(0, 2, 3, 4, 1).reshape(-1, cout)
    torch.testing.assert_close(y.cpu().double(), want, rtol=2e-2, atol=5e-3)
    # transposed: c -> c // 2 channels, output twice the size
    co = c // 2
    wt, bt = torch.randn(c, co, 2, 2, 2, generator=g) * c ** -0.5, torch.randn(co, generator=g)
    wpt = torch.empty(ops.conv_packed_floats(PACK_DECONV_DGRAD, co, c, 8), device="cuda")
    ops.conv_pack_weights(wt.cuda(), wpt, PACK_DECONV_DGRAD, co, c, 8)
    M = n * dd * h * w
    ys = torch.empty(M, 8 * co, device="cuda")
    if ops.linear_supported(M, 8 * co, c, 0):
        ops.linear_fwd(x.cuda(), None, wpt.view(8 * co, c), None, ys, M, 8 * co)
        out, out_r = torch.empty(8 * M, co, device="cuda"), torch.empty(8 * M, co)
        ops.d2s_scatter3d(ys, bt.cuda(), out, n, dd, h, w, co)
        ref.d2s_scatter3d(ys.cpu(), bt, out_r, n, dd, h, w, co)
        assert torch.equal(out.cpu(), out_r)
        want = F.conv_transpose3d(xn, wt.double(), bt.double(), stride=2).permute(0, 2, 3, 4, 1).reshape(-1, co)
        torch.testing.assert_close(out.cpu().double(), want, rtol=2e-2, atol=5e-3)


@pytest.mark.parametrize("case", [(2, 4, 6, 8, 16), (1, 2, 2, 2, 4), (1, 6, 4, 10, 32)])
def test_maxpool3d_and_trilinear_match_torch(case):
    """nn.MaxPool3d(2) fwd/bwd (ties after ReLU: first maximum in scan order) and trilinear x2 (align_corners=False) fwd/bwd
    against torch (code/networks/unet_3D.py:35-48, code/networks/utils.py:264)."""
    n, dd, h, w, c = case
    g = torch.Generator().manual_seed(sum(case))
    a = torch.relu(torch.randn(n * dd * h * w, c, generator=g))                # many exact-zero ties
    a[: a.shape[0] // 3] = 0.0
    M2 = n * dd * h * w // 8
    out, out_r = torch.empty(M2, c, device=DEV), torch.empty(M2, c)
    ops.maxpool3d_fwd(cu(a), out, n, dd, h, w, c)
    ref.maxpool3d_fwd(a, out_r, n, dd, h, w, c)
    assert torch.equal(out.cpu(), out_r)
    dp = torch.randn(M2, c, generator=g)
    da, da_r = torch.full(a.shape, 7.0, device=DEV), torch.full(a.shape, 7.0)
    ops.maxpool3d_bwd(cu(a), cu(dp), da, n, dd, h, w, c, False)
    ref.maxpool3d_bwd(a, dp, da_r, n, dd, h, w, c, False)
    assert torch.equal(da.cpu(), da_r)
    ops.maxpool3d_bwd(cu(a), cu(dp), da, n, dd, h, w, c, True)
    torch.testing.assert_close(da.cpu(), 2 * da_r, rtol=0, atol=0)
    x = torch.randn(n * dd * h * w, c, generator=g)
    y, y_r = torch.empty(8 * x.shape[0], c, device=DEV), torch.empty(8 * x.shape[0], c)
    ops.upsample3d2x_fwd(cu(x), y, n, dd, h, w, c)
    ref.upsample3d2x_fwd(x, y_r, n, dd, h, w, c)
    torch.testing.assert_close(y.cpu(), y_r, rtol=1e-6, atol=1e-6)
    dy = torch.randn(8 * x.shape[0], c, generator=g)
    dx, dx_r = torch.full(x.shape, 3.0, device=DEV), torch.full(x.shape, 3.0)
    ops.upsample3d2x_bwd(cu(dy), dx, n, dd, h, w, c, False)
    ref.upsample3d2x_bwd(dy, dx_r, n, dd, h, w, c, False)
    torch.testing.assert_close(dx.cpu(), dx_r, rtol=1e-5, atol=1e-5)
    ops.upsample3d2x_bwd(cu(dy), dx, n, dd, h, w, c, True)
    torch.testing.assert_close(dx.cpu(), 2 * dx_r, rtol=1e-5, atol=2e-5)
