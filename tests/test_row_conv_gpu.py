"""Row-ring tcgen05 forward / data gradient (csrc/conv_row.cu) against torch fp64 convolutions on the CPU
(code/networks/unet.py:37,41 and their backward at the 256^2 / 128^2 levels), including the fused BatchNorm statistics
(code/networks/unet.py:38,42).  TF32 products with fp32 accumulation: tolerance = TF32 round-off of a 9*Cin long dot product."""
import pytest
import torch
import torch.nn.functional as F

from cv_ssl_mis_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # n, h, w, c0, c1, cout
    (2, 5, 128, 32, 0, 32),
    (1, 4, 128, 32, 32, 32),     # does not fit (two resident planes + 74 KB of weights): must report unsupported for fwd
    (2, 3, 128, 16, 0, 32),
    (2, 6, 256, 16, 0, 16),
    (1, 7, 256, 16, 16, 16),
    (3, 2, 128, 32, 0, 16),
    (1, 3, 384, 16, 0, 16),
    (2, 150, 256, 16, 0, 16),    # ~2 rows per CTA: rolling window across image boundaries
    (1, 320, 256, 16, 16, 16),
    (3, 130, 128, 32, 0, 32),
]


def _mk(case, seed=0):
    n, h, w, c0, c1, cout = case
    g = torch.Generator().manual_seed(seed + sum(case))
    M, cin = n * h * w, c0 + c1
    x0 = torch.randn(M, c0, generator=g)
    x1 = torch.randn(M, c1, generator=g) if c1 else None
    wgt = torch.randn(cout, cin, 3, 3, generator=g) * (cin * 9) ** -0.5
    bias = torch.randn(cout, generator=g)
    dy = torch.randn(M, cout, generator=g)
    return x0, x1, wgt, bias, dy


def _nchw(t, n, h, w, c):
    return t.view(n, h, w, c).permute(0, 3, 1, 2).double()


@pytest.mark.parametrize("case", CASES)
def test_row_fwd_and_stats(case):
    n, h, w, c0, c1, cout = case
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    if not ops.conv_row_supported(d, False):
        assert case == CASES[1]
        return
    x0, x1, wgt, bias, _ = _mk(case)
    M, cin = n * h * w, c0 + c1
    wpk = torch.empty(ops.conv_row_packed_floats(d, False), device=DEV)
    ops.conv_row_pack_weights(d, False, wgt.to(DEV), wpk)
    y = torch.full((M, cout), float("nan"), device=DEV)
    nb = ops.conv_row_stats_blocks(d)
    part = torch.full((nb * 2 * cout,), float("nan"), dtype=torch.float64, device=DEV)
    ops.conv_row_fwd(d, x0.to(DEV), None if x1 is None else x1.to(DEV), wpk, bias.to(DEV), y, part)
    torch.cuda.synchronize()
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    ref = F.conv2d(_nchw(x, n, h, w, cin), wgt.double(), bias.double(), padding=1).permute(0, 2, 3, 1).reshape(M, cout)
    torch.testing.assert_close(y.cpu().double(), ref, rtol=2e-2, atol=5e-3)
    # the fused statistics are the exact sums of the values that were stored
    s = part.view(nb, 2, cout).sum(0).cpu()
    yd = y.cpu().double()
    torch.testing.assert_close(s[0], yd.sum(0), rtol=1e-6, atol=1e-6 * M)
    torch.testing.assert_close(s[1], (yd * yd).sum(0), rtol=1e-6, atol=1e-6 * M)
    # and b200_bn_finalize turns them into the BatchNorm state torch computes
    gamma, beta = torch.rand(cout) + 0.5, torch.randn(cout)
    rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    state = torch.empty(4 * cout, device=DEV)
    ops.bn_finalize(part, nb, M, cout, gamma.to(DEV), beta.to(DEV), 1e-5, 0.1, rm, rv, state)
    mean, var = yd.mean(0), yd.var(0, unbiased=False)
    st = state.view(4, cout).cpu().double()
    torch.testing.assert_close(st[0], mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(st[1], 1 / torch.sqrt(var + 1e-5), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rm.cpu().double(), 0.1 * mean, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", CASES)
def test_row_dgrad(case):
    n, h, w, c0, c1, cout = case
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    if not ops.conv_row_supported(d, True):
        pytest.skip("data gradient of this shape is served by another kernel")
    _, _, wgt, _, dy = _mk(case, 1)
    M, cin = n * h * w, c0 + c1
    wpk = torch.empty(ops.conv_row_packed_floats(d, True), device=DEV)
    ops.conv_row_pack_weights(d, True, wgt.to(DEV), wpk)
    dx0 = torch.full((M, c0), float("nan"), device=DEV)
    dx1 = torch.full((M, c1), float("nan"), device=DEV) if c1 else None
    ops.conv_row_dgrad(d, dy.to(DEV), wpk, dx0, dx1)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(_nchw(dy, n, h, w, cout), wgt.double(), padding=1).permute(0, 2, 3, 1).reshape(M, cin)
    got = dx0.cpu() if dx1 is None else torch.cat([dx0.cpu(), dx1.cpu()], 1)
    torch.testing.assert_close(got.double(), ref, rtol=2e-2, atol=5e-3)
    # accumulate: TMA reduce-add on top of the stored result
    ops.conv_row_dgrad(d, dy.to(DEV), wpk, dx0, dx1, accumulate=True)
    got2 = dx0.cpu() if dx1 is None else torch.cat([dx0.cpu(), dx1.cpu()], 1)
    torch.testing.assert_close(got2, 2 * got, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[3], CASES[4], CASES[5]])
@pytest.mark.parametrize("dgrad", [False, True])
def test_row_pack_batch_equals_single_and_restatement(case, dgrad):
    """The one-launch batch packer (kind 3), the stand-alone packer and the index restatement in tests/fake_ops.py agree."""
    from tests import fake_ops
    n, h, w, c0, c1, cout = case
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    m = ops.conv_row_supported(d, dgrad)
    if not m:
        pytest.skip("not served by the row kernels")
    assert m == fake_ops.conv_row_supported(d, dgrad)
    wgt = torch.randn(cout, c0 + c1, 3, 3, generator=torch.Generator().manual_seed(1))
    nf = ops.conv_row_packed_floats(d, dgrad)
    assert nf == fake_ops.conv_row_packed_floats(d, dgrad)
    single, batch = torch.empty(nf, device=DEV), torch.empty(nf, device=DEV)
    wd = wgt.to(DEV)
    ops.conv_row_pack_weights(d, dgrad, wd, single)
    table = torch.tensor([[wd.data_ptr(), batch.data_ptr(), 3, m - 8, cout, c0 + c1, 9, nf]], dtype=torch.int64, device=DEV)
    ops.conv_pack_batch(table, 1, 8)
    restated = torch.empty(nf)
    fake_ops.TF32_ROUND = True
    try:
        fake_ops.conv_row_pack_weights(d, dgrad, wgt, restated)
    finally:
        fake_ops.TF32_ROUND = False
    assert torch.equal(single, batch)
    assert torch.equal(single.cpu(), restated)


# ------------------------------------------------------------------------------------------------ halo-block kernel
BLK_CASES = [
    # n, h, w, c0, c1, cout
    (2, 64, 64, 32, 0, 64),
    (3, 64, 64, 64, 0, 64),
    (1, 64, 64, 64, 64, 64),
    (2, 32, 32, 64, 0, 128),
    (2, 32, 32, 128, 128, 128),
    (3, 16, 16, 128, 0, 256),
    (2, 16, 16, 256, 0, 256),
    (2, 8, 8, 32, 0, 32),
    (1, 20, 24, 32, 32, 96),
    (2, 14, 14, 64, 0, 32),
]


@pytest.mark.parametrize("case", BLK_CASES)
def test_blk_fwd_and_stats(case):
    n, h, w, c0, c1, cout = case
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    assert ops.conv_blk_supported(d, False)
    x0, x1, wgt, bias, _ = _mk(case)
    M, cin = n * h * w, c0 + c1
    wpk = torch.empty(9 * cout * cin, device=DEV)
    ops.conv_blk_pack_weights(wgt.to(DEV), wpk, False, cout, cin)
    y = torch.full((M, cout), float("nan"), device=DEV)
    nb = ops.conv_blk_stats_blocks(d)
    part = torch.full((nb * 2 * cout,), float("nan"), dtype=torch.float64, device=DEV)
    ops.conv_blk_fwd(d, x0.to(DEV), None if x1 is None else x1.to(DEV), wpk, bias.to(DEV), y, part)
    torch.cuda.synchronize()
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    ref = F.conv2d(_nchw(x, n, h, w, cin), wgt.double(), bias.double(), padding=1).permute(0, 2, 3, 1).reshape(M, cout)
    torch.testing.assert_close(y.cpu().double(), ref, rtol=2e-2, atol=5e-3)
    s = part.view(nb, 2, cout).sum(0).cpu()
    yd = y.cpu().double()
    torch.testing.assert_close(s[0], yd.sum(0), rtol=1e-5, atol=1e-5 * M)
    torch.testing.assert_close(s[1], (yd * yd).sum(0), rtol=1e-5, atol=1e-5 * M)


@pytest.mark.parametrize("case", BLK_CASES)
def test_blk_dgrad(case):
    n, h, w, c0, c1, cout = case
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    if not ops.conv_blk_supported(d, True):
        pytest.skip("data gradient of this shape is served by another kernel")
    _, _, wgt, _, dy = _mk(case, 1)
    M, cin = n * h * w, c0 + c1
    wpk = torch.empty(9 * cout * cin, device=DEV)
    ops.conv_blk_pack_weights(wgt.to(DEV), wpk, True, cout, cin)
    dx0 = torch.full((M, c0), float("nan"), device=DEV)
    dx1 = torch.full((M, c1), float("nan"), device=DEV) if c1 else None
    ops.conv_blk_dgrad(d, dy.to(DEV), wpk, dx0, dx1)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(_nchw(dy, n, h, w, cout), wgt.double(), padding=1).permute(0, 2, 3, 1).reshape(M, cin)
    got = dx0.cpu() if dx1 is None else torch.cat([dx0.cpu(), dx1.cpu()], 1)
    torch.testing.assert_close(got.double(), ref, rtol=2e-2, atol=5e-3)
    ops.conv_blk_dgrad(d, dy.to(DEV), wpk, dx0, dx1, accumulate=True)
    got2 = dx0.cpu() if dx1 is None else torch.cat([dx0.cpu(), dx1.cpu()], 1)
    torch.testing.assert_close(got2, 2 * got, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------ halo-block kernel, 3D
BLK3_CASES = [
    # n, d, h, w, cin, cout  (code/networks/vnet.py:28 at the 32 / 64 / 128 / 256-channel levels, plus ragged shapes)
    (2, 8, 12, 40, 32, 32),
    (1, 28, 28, 20, 64, 64),
    (2, 14, 14, 10, 128, 128),
    (2, 7, 7, 5, 256, 256),
    (1, 5, 9, 7, 32, 64),
    (3, 3, 4, 6, 64, 32),
    (1, 1, 6, 13, 32, 96),
    (2, 6, 10, 96, 16, 16),      # 16-channel level (64-byte halo rows)
    (1, 5, 7, 24, 16, 32),
    (1, 4, 6, 12, 32, 16),
    (2, 3, 3, 8, 16, 48),
]


def _mk3(case, seed=0):
    n, dd, h, w, cin, cout = case
    g = torch.Generator().manual_seed(seed + sum(case))
    M = n * dd * h * w
    return (torch.randn(M, cin, generator=g), torch.randn(cout, cin, 3, 3, 3, generator=g) * (cin * 27) ** -0.5,
            torch.randn(cout, generator=g), torch.randn(M, cout, generator=g))


def _ncdhw(t, n, dd, h, w, c):
    return t.view(n, dd, h, w, c).permute(0, 4, 1, 2, 3).double()


@pytest.mark.parametrize("case", BLK3_CASES)
def test_blk3d_fwd_and_stats(case):
    n, dd, h, w, cin, cout = case
    d = ops.conv_desc(n, dd, h, w, cin, 0, cout, 3, 1, 1, 3)
    mode = ops.conv_blk_supported(d, False) - 8
    assert mode >= 0
    x, wgt, bias, _ = _mk3(case)
    M = n * dd * h * w
    wpk = torch.empty(27 * cout * cin, device=DEV)
    ops.conv_blk_pack_weights(wgt.to(DEV), wpk, mode, cout, cin, 27)
    y = torch.full((M, cout), float("nan"), device=DEV)
    nb = ops.conv_blk_stats_blocks(d)
    part = torch.full((nb * 2 * cout,), float("nan"), dtype=torch.float64, device=DEV)
    ops.conv_blk_fwd(d, x.to(DEV), None, wpk, bias.to(DEV), y, part)
    torch.cuda.synchronize()
    ref = F.conv3d(_ncdhw(x, n, dd, h, w, cin), wgt.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1).reshape(M, cout)
    torch.testing.assert_close(y.cpu().double(), ref, rtol=2e-2, atol=5e-3)
    s = part.view(nb, 2, cout).sum(0).cpu()
    yd = y.cpu().double()
    torch.testing.assert_close(s[0], yd.sum(0), rtol=1e-5, atol=1e-5 * M)
    torch.testing.assert_close(s[1], (yd * yd).sum(0), rtol=1e-5, atol=1e-5 * M)


@pytest.mark.parametrize("case", BLK3_CASES)
def test_blk3d_dgrad(case):
    n, dd, h, w, cin, cout = case
    d = ops.conv_desc(n, dd, h, w, cin, 0, cout, 3, 1, 1, 3)
    mode = ops.conv_blk_supported(d, True) - 8
    assert mode >= 1
    _, wgt, _, dy = _mk3(case, 1)
    M = n * dd * h * w
    wpk = torch.empty(27 * cout * cin, device=DEV)
    ops.conv_blk_pack_weights(wgt.to(DEV), wpk, mode, cout, cin, 27)
    dx = torch.full((M, cin), float("nan"), device=DEV)
    ops.conv_blk_dgrad(d, dy.to(DEV), wpk, dx, None)
    torch.cuda.synchronize()
    ref = F.conv_transpose3d(_ncdhw(dy, n, dd, h, w, cout), wgt.double(), padding=1).permute(0, 2, 3, 4, 1).reshape(M, cin)
    torch.testing.assert_close(dx.cpu().double(), ref, rtol=2e-2, atol=5e-3)
    got = dx.clone()
    ops.conv_blk_dgrad(d, dy.to(DEV), wpk, dx, None, accumulate=True)
    torch.testing.assert_close(dx, 2 * got, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------ production shapes, on device
@pytest.mark.parametrize("case", [(2, 96, 96, 96, 16), (2, 48, 48, 48, 32), (4, 24, 24, 24, 64), (4, 12, 12, 12, 128), (4, 6, 6, 6, 256)])
def test_3d_tcgen05_kernels_at_config4_shapes_match_exact_kernels(case):
    """The VNet layer shapes of BASELINE config 4 (96^3 volumes): halo-block forward / data gradient and the row-ring weight
    gradient (TF32 products) against the generic implicit-GEMM kernels in their exact 3xTF32 mode, all on the device -- every
    tile of the full-size launch is compared, error budget = TF32 round-off of a 27 * C long dot product."""
    from cv_ssl_mis_b200._lib import PACK_CONV_FWD, PACK_CONV_DGRAD
    n, dd, h, w, c = case
    g = torch.Generator(device=DEV).manual_seed(sum(case))
    M = n * dd * h * w
    x = torch.randn(M, c, device=DEV, generator=g)
    dy = torch.randn(M, c, device=DEV, generator=g)
    wgt = torch.randn(c, c, 3, 3, 3, device=DEV, generator=g) * (c * 27) ** -0.5
    bias = torch.randn(c, device=DEV, generator=g)
    d = ops.conv_desc(n, dd, h, w, c, 0, c, 3, 1, 1, 3)

    def generic_pack(mode):
        out = torch.empty(ops.conv_packed_floats(mode, c, c, 27), device=DEV)
        ops.conv_pack_weights(wgt, out, mode, c, c, 27)
        return out

    def check(got, want, what, tol=4e-3):
        scale = float(want.abs().max())
        err = float((got - want).abs().max())
        assert err <= tol * scale, (what, err, scale)

    # forward (+ fused BatchNorm sums)
    mode = ops.conv_blk_supported(d, False) - 8
    assert mode >= 0
    wpk = torch.empty(27 * c * c, device=DEV)
    ops.conv_blk_pack_weights(wgt, wpk, mode, c, c, 27)
    y = torch.full((M, c), float("nan"), device=DEV)
    nb = ops.conv_blk_stats_blocks(d)
    part = torch.zeros(nb * 2 * c, dtype=torch.float64, device=DEV)
    ops.conv_blk_fwd(d, x, None, wpk, bias, y, part)
    y_ref = torch.empty_like(y)
    ops.conv_fwd(d, x, None, generic_pack(PACK_CONV_FWD), bias, y_ref, False, True)
    check(y, y_ref, "forward")
    s = part.view(nb, 2, c).sum(0)
    torch.testing.assert_close(s[0], y.double().sum(0), rtol=1e-6, atol=1e-6 * M)
    torch.testing.assert_close(s[1], (y.double() ** 2).sum(0), rtol=1e-6, atol=1e-6 * M)
    # data gradient
    mode = ops.conv_blk_supported(d, True) - 8
    ops.conv_blk_pack_weights(wgt, wpk, mode, c, c, 27)
    dx = torch.full((M, c), float("nan"), device=DEV)
    ops.conv_blk_dgrad(d, dy, wpk, dx, None)
    dx_ref = torch.empty_like(dx)
    ops.conv_dgrad(d, dy, generic_pack(PACK_CONV_DGRAD), dx_ref, None, False, True)
    check(dx, dx_ref, "data gradient")
    # weight gradient
    assert ops.conv_row_wgrad_supported(d)
    ws = torch.empty(ops.conv_row_wgrad_workspace_bytes(d) // 4 + 4, device=DEV)
    dw = torch.full((c, c, 3, 3, 3), float("nan"), device=DEV)
    ops.conv_row_wgrad(d, x, None, dy, ws, dw)
    ws2 = torch.empty(ops.conv_wgrad_workspace_bytes(d) // 4 + 4, device=DEV)
    dw_ref, db_ref = torch.empty_like(dw), torch.empty(c, device=DEV)
    ops.conv_wgrad(d, x, None, dy, ws2, dw_ref, db_ref, False, True)
    check(dw, dw_ref, "weight gradient", tol=6e-3)
