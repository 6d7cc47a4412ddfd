"""Row-ring tcgen05 weight gradient (csrc/conv_row_wgrad.cu) against torch.nn.grad.conv2d_weight in fp32 on the CPU
(backward of code/networks/unet.py:37,41).  TF32 products, fp32 accumulation: the tolerance is TF32 round-off of a
K = N*H*W long dot product, relative to the largest gradient entry."""
import os

import pytest
import torch

from cv_ssl_mis_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # n, h, w, c0, c1, cout
    (2, 8, 16, 32, 0, 32),
    (3, 12, 24, 32, 32, 32),
    (2, 16, 16, 32, 0, 64),
    (2, 9, 16, 64, 0, 64),
    (2, 8, 8, 64, 64, 64),
    (2, 8, 8, 64, 0, 128),
    (1, 6, 8, 128, 128, 128),
    (2, 4, 8, 128, 0, 256),
    (1, 4, 4, 256, 0, 256),
    (2, 8, 16, 16, 0, 16),
    (3, 10, 32, 16, 16, 16),
    (1, 5, 120, 32, 0, 32),
]


def _ref(n, h, w, x, dy, cin, cout):
    xn = x.view(n, h, w, cin).permute(0, 3, 1, 2).double()
    dyn = dy.view(n, h, w, cout).permute(0, 3, 1, 2).double()
    return torch.nn.grad.conv2d_weight(xn, (cout, cin, 3, 3), dyn, stride=1, padding=1).float()


def _run(case, env=None):
    n, h, w, c0, c1, cout = case
    g = torch.Generator().manual_seed(sum(case))
    M, cin = n * h * w, c0 + c1
    x0 = torch.randn(M, c0, generator=g)
    x1 = torch.randn(M, c1, generator=g) if c1 else None
    dy = torch.randn(M, cout, generator=g)
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    assert ops.conv_row_wgrad_supported(d)
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        ws = torch.full((ops.conv_row_wgrad_workspace_bytes(d) // 4 + 4,), float("nan"), device=DEV)
        dw = torch.full((cout, cin, 3, 3), float("nan"), device=DEV)
        ops.conv_row_wgrad(d, x0.to(DEV), None if x1 is None else x1.to(DEV), dy.to(DEV), ws, dw)
        torch.cuda.synchronize()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    ref = _ref(n, h, w, x, dy, cin, cout)
    return dw.cpu(), ref


@pytest.mark.parametrize("case", CASES)
def test_row_wgrad_matches_torch(case):
    dw, ref = _run(case)
    scale = float(ref.abs().max())
    torch.testing.assert_close(dw, ref, rtol=2e-2, atol=4e-3 * scale)


@pytest.mark.parametrize("env", [{"B200_WGRAD_KHM1": "1"}])
@pytest.mark.parametrize("case", [CASES[1], CASES[3], CASES[10]])
def test_row_wgrad_unstacked_variants(case, env):
    """The kh-stacked MMAs (three dy rows as the row groups of one instruction) and one MMA per dy row agree."""
    dw, ref = _run(case, env)
    scale = float(ref.abs().max())
    torch.testing.assert_close(dw, ref, rtol=2e-2, atol=4e-3 * scale)


def test_row_wgrad_accumulate():
    n, h, w, c0, c1, cout = CASES[0]
    g = torch.Generator().manual_seed(5)
    M = n * h * w
    x0, dy = torch.randn(M, c0, generator=g).to(DEV), torch.randn(M, cout, generator=g).to(DEV)
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    ws = torch.empty(ops.conv_row_wgrad_workspace_bytes(d) // 4 + 4, device=DEV)
    dw = torch.empty(cout, c0, 3, 3, device=DEV)
    ops.conv_row_wgrad(d, x0, None, dy, ws, dw)
    once = dw.clone()
    ops.conv_row_wgrad(d, x0, None, dy, ws, dw, accumulate=True)
    torch.testing.assert_close(dw, 2 * once, rtol=1e-6, atol=0)


# ------------------------------------------------------------------------------------------------ 3D (3x3x3)
CASES3 = [
    # n, d, h, w, c0, c1, cout   (backward of code/networks/vnet.py:28 and of the UNETR residual blocks)
    (2, 6, 8, 16, 32, 0, 32),
    (1, 5, 6, 12, 64, 0, 64),
    (2, 3, 5, 6, 128, 0, 128),
    (1, 4, 4, 4, 256, 0, 256),
    (2, 4, 6, 16, 16, 0, 16),
    (1, 3, 5, 8, 32, 32, 32),
    (1, 1, 7, 10, 64, 0, 32),
    (1, 2, 4, 8, 16, 16, 16),
]


@pytest.mark.parametrize("case", CASES3)
def test_row_wgrad_3d_matches_torch(case):
    n, dd, h, w, c0, c1, cout = case
    g = torch.Generator().manual_seed(sum(case))
    M, cin = n * dd * h * w, c0 + c1
    x0 = torch.randn(M, c0, generator=g)
    x1 = torch.randn(M, c1, generator=g) if c1 else None
    dy = torch.randn(M, cout, generator=g)
    d = ops.conv_desc(n, dd, h, w, c0, c1, cout, 3, 1, 1, 3)
    assert ops.conv_row_wgrad_supported(d)
    ws = torch.full((ops.conv_row_wgrad_workspace_bytes(d) // 4 + 4,), float("nan"), device=DEV)
    dw = torch.full((cout, cin, 3, 3, 3), float("nan"), device=DEV)
    db = torch.full((cout,), float("nan"), device=DEV)
    ops.conv_row_wgrad(d, x0.to(DEV), None if x1 is None else x1.to(DEV), dy.to(DEV), ws, dw, False, db)
    torch.cuda.synchronize()
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    xn = x.view(n, dd, h, w, cin).permute(0, 4, 1, 2, 3).double()
    dyn = dy.view(n, dd, h, w, cout).permute(0, 4, 1, 2, 3).double()
    ref = torch.nn.grad.conv3d_weight(xn, (cout, cin, 3, 3, 3), dyn, stride=1, padding=1).float()
    scale = float(ref.abs().max())
    torch.testing.assert_close(dw.cpu(), ref, rtol=2e-2, atol=4e-3 * scale)
    assert float(db.abs().max()) == 0.0
    once = dw.clone()
    ops.conv_row_wgrad(d, x0.to(DEV), None if x1 is None else x1.to(DEV), dy.to(DEV), ws, dw, True)
    torch.testing.assert_close(dw, 2 * once, rtol=1e-6, atol=0)
