"""tcgen05/TMA Linear kernels (gemm_umma.cu) against a plain torch fp32 reference of torch.nn.functional.linear and its
autograd (the ops the reference's Mlp / qkv / proj / reduction / expand layers run,
code/networks/swin_transformer_unet_skip_expand_decoder_sys.py:19-25,115-150,346,378).
Tolerance: TF32 operands (10-bit mantissa, truncated by the tensor core), fp32 accumulation -> max abs error below
3e-3 of the largest reference magnitude."""
import pytest
import torch

from cv_ssl_mis_b200 import ops

pytestmark = pytest.mark.gpu

SHAPES = [  # M, O, c0, c1
    (3136, 288, 96, 0),        # stage-1 qkv
    (784, 192, 192, 192),      # concat_back_dim (virtual concat)
    (196, 1152, 384, 0),
    (98, 3072, 768, 0),        # stage-4 fc1, M < 128
    (3136, 96, 48, 0),         # patch embedding, K = 48 (partial k-block)
    (1000, 20, 96, 0),         # ragged rows and columns
    (1000, 96, 32, 64),        # concat with unequal halves
    (4500, 384, 96, 0),        # several splits in wgrad
]
TOL = 3e-3


def _err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("M,O,c0,c1", SHAPES)
def test_linear_fwd_dgrad_wgrad(M, O, c0, c1):
    assert ops.linear_supported(M, O, c0, c1)
    g = torch.Generator(device="cuda").manual_seed(M + O + c0)
    I = c0 + c1
    x0 = torch.randn(M, c0, device="cuda", generator=g)
    x1 = torch.randn(M, c1, device="cuda", generator=g) if c1 else None
    w = torch.randn(O, I, device="cuda", generator=g) * 0.1
    b = torch.randn(O, device="cuda", generator=g)
    dy = torch.randn(M, O, device="cuda", generator=g)
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y_ref = x @ w.t() + b
        dx_ref = dy @ w
        dw_ref = dy.t() @ x
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev

    y = torch.full((M, O), float("nan"), device="cuda")
    ops.linear_fwd(x0, x1, w, b, y, M, O)
    assert _err(y, y_ref) < TOL
    y2 = torch.empty_like(y)
    ops.linear_fwd(x0, x1, w, None, y2, M, O)
    assert _err(y2, y_ref - b) < TOL

    dx0 = torch.full((M, c0), float("nan"), device="cuda")
    dx1 = torch.full((M, c1), float("nan"), device="cuda") if c1 else None
    ops.linear_dgrad(dy, w, dx0, dx1, False, M, O)
    assert _err(dx0, dx_ref[:, :c0]) < TOL
    if c1:
        assert _err(dx1, dx_ref[:, c0:]) < TOL
    # accumulate
    ops.linear_dgrad(dy, w, dx0, dx1, True, M, O)
    assert _err(dx0, 2 * dx_ref[:, :c0]) < TOL

    ws = torch.empty(max(ops.linear_wgrad_workspace_bytes(M, O, I) // 4, 4), device="cuda")
    dw = torch.full((O, I), float("nan"), device="cuda")
    ops.linear_wgrad(x0, x1, dy, dw, ws, M, O)
    assert _err(dw, dw_ref) < TOL
    dw_again = torch.empty_like(dw)
    ops.linear_wgrad(x0, x1, dy, dw_again, ws, M, O)
    assert torch.equal(dw, dw_again), "wgrad must be deterministic"
    ops.linear_wgrad(x0, x1, dy, dw, ws, M, O, accumulate=True)
    assert _err(dw, 2 * dw_ref) < TOL


def test_linear_rejects_bad_shapes():
    assert not ops.linear_supported(128, 8, 96, 0)       # O < 16
    assert not ops.linear_supported(128, 96, 48, 48)     # concat needs c0 % 32 == 0
    assert not ops.linear_supported(128, 96, 6, 0)


@pytest.mark.parametrize("M,C", [(3136, 96), (500, 288), (98, 3072), (12544, 768)])
def test_colsum_any_width(M, C):
    """bias gradient of a Linear layer = column sums of dy (deterministic two-step reduction)"""
    g = torch.randn(M, C, device="cuda", generator=torch.Generator(device="cuda").manual_seed(C))
    ws = torch.empty(ops.colsum_workspace_bytes(M, C) // 4 + 4, device="cuda")
    out = torch.empty(C, device="cuda")
    ops.colsum(g, M, C, out, ws)
    torch.testing.assert_close(out, g.double().sum(0).float(), rtol=1e-5, atol=1e-4)
    out2 = torch.empty(C, device="cuda")
    ops.colsum(g, M, C, out2, ws)
    assert torch.equal(out, out2)
