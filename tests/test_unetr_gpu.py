"""UNETR path on the GPU (config 5): the new kernels against plain torch fp32 references, then the network's logits,
loss and gradients against the restated oracle (oracle/unetr_oracle.py -- parity unpinned, see its header).
Tolerances: fp32 elementwise ops 1e-5; attention (FFMA, fp32) 1e-4; `exact` (3xTF32) network: logits 2e-3, gradients
1e-2 of the largest entry; TF32 network (tcgen05 GEMMs, TF32 convolutions): logits 5e-2 of the largest logit, loss 1e-2,
per-tensor gradients 25 % in relative L2 norm (cancelling sums, see the test)."""
import pytest
import torch
import torch.nn.functional as F

from cv_ssl_mis_b200 import ops
from oracle import ssl_oracle as O
from oracle import unetr_oracle as UO

pytestmark = pytest.mark.gpu


def test_patch3d_gather():
    x = torch.randn(2, 2, 32, 48, 16, device="cuda")
    P = 16
    y = torch.empty(2 * 2 * 3 * 1, P ** 3 * 2, device="cuda")
    ops.patch3d_gather(x, y, 2, 2, 32, 48, 16, P)
    ref = x.reshape(2, 2, 2, P, 3, P, 1, P).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(y.shape)
    assert torch.equal(y, ref)


@pytest.mark.parametrize("B,N,heads,hd", [(2, 216, 12, 64), (1, 8, 2, 32), (3, 50, 3, 16), (1, 256, 1, 64)])
def test_mha_fwd_bwd(B, N, heads, hd):
    C = heads * hd
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.randn(B * N, 3 * C, device="cuda", generator=g)
    dout = torch.randn(B * N, C, device="cuda", generator=g)
    out = torch.empty(B * N, C, device="cuda")
    probs = torch.empty(ops.mha_probs_floats(B, N, heads), device="cuda")
    ops.mha_fwd(qkv, out, probs, B, N, heads, hd)
    leaf = qkv.clone().requires_grad_(True)
    q, k, v = leaf.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(B * N, C)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(probs.view(B, heads, N, N), att, rtol=1e-4, atol=1e-6)
    ref.backward(dout)
    dqkv = torch.full_like(qkv, float("nan"))
    ws = torch.empty_like(probs)
    ops.mha_bwd(qkv, probs, dout, dqkv, ws, B, N, heads, hd)
    torch.testing.assert_close(dqkv, leaf.grad, rtol=1e-3, atol=1e-4)


def test_add_lrelu():
    a, b = torch.randn(4096, device="cuda"), torch.randn(4096, device="cuda")
    out = torch.empty_like(a)
    ops.add_lrelu_fwd(a, b, out, 0.01)
    torch.testing.assert_close(out, F.leaky_relu(a + b, 0.01))
    g = torch.randn_like(a)
    dx = torch.empty_like(a)
    ops.lrelu_bwd(out, g, dx, 0.01)
    torch.testing.assert_close(dx, torch.where(a + b > 0, g, 0.01 * g))


SMALL = dict(img_size=(32, 32, 32), feature_size=8, hidden_size=64, mlp_dim=128, num_heads=2, conv_block=True, res_block=True)
MEDIUM = dict(img_size=(64, 64, 64), feature_size=16, hidden_size=192, mlp_dim=384, num_heads=3, conv_block=True, res_block=True)


def _run(cfg, exact, B=2, seed=3):
    from cv_ssl_mis_b200.networks.unetr import UNETR
    torch.manual_seed(seed)
    net = UNETR(1, 2, **cfg, exact=exact)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    S = cfg["img_size"]
    x = torch.randn(B, 1, *S, generator=g)
    low = torch.randint(0, 2, (B, S[0] // 8, S[1] // 8, S[2] // 8), generator=g)
    y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_ref, logits_ref = UO.fully_supervised_loss(leaf, x, y, cfg["num_heads"], 2)
    loss_ref.backward()
    net = net.cuda()
    out = net(x.cuda())
    loss = O.supervised_loss(out, y.cuda(), 2)[0]
    loss.backward()
    return net, out.detach().cpu(), float(loss.detach()), leaf, logits_ref.detach(), float(loss_ref.detach())


def _grad_err(net, leaf, l2=False):
    """worst per-tensor error: max |g - ref| / max |ref|, or the relative L2 error ||g - ref|| / ||ref||"""
    worst, worst_k = 0.0, None
    for k, p in net.named_parameters():
        if leaf[k].grad is None:
            continue
        d = p.grad.cpu() - leaf[k].grad
        e = float(d.norm() / (leaf[k].grad.norm() + 1e-20)) if l2 else float(d.abs().max()) / (float(leaf[k].grad.abs().max()) + 1e-12)
        if e > worst:
            worst, worst_k = e, k
    return worst, worst_k


def test_unetr_exact_matches_oracle():
    net, out, loss, leaf, logits_ref, loss_ref = _run(SMALL, exact=True)
    assert float((out - logits_ref).abs().max()) < 2e-3 * float(logits_ref.abs().max())
    assert abs(loss - loss_ref) < 1e-4 * abs(loss_ref)
    worst, k = _grad_err(net, leaf)
    assert worst < 1e-2, (k, worst)


@pytest.mark.parametrize("cfg", [SMALL, MEDIUM])
def test_unetr_tf32_matches_oracle(cfg):
    net, out, loss, leaf, logits_ref, loss_ref = _run(cfg, exact=False)
    assert float((out - logits_ref).abs().max()) < 5e-2 * float(logits_ref.abs().max())
    assert abs(loss - loss_ref) < 1e-2 * abs(loss_ref)
    worst, k = _grad_err(net, leaf, l2=True)
    # TF32 products through 12 transformer blocks and 5 InstanceNorm conv levels; the worst tensors are LayerNorm biases
    # in front of the attention, whose gradient is a heavily cancelling sum over tokens (measured 0.10 at MEDIUM; the
    # 3xTF32 `exact` schedule above pins the same quantities to 1e-2)
    assert worst < 0.25, (k, worst)


def test_unetr_full_size_step_is_finite_and_learns():
    """BASELINE config 5 shape (96^3, bs2, ViT-B): size-independent properties -- finite loss, loss decreases over a few
    SGD steps on a fixed batch, CUDA-graph replay equals the eager schedule."""
    from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
    from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
    torch.manual_seed(0)
    net = net_factory_3d("unetr", 1, 2)
    tr = MeanTeacherTrainer(net, None, batch_size=2, labeled_bs=2, patch_size=(96, 96, 96), num_classes=2, base_lr=0.01,
                            use_cuda_graph=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 1, 96, 96, 96, generator=g).cuda()
    low = torch.randint(0, 2, (2, 12, 12, 12), generator=g)
    y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3).cuda()
    losses = [tr.step(x, y, read_loss=True)[3] for _ in range(6)]
    assert all(l == l and l < 10 for l in losses), losses
    assert losses[-1] < losses[0], losses
