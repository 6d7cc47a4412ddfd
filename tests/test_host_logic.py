"""Host-side launch schedules against the oracle, on CPU, with tests/fake_ops.py standing in for the
C ABI (so wiring bugs -- wrong buffer, wrong order, missing accumulate -- are caught without a GPU)."""
import numpy as np
import pytest
import torch

from oracle import philox, ssl_oracle as O
from tests import fake_ops
from cv_ssl_mis_b200.networks import unet as unet_mod
from cv_ssl_mis_b200.networks._engine import FlatParams, Runtime


@pytest.fixture()
def fake(monkeypatch):
    fake_ops.install(monkeypatch)
    return fake_ops


def unet_masks(seed, B, H, W):
    masks = []
    for i, (c, p) in enumerate(zip(O.UNET_FT, O.UNET_DROPOUT)):
        h, w = H >> i, W >> i
        m = philox.keep_mask(seed, i, B * h * w, c, p, 1)
        masks.append(torch.from_numpy(m).reshape(B, h, w, c).permute(0, 3, 1, 2).contiguous())
    return masks


@pytest.mark.parametrize("dropout", [False, True])
def test_unet_plan_matches_oracle(fake, dropout):
    torch.manual_seed(11)
    net = unet_mod.UNet(1, 4)
    if not dropout:
        unet_mod_drop = [0.0] * 5
    B, H, W = 2, 32, 32
    g = torch.Generator().manual_seed(3)
    x = torch.rand(B, 1, H, W, generator=g)
    y = torch.randint(0, 4, (B, H, W), generator=g).to(torch.uint8)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    flat = FlatParams(net, "cpu")
    rt = Runtime("cpu", seed=77)
    old = list(unet_mod.DROPOUT)
    try:
        if not dropout:
            unet_mod.DROPOUT[:] = [0.0] * 5
        plan = unet_mod.UNetPlan(net, rt, B, H, W, True)
    finally:
        unet_mod.DROPOUT[:] = old
    logits = plan.forward(x, train=True).view(B, 4, H, W)

    keys = O.param_keys(sd0)
    leaf = {k: (v.clone().requires_grad_(True) if k in keys else v.clone()) for k, v in sd0.items()}
    masks = unet_masks(77, B, H, W) if dropout else None
    ref = O.unet_forward(leaf, x, True, masks, update_running=True)
    torch.testing.assert_close(logits, ref, rtol=1e-4, atol=1e-5)
    loss, _, _ = O.supervised_loss(ref, y, 4)
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys] + [ref])
    dlogits = grads[-1]
    plan.backward(dlogits.permute(0, 2, 3, 1).reshape(B * H * W, 4).contiguous())
    named = dict(net.named_parameters())
    for k, gr in zip(keys, grads[:-1]):
        torch.testing.assert_close(named[k].grad, gr, rtol=2e-3, atol=2e-6, msg=lambda m, k=k: f"{k}: {m}")
    # BN running statistics follow torch's train-mode update
    sd1 = net.state_dict()
    for k in sd1:
        if "running" in k:
            torch.testing.assert_close(sd1[k], leaf[k], rtol=1e-5, atol=1e-6)


def test_flat_params_alias_module_parameters():
    torch.manual_seed(0)
    net = unet_mod.UNet(1, 4)
    ref = [p.detach().clone() for p in net.parameters()]
    flat = FlatParams(net, "cpu")
    assert flat.numel == 1813764                      # UNet parameter count
    for p, r, o in zip(net.parameters(), ref, flat.offsets):
        assert torch.equal(p.detach(), r)
        assert o % 4 == 0
        assert p.data_ptr() == flat.data[o:].data_ptr()
        assert p.grad.data_ptr() == flat.grad[o:].data_ptr()
    flat.data.zero_()
    assert all(float(p.abs().sum()) == 0 for p in net.parameters())


def test_mean_teacher_trainer_matches_oracle(fake):
    """Three iterations straddling the iter<1000 consistency gate: trainer schedule vs oracle.mt2d_step."""
    from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
    torch.manual_seed(21)
    student, teacher = unet_mod.UNet(1, 4, seed=101), unet_mod.UNet(1, 4, seed=202)
    for p in teacher.parameters():
        p.detach_()
    s_sd = {k: v.clone() for k, v in student.state_dict().items()}
    t_sd = {k: v.clone() for k, v in teacher.state_dict().items()}
    B, Lb, H, W = 4, 2, 32, 32
    tr = MeanTeacherTrainer(student, teacher, batch_size=B, labeled_bs=Lb, patch_size=(H, W), num_classes=4,
                            start_iter=999, noise_seed=555)
    tr.lr = O.poly_lr(0.01, 998, 30000)
    bufs = {k: torch.zeros_like(s_sd[k]) for k in O.param_keys(s_sd)}
    g = torch.Generator().manual_seed(8)
    for step in range(3):
        it = 999 + step
        x = torch.rand(B, 1, H, W, generator=g)
        y = torch.randint(0, 4, (B, H, W), generator=g).to(torch.uint8)
        lossbuf = tr.step(x, y).clone()
        off = step + 1                                         # seed_off after this step's bump
        noise = torch.from_numpy(philox.clamp_noise(555 + off, 1000, (B - Lb) * H * W)).reshape(B - Lb, 1, H, W)
        r = O.mt2d_step(s_sd, t_sd, bufs, x, y, noise, it, labeled_bs=Lb,
                        student_masks=unet_masks(101 + off, B, H, W), teacher_masks=unet_masks(202 + off, B - Lb, H, W))
        torch.testing.assert_close(lossbuf[3], r["loss"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(lossbuf[0], r["ce"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(lossbuf[1], r["dice"], rtol=1e-4, atol=1e-5)
        if it >= 1000:
            torch.testing.assert_close(lossbuf[2], r["cons"], rtol=1e-3, atol=1e-6)
        sd_now, td_now = student.state_dict(), teacher.state_dict()
        for k in s_sd:
            if s_sd[k].dtype.is_floating_point:
                torch.testing.assert_close(sd_now[k], s_sd[k], rtol=2e-3, atol=1e-5, msg=lambda m, k=k: f"student {k}: {m}")
                torch.testing.assert_close(td_now[k], t_sd[k], rtol=2e-3, atol=1e-5, msg=lambda m, k=k: f"teacher {k}: {m}")
    assert tr.iter_num == 1002
    assert abs(tr.lr - O.poly_lr(0.01, 1001, 30000)) < 1e-15


def vnet_drops(seed, B, nf=16):
    """(drop5, drop9) channel keep-masks [B, C] of one VNet forward (rng streams 0 and 1, Dropout3d p = 0.5)."""
    d5 = torch.from_numpy(philox.keep_mask(seed, 0, B, 16 * nf, 0.5, 2, 1)).reshape(B, 16 * nf)
    d9 = torch.from_numpy(philox.keep_mask(seed, 1, B, nf, 0.5, 2, 1)).reshape(B, nf)
    return d5, d9


def test_vnet_plan_matches_oracle(fake):
    from cv_ssl_mis_b200.networks import vnet as vnet_mod
    torch.manual_seed(13)
    net = vnet_mod.VNet(1, 2, has_dropout=True)
    B, P = 2, 32                      # 16^3 would leave 2 samples per channel in the deepest BatchNorm: ill-conditioned
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, 1, P, P, P, generator=g)
    y = torch.randint(0, 2, (B, P, P, P), generator=g)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    FlatParams(net, "cpu")
    rt = Runtime("cpu", seed=55)
    plan = vnet_mod.VNetPlan(net, rt, B, P, P, P, True)
    logits = plan.forward(x, train=True).view(B, 2, P, P, P)
    keys = O.param_keys(sd0)
    leaf = {k: (v.clone().requires_grad_(True) if k in keys else v.clone()) for k, v in sd0.items()}
    d5, d9 = vnet_drops(55, B)
    ref = O.vnet_forward(leaf, x, True, d5, d9, update_running=True)
    torch.testing.assert_close(logits, ref, rtol=1e-3, atol=2e-4)      # 30 BatchNorms deep, 16 samples in the deepest
    loss, _, _ = O.supervised_loss(ref, y, 2)
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys] + [ref])
    plan.backward(grads[-1].permute(0, 2, 3, 4, 1).reshape(B * P ** 3, 2).contiguous())
    named = dict(net.named_parameters())
    for k, gr in zip(keys, grads[:-1]):
        rel = float((named[k].grad - gr).norm() / (gr.norm() + 1e-12))
        # a conv bias in front of a train-mode BatchNorm has an analytically zero gradient (only round-off remains)
        # the fp32 oracle itself sits 0.5-1.2% from an fp64 run on this tiny batch (30 BatchNorm backward passes)
        assert rel < 5e-2 or float(gr.norm()) < 1e-5, (k, rel)
    sd1 = net.state_dict()
    for k in sd1:
        if "running" in k:
            torch.testing.assert_close(sd1[k], leaf[k], rtol=1e-5, atol=1e-6)


def test_uamt_trainer_matches_oracle(fake):
    from cv_ssl_mis_b200.networks import vnet as vnet_mod
    from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
    torch.manual_seed(31)
    student, teacher = vnet_mod.VNet(1, 2, has_dropout=True, seed=301), vnet_mod.VNet(1, 2, has_dropout=True, seed=402)
    for p in teacher.parameters():
        p.detach_()
    s_sd = {k: v.clone() for k, v in student.state_dict().items()}
    t_sd = {k: v.clone() for k, v in teacher.state_dict().items()}
    B, Lb, P, T = 4, 2, 32, 8
    U = B - Lb
    tr = MeanTeacherTrainer(student, teacher, batch_size=B, labeled_bs=Lb, patch_size=(P, P, P), num_classes=2,
                            start_iter=2000, consistency_gate_iters=0, uncertainty_T=T, noise_seed=777)
    tr.lr = O.poly_lr(0.01, 1999, 30000)
    bufs = {k: torch.zeros_like(s_sd[k]) for k in O.param_keys(s_sd)}
    g = torch.Generator().manual_seed(9)
    t_epoch = 0
    for step in range(2):
        it = 2000 + step
        x = torch.randn(B, 1, P, P, P, generator=g)
        y = torch.randint(0, 2, (B, P, P, P), generator=g)
        lossbuf = tr.step(x, y).clone()
        noises, tdrops = [], []
        for k in range(1 + T // 2):                           # teacher RNG epoch advances before every teacher forward
            t_epoch += 1
            nb = U if k == 0 else 2 * U
            noises.append(torch.from_numpy(philox.clamp_noise(777 + t_epoch, 1000, nb * P ** 3)).reshape(nb, 1, P, P, P))
            tdrops.append(vnet_drops(402 + t_epoch, nb))
        r = O.uamt3d_step(s_sd, t_sd, bufs, x, y, noises, it, labeled_bs=Lb, T=T,
                          student_drops=vnet_drops(301 + step + 1, B), teacher_drops=tdrops)
        torch.testing.assert_close(lossbuf[3], r["loss"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(lossbuf[2], r["cons"], rtol=1e-3, atol=1e-6)
        sd_now, td_now = student.state_dict(), teacher.state_dict()
        for k in s_sd:
            if s_sd[k].dtype.is_floating_point:
                torch.testing.assert_close(sd_now[k], s_sd[k], rtol=5e-3, atol=2e-5, msg=lambda m, k=k: f"student {k}: {m}")
                torch.testing.assert_close(td_now[k], t_sd[k], rtol=5e-3, atol=2e-5, msg=lambda m, k=k: f"teacher {k}: {m}")


# ------------------------------------------------------------------ Swin-UNet / Cross-Teaching
SWIN_SMALL = dict(img_size=64, embed_dim=32, num_heads=(1, 2, 4, 8), window_size=4, drop_path_rate=0.2)


@pytest.mark.parametrize("drop_path", [0.0, 0.3])
def test_swin_plan_matches_oracle(fake, drop_path):
    """Forward/backward tape (aliasing residual gradients, pooled buffers, accumulate flags) against autograd."""
    from oracle import swin_oracle as SO
    from tests import swin_common as SC
    from cv_ssl_mis_b200.networks import swin_unet as S
    torch.manual_seed(5)
    net = S.SwinUnet(dict(SWIN_SMALL, drop_path_rate=drop_path), num_classes=4, seed=99)
    g = torch.Generator().manual_seed(6)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    sd0 = SC.swin_sd(net)
    B = 2
    x = torch.randn(B, 1, 64, 64, generator=g)
    y = torch.randint(0, 4, (B, 64, 64), generator=g)
    net.materialize()
    plan = net._get_plan(B, True)
    net._bump_seed()
    logits = plan.forward(x, train=True).view(B, 4, 64, 64)
    cfg = SO.swin_config(sd0, 64, 4, drop_path)
    keys = [k for k, v in sd0.items() if v.dtype.is_floating_point and not k.endswith("attn_mask")]
    leaf = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in sd0.items()}
    keeps = SC.drop_keeps(99 + 1, B, cfg["depths"], drop_path) if drop_path > 0 else None
    ref = SO.swin_unet_forward(leaf, x, cfg, True, keeps)
    torch.testing.assert_close(logits, ref, rtol=1e-4, atol=1e-5)
    loss, _, _ = O.supervised_loss(ref, y, 4)
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys] + [ref])
    plan.backward(grads[-1].permute(0, 2, 3, 1).reshape(-1, 4).contiguous())
    named = dict(net.swin_unet.named_parameters())
    for k, gr in zip(keys, grads[:-1]):
        torch.testing.assert_close(named[k].grad, gr, rtol=2e-3, atol=1e-7, msg=lambda m, k=k: f"{k}: {m}")
    # eval mode: DropPath is the identity
    with torch.no_grad():
        ev = net._get_plan(B, False).forward(x, train=False).view(B, 4, 64, 64)
        torch.testing.assert_close(ev, SO.swin_unet_forward(sd0, x, cfg, False), rtol=1e-4, atol=1e-5)


def test_swin_autograd_bridge(fake):
    """`loss.backward()` through SwinUnet.forward, as the reference trainers call it."""
    from oracle import swin_oracle as SO
    from tests import swin_common as SC
    from cv_ssl_mis_b200.networks import swin_unet as S
    torch.manual_seed(2)
    net = S.SwinUnet(dict(SWIN_SMALL, drop_path_rate=0.0), num_classes=3)
    sd0 = SC.swin_sd(net)
    x = torch.randn(1, 1, 64, 64)
    out = net(x)
    out.square().mean().backward()
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd0.items()}
    ref = SO.swin_unet_forward(leaf, x, SO.swin_config(sd0, 64, 4, 0.0), True)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
    (gw,) = torch.autograd.grad(ref.square().mean(), leaf["layers.2.blocks.0.mlp.fc1.weight"])
    torch.testing.assert_close(net.swin_unet.layers[2].blocks[0].mlp.fc1.weight.grad, gw, rtol=2e-3, atol=1e-8)
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 1, 32, 32))


def test_cross_teaching_trainer_matches_oracle(fake):
    """Two iterations of CrossTeachingTrainer vs oracle.ct2d_step (losses, both models' weights, LR schedule)."""
    from oracle import swin_oracle as SO
    from tests import swin_common as SC
    from cv_ssl_mis_b200.networks import swin_unet as S
    from cv_ssl_mis_b200.trainers import CrossTeachingTrainer
    torch.manual_seed(31)
    m1 = unet_mod.UNet(1, 4, seed=11)
    m2 = S.SwinUnet(dict(SWIN_SMALL), num_classes=4, seed=22)
    sd1 = {k: v.clone() for k, v in m1.state_dict().items()}
    sd2 = SC.swin_sd(m2)
    B, Lb, P = 4, 2, 64
    it0 = 23999
    tr = CrossTeachingTrainer(m1, m2, batch_size=B, labeled_bs=Lb, patch_size=(P, P), num_classes=4, start_iter=it0)
    cfg = SO.swin_config(sd2, P, 4, 0.2)
    bufs1 = {k: torch.zeros_like(sd1[k]) for k in O.param_keys(sd1)}
    bufs2 = {k: torch.zeros_like(v) for k, v in sd2.items() if v.dtype.is_floating_point}
    g = torch.Generator().manual_seed(9)
    for step in range(2):
        it = it0 + step
        x = torch.rand(B, 1, P, P, generator=g)
        y = SC.blocky_labels(g, B, P, P, 4)
        got = tr.step(x, y, read_loss=True)
        off = step + 1
        r = SO.ct2d_step(sd1, sd2, bufs1, bufs2, x, y, it, cfg, labeled_bs=Lb, masks1=unet_masks(11 + off, B, P, P),
                         drop_keep=SC.drop_keeps(22 + off, B, cfg["depths"], 0.2))
        want = [r["ce1"], r["dice1"], r["ps1"], r["model1_loss"], r["ce2"], r["dice2"], r["ps2"], r["model2_loss"]]
        torch.testing.assert_close(torch.tensor(got), torch.stack(want), rtol=1e-4, atol=1e-5)
        now1, now2 = m1.state_dict(), SC.swin_sd(m2)
        for k in sd1:
            if sd1[k].dtype.is_floating_point:
                torch.testing.assert_close(now1[k], sd1[k], rtol=2e-3, atol=1e-5, msg=lambda m, k=k: f"unet {k}: {m}")
        for k in sd2:
            if sd2[k].dtype.is_floating_point:
                torch.testing.assert_close(now2[k], sd2[k], rtol=2e-3, atol=1e-6, msg=lambda m, k=k: f"swin {k}: {m}")
    assert tr.iter_num == it0 + 2
    assert abs(tr.lr - O.poly_lr(0.01, it0 + 2, 30000)) < 1e-15


# ------------------------------------------------------------------ UNETR (config 5)
UNETR_SMALL = dict(img_size=(32, 32, 32), feature_size=8, hidden_size=64, mlp_dim=128, num_heads=2, conv_block=True,
                   res_block=True)


def test_unetr_plan_matches_oracle(fake):
    """Forward logits and every parameter gradient of the UNETR launch schedule (ViT tape + per-sample conv decoder,
    InstanceNorm as batch-of-one BatchNorm, weight gradients accumulated over samples) against the restated oracle."""
    from oracle import unetr_oracle as UO
    from cv_ssl_mis_b200.networks import unetr as U
    torch.manual_seed(5)
    net = U.UNETR(1, 2, **UNETR_SMALL)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 1, 32, 32, 32, generator=g)
    y = torch.randint(0, 2, (2, 32, 32, 32), generator=g)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss, logits = UO.fully_supervised_loss(leaf, x, y, 2, 2)
    loss.backward()
    out = net(x)
    torch.testing.assert_close(out, logits, rtol=1e-4, atol=1e-4)
    O.supervised_loss(out, y, 2)[0].backward()
    for k, p in net.named_parameters():
        if leaf[k].grad is None:          # cls_token: unused when classification=False
            continue
        # fp32 reassociation through 12 transformer blocks and 5 InstanceNorm levels (depends on the CPU thread count):
        # relative L2 error per tensor, single entries may move by a few per cent of the largest one
        ref = leaf[k].grad
        assert float((p.grad - ref).norm() / (ref.norm() + 1e-20)) < 2e-2, k
        assert float((p.grad - ref).abs().max()) / (float(ref.abs().max()) + 1e-12) < 1e-1, k
    # MONAI's state_dict schema
    for key in ("vit.patch_embedding.patch_embeddings.1.weight", "vit.patch_embedding.position_embeddings",
                "vit.patch_embedding.cls_token", "vit.blocks.11.attn.qkv.weight", "vit.blocks.0.mlp.linear1.bias",
                "vit.norm.weight", "encoder1.layer.conv3.conv.weight", "encoder2.transp_conv_init.conv.weight",
                "encoder2.blocks.1.1.conv2.conv.weight", "decoder5.conv_block.conv3.conv.weight",
                "decoder2.transp_conv.conv.weight", "out.conv.conv.bias"):
        assert key in sd, key
    assert "encoder2.blocks.0.1.conv3.conv.weight" not in sd and "vit.blocks.0.attn.qkv.bias" not in sd


def test_unetr_fully_supervised_trainer_step(fake):
    """One fully-supervised step (train_fully_supervised_3D_ViT.py: SGD on 0.5 (CE + Dice)) through MeanTeacherTrainer."""
    from oracle import unetr_oracle as UO
    from cv_ssl_mis_b200.networks import unetr as U
    from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
    torch.manual_seed(6)
    net = U.UNETR(1, 2, **UNETR_SMALL)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(10)
    x = torch.randn(2, 1, 32, 32, 32, generator=g)
    y = torch.randint(0, 2, (2, 32, 32, 32), generator=g)
    tr = MeanTeacherTrainer(net, None, batch_size=2, labeled_bs=2, patch_size=(32, 32, 32), num_classes=2, base_lr=0.01)
    pending = tr.submit(x, y)                                             # pipelined form of step(..., read_loss=True)
    ce, dice, cons, total = pending.result()
    assert tr.iter_num == 1 and [ce, dice, cons, total] == tr.lossbuf[:4].tolist()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss, _ = UO.fully_supervised_loss(leaf, x, y, 2, 2)
    loss.backward()
    assert abs(total - float(loss)) < 1e-4 * abs(float(loss))
    k = "decoder2.conv_block.conv1.conv.weight"
    expect = sd[k] - 0.01 * (leaf[k].grad + 1e-4 * sd[k])                 # first SGD step: buf = g + wd * p
    torch.testing.assert_close(dict(net.named_parameters())[k].detach(), expect, rtol=1e-3, atol=1e-6)


def test_mean_teacher_with_swin_unets_matches_oracle(fake):
    """code/train_mean_teacher_ViT.py:201-233: the Mean-Teacher loop over two ViT_seg (Swin-UNet) models.  One step with
    the consistency term live (iter >= 1000): loss terms, the student's SGD update and the teacher's EMA update."""
    from oracle import swin_oracle as SO
    from tests import swin_common as SC
    from cv_ssl_mis_b200.networks import swin_unet as S
    from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
    torch.manual_seed(41)
    cfgd = dict(SWIN_SMALL, drop_path_rate=0.0)
    student, teacher = S.SwinUnet(dict(cfgd), num_classes=4, seed=5), S.SwinUnet(dict(cfgd), num_classes=4, seed=6)
    s_sd, t_sd = SC.swin_sd(student), SC.swin_sd(teacher)
    B, Lb, P, it = 4, 2, 64, 1500
    tr = MeanTeacherTrainer(student, teacher, batch_size=B, labeled_bs=Lb, patch_size=(P, P), num_classes=4, start_iter=it,
                            noise_seed=7)
    g = torch.Generator().manual_seed(12)
    x = torch.rand(B, 1, P, P, generator=g)
    y = SC.blocky_labels(g, B, P, P, 4)
    ce, dice, cons, total = tr.step(x, y, read_loss=True)
    cfg = SO.swin_config(s_sd, P, 4, 0.0)
    noise = torch.from_numpy(philox.clamp_noise(7 + 1, 1000, (B - Lb) * P * P)).reshape(B - Lb, 1, P, P)
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in s_sd.items()}
    s_logits = SO.swin_unet_forward(leaf, x, cfg, True)
    with torch.no_grad():
        t_logits = SO.swin_unet_forward(t_sd, x[Lb:] + noise, cfg, True)
    w = O.consistency_weight(it)
    loss, ce_r, dice_r, cons_r = O.mt_loss(s_logits, t_logits, y, Lb, 4, w)
    loss.backward()
    torch.testing.assert_close(torch.tensor([ce, dice, cons, total]), torch.stack([ce_r, dice_r, cons_r, loss]).detach(),
                               rtol=1e-4, atol=1e-6)
    alpha = O.ema_alpha(it)
    now_s, now_t = SC.swin_sd(student), SC.swin_sd(teacher)
    for k in ("layers.0.blocks.0.mlp.fc1.weight", "layers_up.3.blocks.1.attn.qkv.bias", "output.weight"):
        new = s_sd[k] - 0.01 * (leaf[k].grad + 1e-4 * s_sd[k])
        torch.testing.assert_close(now_s[k], new, rtol=2e-3, atol=1e-6, msg=lambda m, k=k: f"student {k}: {m}")
        torch.testing.assert_close(now_t[k], alpha * t_sd[k] + (1 - alpha) * new, rtol=2e-3, atol=1e-6,
                                   msg=lambda m, k=k: f"teacher {k}: {m}")


def test_cross_pseudo_supervision_trainer_matches_oracle(fake):
    """code/train_cross_pseudo_supervision_2D.py:176-212 -- two UNets, each supervised by the other's argmax pseudo labels
    through a CE term: CrossTeachingTrainer(pseudo_loss="ce").  One iteration: both losses and both SGD updates."""
    import torch.nn.functional as F
    from cv_ssl_mis_b200.trainers import CrossTeachingTrainer
    torch.manual_seed(51)
    m1, m2 = unet_mod.UNet(1, 4, seed=11), unet_mod.UNet(1, 4, seed=22)
    sd = [{k: v.clone() for k, v in m.state_dict().items()} for m in (m1, m2)]
    B, Lb, P, it = 4, 2, 32, 4000
    tr = CrossTeachingTrainer(m1, m2, batch_size=B, labeled_bs=Lb, patch_size=(P, P), num_classes=4, start_iter=it,
                              pseudo_loss="ce")
    g = torch.Generator().manual_seed(14)
    x = torch.rand(B, 1, P, P, generator=g)
    y = torch.randint(0, 4, (B, P, P), generator=g).to(torch.uint8)
    got = tr.step(x, y, read_loss=True)
    leaf = [{k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v.clone()) for k, v in s.items()} for s in sd]
    outs = [O.unet_forward(leaf[i], x, True, unet_masks((11, 22)[i] + 1, B, P, P)) for i in range(2)]
    w = O.consistency_weight(it)                                         # get_current_consistency_weight(iter_num // 150)
    want, total = [], 0.0
    for i in range(2):
        sup, ce, dice = O.supervised_loss(outs[i][:Lb], y[:Lb], 4)                       # :182-185
        pseudo = torch.argmax(torch.softmax(outs[1 - i][Lb:].detach(), 1), 1)            # :187-188
        ps = F.cross_entropy(outs[i][Lb:], pseudo)                                       # :190-191
        m = sup + w * ps                                                                 # :193-194
        total = total + m
        want += [ce, dice, ps, m]
    total.backward()
    torch.testing.assert_close(torch.tensor(got), torch.stack(want).detach(), rtol=1e-4, atol=1e-5)
    lr = O.poly_lr(0.01, it, 30000)                                      # the rate in effect for iteration `it`
    for i, m in enumerate((m1, m2)):
        k = "decoder.up4.conv.conv_conv.0.weight"
        new = sd[i][k] - lr * (leaf[i][k].grad + 1e-4 * sd[i][k])
        torch.testing.assert_close(m.state_dict()[k], new, rtol=2e-3, atol=1e-6)
    assert tr.iter_num == it + 1


def test_uamt_2d_unet_trainer_matches_oracle(fake):
    """code/train_uncertainty_aware_mean_teacher_2D.py:147-201 -- the 2D twin of UAMT: UNet student/teacher, T = 8 stochastic
    teacher passes (4 forwards of the twice-repeated unlabeled batch, fresh noise and dropout each), entropy mask.
    One step of MeanTeacherTrainer(uncertainty_T=8) against the oracle's loss with OUR Philox noise / dropout streams."""
    from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
    torch.manual_seed(61)
    student, teacher = unet_mod.UNet(1, 4, seed=11), unet_mod.UNet(1, 4, seed=22)
    s_sd = {k: v.clone() for k, v in student.state_dict().items()}
    t_sd = {k: v.clone() for k, v in teacher.state_dict().items()}
    B, Lb, P, it, T = 4, 2, 32, 2000, 8
    U = B - Lb
    tr = MeanTeacherTrainer(student, teacher, batch_size=B, labeled_bs=Lb, patch_size=(P, P), num_classes=4, start_iter=it,
                            noise_seed=7, uncertainty_T=T, consistency_gate_iters=0)
    g = torch.Generator().manual_seed(16)
    x = torch.rand(B, 1, P, P, generator=g)
    y = torch.randint(0, 4, (B, P, P), generator=g).to(torch.uint8)
    ce, dice, cons, total = tr.step(x, y, read_loss=True)

    noise = lambda off, n: torch.from_numpy(philox.clamp_noise(7 + off, 1000, n * P * P)).reshape(n, 1, P, P)
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v.clone()) for k, v in s_sd.items()}
    s_logits = O.unet_forward(leaf, x, True, unet_masks(11 + 1, B, P, P))
    with torch.no_grad():
        # the teacher's RNG epoch is bumped before every teacher forward: pass 0 -> +1, MC pass i -> +2 + i
        t_logits = O.unet_forward(t_sd, x[Lb:] + noise(1, U), True, unet_masks(22 + 1, U, P, P))
        x_rep = x[Lb:].repeat(2, 1, 1, 1)
        mc = [O.unet_forward(t_sd, x_rep + noise(2 + i, 2 * U), True, unet_masks(22 + 2 + i, 2 * U, P, P)) for i in range(T // 2)]
    w = O.consistency_weight(it)
    thr = (0.75 + 0.25 * O.sigmoid_rampup(it, 30000)) * np.log(2)
    loss, ce_r, dice_r, cons_r, mask = O.uamt_loss(s_logits, t_logits, mc, y, Lb, 4, w, thr, T)
    torch.testing.assert_close(torch.tensor([ce, dice, cons, total]), torch.stack([ce_r, dice_r, cons_r, loss]).detach(),
                               rtol=1e-4, atol=1e-6)
    loss.backward()
    k = "encoder.down2.maxpool_conv.1.conv_conv.0.weight"
    new = s_sd[k] - 0.01 * (leaf[k].grad + 1e-4 * s_sd[k])
    torch.testing.assert_close(student.state_dict()[k], new, rtol=2e-3, atol=1e-6)


def test_ict_trainer_matches_oracle(fake):
    """code/train_interpolation_consistency_training_2D.py:150-193 -- input mix-up of unlabeled pairs, teacher probabilities
    mixed with the same factors, MSE consistency.  One ICTTrainer step against the oracle pieces, then the SGD / EMA update."""
    from cv_ssl_mis_b200.trainers import ICTTrainer
    torch.manual_seed(71)
    student, teacher = unet_mod.UNet(1, 4, seed=11), unet_mod.UNet(1, 4, seed=22)
    s_sd = {k: v.clone() for k, v in student.state_dict().items()}
    t_sd = {k: v.clone() for k, v in teacher.state_dict().items()}
    B, Lb, P, it = 8, 4, 32, 700                      # below the MT2D gate: ICT applies the consistency term from the start
    h = Lb // 2
    tr = ICTTrainer(student, teacher, batch_size=B, labeled_bs=Lb, patch_size=(P, P), num_classes=4, start_iter=it)
    g = torch.Generator().manual_seed(18)
    x = torch.rand(B, 1, P, P, generator=g)
    y = torch.randint(0, 4, (B, P, P), generator=g).to(torch.uint8)
    f = torch.tensor([0.3, 0.85])
    ce, dice, cons, total = tr.step(x, y, read_loss=True, mix_factors=f)

    fm = f.view(h, 1, 1, 1)
    u0, u1 = x[Lb:Lb + h], x[Lb + h:]
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v.clone()) for k, v in s_sd.items()}
    s_in = torch.cat([x[:Lb], u0 * (1 - fm) + u1 * fm], 0)                                           # :163-167
    outputs = O.unet_forward(leaf, s_in, True, unet_masks(11 + 1, Lb + h, P, P))
    with torch.no_grad():                                                                            # :170-176
        p0 = torch.softmax(O.unet_forward(t_sd, u0, True, unet_masks(22 + 1, h, P, P)), 1)
        p1 = torch.softmax(O.unet_forward(t_sd, u1, True, unet_masks(22 + 2, h, P, P)), 1)
        mixed = p0 * (1 - fm) + p1 * fm
    sup, ce_r, dice_r = O.supervised_loss(outputs[:Lb], y[:Lb], 4)
    cons_r = torch.mean((torch.softmax(outputs, 1)[Lb:] - mixed) ** 2)                               # :184-185
    w = O.consistency_weight(it)
    assert w > 0
    loss = sup + w * cons_r
    loss.backward()
    torch.testing.assert_close(torch.tensor([ce, dice, cons, total]), torch.stack([ce_r, dice_r, cons_r, loss]).detach(),
                               rtol=1e-4, atol=1e-6)
    k = "decoder.up1.conv1x1.weight"
    new = s_sd[k] - 0.01 * (leaf[k].grad + 1e-4 * s_sd[k])
    torch.testing.assert_close(student.state_dict()[k], new, rtol=2e-3, atol=1e-6)
    alpha = O.ema_alpha(it)
    torch.testing.assert_close(teacher.state_dict()[k], alpha * t_sd[k] + (1 - alpha) * new, rtol=2e-3, atol=1e-6)
    # default mix factors come from numpy's Beta(alpha, alpha)
    tr2 = ICTTrainer(unet_mod.UNet(1, 4, seed=1), unet_mod.UNet(1, 4, seed=2), batch_size=B, labeled_bs=Lb, patch_size=(P, P),
                     num_classes=4, mix_seed=5)
    out = tr2.step(x, y, read_loss=True)
    assert all(v == v for v in out) and 0.0 <= float(tr2.mix.min()) and float(tr2.mix.max()) <= 1.0


def unet3d_drops(seed, B, P, filters=(16, 32, 64, 128, 256), p=0.3):
    """(center, up1) element-wise keep masks / (1 - p) of one unet_3D forward: Philox stream 2 b + site per sample."""
    S4, S0 = (P // 16) ** 3, P ** 3
    m0 = [torch.from_numpy(philox.keep_mask(seed, 2 * b, S4, filters[4], p, 1)) for b in range(B)]
    m1 = [torch.from_numpy(philox.keep_mask(seed, 2 * b + 1, S0, filters[0], p, 1)) for b in range(B)]
    cl = lambda ms, side, c: torch.stack(ms).reshape(B, side, side, side, c).permute(0, 4, 1, 2, 3) / (1 - p)
    return cl(m0, P // 16, filters[4]), cl(m1, P, filters[0])


@pytest.mark.parametrize("dropout", [False, True])
def test_unet3d_plan_matches_oracle(fake, dropout):
    """unet_3D launch schedule (per-sample InstanceNorm plans, 3-D pooling / trilinear up-sampling, virtual concats,
    element-wise dropout) against the functional oracle and its autograd gradients."""
    from oracle import unet3d_oracle as U3
    from cv_ssl_mis_b200.networks import unet_3d as u3
    net = u3.unet_3D(n_classes=2, in_channels=1)
    sd = U3.fixture_state_dict(31)
    net.load_state_dict(sd)
    B, P = 2, 32
    x, y = U3.fixture_inputs(32, B, P)
    FlatParams(net, "cpu")
    rt = Runtime("cpu", seed=77)
    plan = u3.UNet3DPlan(net, rt, B, P, P, P, True)
    logits = plan.forward(x, train=dropout).view(B, 2, P, P, P)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = U3.unet3d_forward(leaf, x, unet3d_drops(77, B, P) if dropout else None)
    torch.testing.assert_close(logits, ref, rtol=1e-3, atol=2e-4)
    loss, _, _ = O.supervised_loss(ref, y, 2)
    keys = list(sd.keys())
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys] + [ref], allow_unused=True)
    plan.backward(grads[-1].permute(0, 2, 3, 4, 1).reshape(B * P ** 3, 2).contiguous())
    named = dict(net.named_parameters())
    for k, gr in zip(keys, grads[:-1]):
        if k.endswith("0.bias"):          # a bias in front of InstanceNorm: analytically zero gradient
            assert float(named[k].grad.abs().max()) < 1e-4
            continue
        rel = float((named[k].grad - gr).norm() / (gr.norm() + 1e-12))
        assert rel < 2e-2, (k, rel)


def test_cross_pseudo_supervision_3d_trainer_matches_oracle(fake, monkeypatch):
    """code/train_cross_pseudo_supervision_3D.py:152-176 over two unet_3Ds (dropout off): both losses and one SGD update of
    CrossTeachingTrainer(pseudo_loss="ce") on 3-D patches with int64 labels."""
    import torch.nn.functional as F
    from oracle import unet3d_oracle as U3
    from cv_ssl_mis_b200.networks import unet_3d as u3
    from cv_ssl_mis_b200.trainers import CrossTeachingTrainer
    monkeypatch.setattr(u3, "P_DROP", 0.0)
    sds = [U3.fixture_state_dict(61), U3.fixture_state_dict(62)]
    nets = []
    for sd in sds:
        n = u3.unet_3D(n_classes=2, in_channels=1)
        n.load_state_dict(sd)
        nets.append(n)
    B, Lb, P, it = 2, 1, 32, 3000              # 16^3 would leave a single voxel per channel in `center` (InstanceNorm undefined)
    tr = CrossTeachingTrainer(nets[0], nets[1], batch_size=B, labeled_bs=Lb, patch_size=(P, P, P), num_classes=2, start_iter=it,
                              label_dtype=torch.int64, pseudo_loss="ce")
    x, y = U3.fixture_inputs(63, B, P)
    got = tr.step(x, y, read_loss=True)
    leaf = [{k: v.clone().requires_grad_(True) for k, v in sd.items()} for sd in sds]
    outs = [U3.unet3d_forward(leaf[i], x) for i in range(2)]
    w = O.consistency_weight(it)
    want, total = [], 0.0
    for i in range(2):
        sup, ce, dice = O.supervised_loss(outs[i][:Lb], y[:Lb], 2)
        pseudo = torch.argmax(torch.softmax(outs[1 - i][Lb:].detach(), 1), 1)
        ps = F.cross_entropy(outs[i][Lb:], pseudo)
        m = sup + w * ps
        total = total + m
        want += [ce, dice, ps, m]
    total.backward()
    torch.testing.assert_close(torch.tensor(got), torch.stack(want).detach(), rtol=1e-4, atol=1e-5)
    lr = O.poly_lr(0.01, it, 30000)
    for i, n in enumerate(nets):
        k = "up_concat1.conv.conv2.0.weight"
        new = sds[i][k] - lr * (leaf[i][k].grad + 1e-4 * sds[i][k])
        torch.testing.assert_close(n.state_dict()[k], new, rtol=2e-3, atol=1e-6)
