"""VNet and the uncertainty-aware Mean-Teacher step (BASELINE config 4) on the GPU against the reference fixture,
the oracle, and size-independent properties at the real 96^3 patch size."""
import pytest
import torch

from oracle import philox, ssl_oracle as O
from cv_ssl_mis_b200 import ops
from cv_ssl_mis_b200.networks import vnet as vnet_mod
from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
from tests import fake_ops as ref
from tests.test_host_logic import vnet_drops
from tests.test_oracle_golden import checksum

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("exact", [True, False])
def test_vnet_matches_reference_fixture(golden, exact):
    """Student logits of the reference's own VNet (tests/golden/vnet_uamt.pt): exact (3xTF32) mode, and the production TF32
    path (3-D halo-block tcgen05 forward / data gradient, row-ring weight gradient) with TF32 tolerances."""
    g = golden("vnet_uamt.pt")
    torch.manual_seed(g["seed"])
    net = vnet_mod.VNet(1, 2, has_dropout=False, exact=exact)         # the fixture ran with Dropout3d p = 0
    if abs(checksum(net.state_dict()) - g["init_ck"][0]) > 1e-6 * g["init_ck"][0]:
        pytest.skip("torch RNG stream differs from the fixture's")
    x, y, _ = O.vnet_fixture_inputs(g["gen_seed"], g["B"], g["labeled_bs"], g["P"])
    net = net.cuda().train()
    logits = net(x.cuda())
    if exact:
        torch.testing.assert_close(logits[:, :, ::4, ::4, ::4].cpu(), g["logits_sub"], rtol=2e-3, atol=3e-4)
    else:
        err = float((logits[:, :, ::4, ::4, ::4].cpu() - g["logits_sub"]).abs().max())
        assert err <= 5e-2 * float(g["logits_sub"].abs().max()), err
    assert abs(float(logits.detach().abs().mean()) - g["logits_stat"][1]) < (1e-3 if exact else 1e-2) * g["logits_stat"][1]
    Lb = g["labeled_bs"]
    loss, ce, dice = O.supervised_loss(logits[:Lb], y[:Lb].cuda(), 2)
    torch.testing.assert_close(ce.cpu(), g["ce"], rtol=1e-3 if exact else 1e-2, atol=1e-5)
    torch.testing.assert_close(dice.cpu(), g["dice"], rtol=1e-3 if exact else 1e-2, atol=1e-5)
    loss.backward()                                                    # backward runs (values are checked in the step test)
    assert all(torch.isfinite(p.grad).all() for p in net.parameters())


def test_mc_uncertainty_loss_matches_reference_op():
    g = torch.Generator().manual_seed(29)
    B, Lb, C, S, T = 4, 2, 2, 12 * 12 * 12, 8
    U = B - Lb
    logits, teacher = torch.randn(B, C, S, generator=g), torch.randn(U, C, S, generator=g)
    y = torch.randint(0, C, (B, S), generator=g)
    psum_d, psum_r = torch.empty(U, C, S, device=DEV), torch.empty(U, C, S)
    for i in range(T // 2):
        mc = torch.randn(2 * U, C, S, generator=g) * 3
        ops.mc_softmax_accumulate(mc.to(DEV), psum_d, 2, U, C, S, False, i == 0)
        ref.mc_softmax_accumulate(mc, psum_r, 2, U, C, S, False, i == 0)
    torch.testing.assert_close(psum_d.cpu(), psum_r, rtol=1e-5, atol=1e-6)
    w, thr = torch.tensor([0.05]), torch.tensor([0.55])
    ws = torch.empty(ops.ssl_loss_workspace_bytes(B, S) // 4 + 4, device=DEV)
    lb, lb_r = torch.zeros(32, device=DEV), torch.zeros(32)
    ops.ssl_loss_fwd(logits.to(DEV), teacher.to(DEV), y.to(DEV), False, B, Lb, C, S, w.to(DEV), lb, ws, psum_d, float(T), thr.to(DEV))
    ref.ssl_loss_fwd(logits, teacher, y, False, B, Lb, C, S, w, lb_r, None, psum_r, float(T), thr)
    torch.testing.assert_close(lb[:4].cpu(), lb_r[:4], rtol=1e-5, atol=1e-6)
    assert float(lb[2]) > 0
    dl, dl_r = torch.empty(B, S, C, device=DEV), torch.empty(B, S, C)
    ops.ssl_loss_bwd(logits.to(DEV), teacher.to(DEV), y.to(DEV), False, B, Lb, C, S, w.to(DEV), lb, 1.0, dl, True, psum_d, float(T), thr.to(DEV))
    ref.ssl_loss_bwd(logits, teacher, y, False, B, Lb, C, S, w, lb_r, 1.0, dl_r, True, psum_r, float(T), thr)
    torch.testing.assert_close(dl.cpu(), dl_r, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("use_graph", [False, True])
def test_uamt_step_matches_oracle(use_graph):
    torch.manual_seed(31)
    student = vnet_mod.VNet(1, 2, has_dropout=True, seed=301, exact=True)
    teacher = vnet_mod.VNet(1, 2, has_dropout=True, seed=402, exact=True)
    for p in teacher.parameters():
        p.detach_()
    s_sd = {k: v.clone() for k, v in student.state_dict().items()}
    t_sd = {k: v.clone() for k, v in teacher.state_dict().items()}
    student, teacher = student.cuda(), teacher.cuda()
    B, Lb, P, T = 4, 2, 32, 8
    U = B - Lb
    tr = MeanTeacherTrainer(student, teacher, batch_size=B, labeled_bs=Lb, patch_size=(P, P, P), num_classes=2,
                            start_iter=2000, consistency_gate_iters=0, uncertainty_T=T, noise_seed=777, use_cuda_graph=use_graph)
    tr.lr = O.poly_lr(0.01, 1999, 30000)
    bufs = {k: torch.zeros_like(s_sd[k]) for k in O.param_keys(s_sd)}
    g = torch.Generator().manual_seed(9)
    t_epoch = 0
    for step in range(2):
        it = 2000 + step
        x = torch.randn(B, 1, P, P, P, generator=g)
        y = torch.randint(0, 2, (B, P, P, P), generator=g)
        ce, dice, cons, total = tr.step(x.pin_memory(), y.pin_memory(), read_loss=True)
        noises, tdrops = [], []
        for k in range(1 + T // 2):
            t_epoch += 1
            nb = U if k == 0 else 2 * U
            noises.append(torch.from_numpy(philox.clamp_noise(777 + t_epoch, 1000, nb * P ** 3)).reshape(nb, 1, P, P, P))
            tdrops.append(vnet_drops(402 + t_epoch, nb))
        r = O.uamt3d_step(s_sd, t_sd, bufs, x, y, noises, it, labeled_bs=Lb, T=T,
                          student_drops=vnet_drops(301 + step + 1, B), teacher_drops=tdrops)
        assert abs(total - float(r["loss"])) < 1e-3 * abs(float(r["loss"])) + 1e-5, (total, float(r["loss"]))
        assert abs(ce - float(r["ce"])) < 1e-3 and abs(dice - float(r["dice"])) < 1e-3
        assert abs(cons - float(r["cons"])) < 5e-2 * float(r["cons"]) + 1e-5, (cons, float(r["cons"]))   # mask flips at the threshold
        sd_now, td_now = student.state_dict(), teacher.state_dict()
        for k in s_sd:
            if s_sd[k].dtype.is_floating_point:
                torch.testing.assert_close(sd_now[k].cpu(), s_sd[k], rtol=1e-2, atol=5e-5, msg=lambda m, k=k: f"student {k}: {m}")
                torch.testing.assert_close(td_now[k].cpu(), t_sd[k], rtol=1e-2, atol=5e-5, msg=lambda m, k=k: f"teacher {k}: {m}")


def test_full_size_uamt_step_properties():
    """BASELINE config 4: bs4 (2 labeled), 96^3, T = 8 -- runs, is finite, graph replay == eager schedule."""
    B, Lb, P = 4, 2, 96
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 1, P, P, P, generator=g).pin_memory()
    low = torch.randint(0, 2, (B, P // 8, P // 8, P // 8), generator=g)
    y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3).long().pin_memory()

    def run(graph):
        torch.manual_seed(9)
        s = net_factory_3d("vnet", 1, 2, seed=1)
        t = net_factory_3d("vnet", 1, 2, seed=2)
        tr = MeanTeacherTrainer(s, t, batch_size=B, labeled_bs=Lb, patch_size=(P, P, P), num_classes=2, start_iter=3000,
                                consistency_gate_iters=0, uncertainty_T=8, use_cuda_graph=graph)
        losses = [tr.step(x, y, read_loss=True) for _ in range(2)]
        return losses, tr.flat.data.clone(), tr.ema_flat.data.clone()

    l_e, p_e, e_e = run(False)
    l_g, p_g, e_g = run(True)
    assert all(all(v == v and abs(v) < 1e3 for v in l) for l in l_e)
    assert l_e == l_g and torch.equal(p_e, p_g) and torch.equal(e_e, e_g)
    assert torch.isfinite(p_e).all() and torch.isfinite(e_e).all()
    assert net_factory_3d("does_not_exist") is None
    sd = net_factory_3d("vnet", 1, 2).state_dict()
    assert len(sd) == 205 and sum(1 for k in sd if "num_batches" not in k and "running" not in k) == 118


@pytest.mark.parametrize("exact", [False, True])
def test_full_size_vnet_fwd_bwd_matches_cpu_oracle(exact):
    """BASELINE config-4 volume size (96^3, two volumes) through the kernels the bench times -- exact=False: 3-D halo-block
    tcgen05 forward / data gradient (64-byte rows at the 16-channel level), row-ring weight gradient, GEMM-formulated
    2x2x2 stride-2 convs -- against the CPU oracle (torch fp32) with the same Philox dropout masks: logits within
    5e-2 of the largest logit (1e-3 exact), loss 1e-2 (1e-4).  Weight gradients: the 30 BatchNorm backward passes of this
    network amplify round-off ~300x from the head to the first block -- measured on the B200 (this test, -s): the `exact`
    (3xTF32) schedule goes from 1e-5 at `out_conv` to 1.1e-2 at `block_one` against the fp32 oracle, the TF32 path from
    1e-4 to 0.34 with the SAME profile (a constant ~30x = the ratio of the per-product errors), cosine 0.94 at the
    worst tensor.  Gates: exact 2e-2 / cosine 0.9999 (pins every index of the full-size schedule), TF32 0.4 / cosine 0.93
    (catches any mis-indexed or mis-signed tile; cuDNN's TF32 -- the reference's default numerics -- sits in the same class)."""
    torch.manual_seed(17)
    net = vnet_mod.VNet(1, 2, has_dropout=True, seed=501, exact=exact)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda().train()
    B, P = 2, 96
    g = torch.Generator().manual_seed(12)
    x = torch.randn(B, 1, P, P, P, generator=g)
    low = torch.randint(0, 2, (B, P // 8, P // 8, P // 8), generator=g)
    y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3).long()
    logits = net(x.cuda())                                             # the autograd bridge bumps the RNG epoch once
    loss, ce, dice = O.supervised_loss(logits, y.cuda(), 2)
    loss.backward()
    keys = O.param_keys(sd0)
    leaf = {k: (v.clone().requires_grad_(True) if k in keys else v.clone()) for k, v in sd0.items()}
    d5, d9 = vnet_drops(501 + 1, B)
    ref = O.vnet_forward(leaf, x, True, d5, d9, update_running=False)
    rloss, _, _ = O.supervised_loss(ref, y, 2)
    rloss.backward()
    big = float(ref.abs().max())
    err = float((logits.detach().cpu() - ref.detach()).abs().max())
    print(f"logits: max |err| {err:.3e} of max |logit| {big:.3f}")
    assert err <= (1e-3 if exact else 5e-2) * big, (err, big)
    assert abs(float(loss) - float(rloss)) <= (1e-4 if exact else 1e-2) * abs(float(rloss))
    named = dict(net.named_parameters())
    worst, bad = 0.0, []
    for k in keys:
        if not k.endswith("weight") or leaf[k].grad.dim() < 2:          # conv weights (biases in front of BatchNorm: ~0 gradient)
            continue
        a, b = named[k].grad.detach().cpu().flatten(), leaf[k].grad.flatten()
        rel = float((a - b).norm() / (b.norm() + 1e-20))
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-20))
        worst = max(worst, rel)
        print(f"{k:40s} rel {rel:.4f} cos {cos:.5f}")
        bad = bad + [(k, rel, cos)] if not (rel < (2e-2 if exact else 0.4) and cos > (0.9999 if exact else 0.93)) else bad
    print("worst relative weight-gradient error", worst)
    assert not bad, bad
