"""Generate the golden fixtures in tests/golden/ by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports networks/unet.py, utils/losses.py, utils/ramps.py from /root/reference/code (read-only) and stores
small seeded input/output vectors.  The trainer loop cannot be imported (tensorboardX/medpy/h5py missing,
hard-coded .cuda()), so the Mean-Teacher fixture drives the reference's UNet / DiceLoss / ramps /
torch.optim.SGD through the loop body of code/train_mean_teacher_2D.py:204-236, quoted line by line below.
Weights are not stored (7 MB): both the reference UNet and our parameter containers draw them from
torch.manual_seed(seed) in the same order; a checksum is stored and this script asserts they agree.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/code"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from networks.unet import UNet as RefUNet          # noqa: E402
from utils import losses as ref_losses              # noqa: E402
from utils import ramps as ref_ramps                # noqa: E402

from cv_ssl_mis_b200.networks.unet import UNet as OurUNet   # noqa: E402  (parameter containers only; no kernels run)

torch.set_num_threads(4)
torch.backends.mkldnn.enabled = True


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sd.items() if v.dtype.is_floating_point))


def blocky_labels(gen, B, H, W, ncls, dtype=torch.uint8):
    low = torch.randint(0, ncls, (B, H // 8, W // 8), generator=gen)
    return low.repeat_interleave(8, 1).repeat_interleave(8, 2).to(dtype)


def no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0


def unet_fixture():
    seed = 1234
    torch.manual_seed(seed)
    ref = RefUNet(in_chns=1, class_num=4)
    torch.manual_seed(seed)
    ours = OurUNet(1, 4)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys()), "state_dict key schema differs from the reference"
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ours[k]), k
    assert [tuple(p.shape) for p in ref.parameters()] == [tuple(p.shape) for p in ours.parameters()]
    ck0 = checksum(sd_ref)          # before the train-mode forward moves the BN running stats
    no_dropout(ref)
    g = torch.Generator().manual_seed(99)
    x = torch.rand(2, 1, 32, 32, generator=g)
    y = blocky_labels(g, 2, 32, 32, 4)
    ref.train()
    logits = ref(x)
    soft = torch.softmax(logits, dim=1)
    ce = torch.nn.CrossEntropyLoss()(logits, y.long())
    dice = ref_losses.DiceLoss(4)(soft, y.unsqueeze(1))
    loss = 0.5 * (dice + ce)
    loss.backward()
    names = [n for n, _ in ref.named_parameters()]
    grad_norm = {n: float(p.grad.norm()) for n, p in ref.named_parameters()}
    grad_head = {n: p.grad.flatten()[:8].clone() for n, p in ref.named_parameters()}
    sd_after = ref.state_dict()
    running = {k: sd_after[k].clone() for k in sd_after if "running" in k and ("in_conv" in k or "up4" in k)}
    ref.eval()
    with torch.no_grad():
        logits_eval = ref(x)
    torch.save(dict(seed=seed, checksum=ck0, keys=list(sd_ref.keys()), param_names=names, x=x, y=y,
                    logits=logits.detach(), logits_eval=logits_eval, ce=ce.detach(), dice=dice.detach(),
                    loss=loss.detach(), grad_norm=grad_norm, grad_head=grad_head, running=running),
               os.path.join(HERE, "unet_small.pt"))
    print("unet_small: loss", float(loss), "checksum", checksum(sd_ref))


def losses_fixture():
    g = torch.Generator().manual_seed(7)
    out = {}
    for name, (B, C, shape, dt) in {"2d": (3, 4, (16, 16), torch.uint8), "3d": (2, 2, (8, 8, 8), torch.int64)}.items():
        logits = torch.randn(B, C, *shape, generator=g, requires_grad=True)
        tlogits = torch.randn(B, C, *shape, generator=g)
        y = torch.randint(0, C, (B, *shape), generator=g).to(dt)
        soft = torch.softmax(logits, dim=1)
        dice = ref_losses.DiceLoss(C)(soft, y.unsqueeze(1))
        ce = torch.nn.CrossEntropyLoss()(logits, y.long())
        mse = ref_losses.softmax_mse_loss(logits, tlogits)
        total = 0.5 * (dice + ce) + 0.37 * mse.mean()
        (grad,) = torch.autograd.grad(total, logits)
        out[name] = dict(logits=logits.detach(), teacher=tlogits, y=y, dice=dice.detach(), ce=ce.detach(),
                         mse=mse.detach(), total=total.detach(), grad=grad, w=0.37)
    torch.save(out, os.path.join(HERE, "losses.pt"))
    print("losses: ", {k: float(v["total"]) for k, v in out.items()})


def losses_dropin_fixture():
    """The stand-alone entry points of code/utils/losses.py driven the way the reference trainers drive them, plus
    `update_ema_variables` (code/train_mean_teacher_2D.py:124-128, restated: the script itself needs tensorboardX/medpy
    at import) -- checked by tests/test_losses_dropin_gpu.py against cv_ssl_mis_b200/utils/losses.py."""
    g = torch.Generator().manual_seed(11)
    out = {}
    for name, (B, C, shape, dt) in {"2d": (3, 4, (12, 20), torch.uint8), "3d": (2, 2, (6, 8, 10), torch.int64)}.items():
        logits = torch.randn(B, C, *shape, generator=g, requires_grad=True)
        tlogits = torch.randn(B, C, *shape, generator=g)
        y = torch.randint(0, C, (B, *shape), generator=g).to(dt)
        weight = [0.5 + 0.25 * i for i in range(C)]
        dl = ref_losses.DiceLoss(C)
        dice_sm = dl(logits, y.unsqueeze(1), weight=weight, softmax=True)                       # logits in, softmax inside
        (g_dice_sm,) = torch.autograd.grad(1.7 * dice_sm, logits)
        probs = torch.softmax(logits.detach(), dim=1).requires_grad_(True)
        dice_p = dl(probs, y.unsqueeze(1))                                                      # probabilities in (:214-215)
        (g_dice_p,) = torch.autograd.grad(dice_p, probs)
        mse = ref_losses.softmax_mse_loss(logits, tlogits)                                      # element-wise
        up = torch.rand(mse.shape, generator=g)
        (g_mse,) = torch.autograd.grad((mse * up).sum(), logits)
        with __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            kl = ref_losses.softmax_kl_loss(logits, tlogits)                                    # scalar, reduction='mean'
        (g_kl,) = torch.autograd.grad(0.3 * kl, logits)
        out[name] = dict(logits=logits.detach(), teacher=tlogits, y=y, weight=weight, dice_sm=dice_sm.detach(), g_dice_sm=g_dice_sm,
                         probs=probs.detach(), dice_p=dice_p.detach(), g_dice_p=g_dice_p, mse=mse.detach(), up=up, g_mse=g_mse,
                         kl=kl.detach(), g_kl=g_kl)
    # update_ema_variables, three consecutive global steps on a small parameter list
    student = [torch.randn(5, 3, generator=g), torch.randn(7, generator=g)]
    teacher = [torch.randn(5, 3, generator=g), torch.randn(7, generator=g)]
    ema = dict(student=[t.clone() for t in student], teacher0=[t.clone() for t in teacher], alpha=0.99, steps=[0, 1, 250], teacher=[])
    for step in ema["steps"]:
        alpha = min(1 - 1 / (step + 1), 0.99)                                                   # :126
        for ema_param, param in zip(teacher, student):
            ema_param.mul_(alpha).add_(param, alpha=1 - alpha)                                  # :127-128
        ema["teacher"].append([t.clone() for t in teacher])
    out["ema"] = ema
    torch.save(out, os.path.join(HERE, "losses_dropin.pt"))
    print("losses_dropin:", {k: (float(v["dice_sm"]), float(v["kl"])) for k, v in out.items() if k != "ema"})


def ramps_fixture():
    pts = [(0, 200), (1, 200), (6, 200.0), (50, 200.0), (199, 200), (200, 200), (500, 200), (3, 0), (10, 40.0)]
    vals = [ref_ramps.sigmoid_rampup(c, l) for c, l in pts]
    torch.save(dict(points=pts, values=vals), os.path.join(HERE, "ramps.pt"))
    print("ramps:", vals[:4])


def update_ema_variables(model, ema_model, alpha, global_step):
    # code/train_mean_teacher_2D.py:124-128 (deprecated add_(scalar, tensor) overload spelled in its modern form)
    alpha = min(1 - 1 / (global_step + 1), alpha)
    for ema_param, param in zip(ema_model.parameters(), model.parameters()):
        ema_param.data.mul_(alpha).add_(param.data, alpha=1 - alpha)


def mt_step_fixture():
    """code/train_mean_teacher_2D.py:137-238 on a tiny batch, two iterations straddling the iter<1000 gate."""
    seed = 4321
    torch.manual_seed(seed)
    model = RefUNet(in_chns=1, class_num=4)          # create_model()           :137-144
    ema_model = RefUNet(in_chns=1, class_num=4)      # create_model(ema=True)
    for p in ema_model.parameters():
        p.detach_()
    no_dropout(model)
    no_dropout(ema_model)
    model.train()                                    # :184 (ema_model is never put in eval mode)
    base_lr, max_iterations, labeled_bs, ema_decay = 0.01, 30000, 2, 0.99
    consistency, consistency_rampup = 0.1, 200.0
    optimizer = torch.optim.SGD(model.parameters(), lr=base_lr, momentum=0.9, weight_decay=0.0001)   # :189-190
    ce_loss = torch.nn.CrossEntropyLoss()
    dice_loss = ref_losses.DiceLoss(4)
    g = torch.Generator().manual_seed(5)
    init_ck = (checksum(model.state_dict()), checksum(ema_model.state_dict()))
    # the reference installs lr after each step; emulate having just finished iteration 998
    iter_num = 999
    lr_ = base_lr * (1.0 - (iter_num - 1) / max_iterations) ** 0.9
    for pg in optimizer.param_groups:
        pg["lr"] = lr_
    steps = []
    for _ in range(3):
        volume_batch = torch.rand(4, 1, 32, 32, generator=g)
        label_batch = blocky_labels(g, 4, 32, 32, 4)
        unlabeled_volume_batch = volume_batch[labeled_bs:]                                     # :206
        noise = torch.clamp(torch.randn(unlabeled_volume_batch.shape, generator=g) * 0.1, -0.2, 0.2)   # :208-209
        ema_inputs = unlabeled_volume_batch + noise                                             # :210
        outputs = model(volume_batch)                                                           # :212
        outputs_soft = torch.softmax(outputs, dim=1)
        with torch.no_grad():
            ema_output = ema_model(ema_inputs)                                                  # :215
            ema_output_soft = torch.softmax(ema_output, dim=1)
        loss_ce = ce_loss(outputs[:labeled_bs], label_batch[:][:labeled_bs].long())             # :218-219
        loss_dice = dice_loss(outputs_soft[:labeled_bs], label_batch[:labeled_bs].unsqueeze(1))   # :220-221
        supervised_loss = 0.5 * (loss_dice + loss_ce)
        consistency_weight = consistency * ref_ramps.sigmoid_rampup(iter_num // 150, consistency_rampup)   # :223
        if iter_num < 1000:
            consistency_loss = 0.0
        else:
            consistency_loss = torch.mean((outputs_soft[labeled_bs:] - ema_output_soft) ** 2)   # :227-228
        loss = supervised_loss + consistency_weight * consistency_loss
        lr_used = optimizer.param_groups[0]["lr"]
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        update_ema_variables(model, ema_model, ema_decay, iter_num)                             # :233
        lr_ = base_lr * (1.0 - iter_num / max_iterations) ** 0.9                                # :234
        for param_group in optimizer.param_groups:
            param_group["lr"] = lr_
        steps.append(dict(iter_num=iter_num, x=volume_batch, y=label_batch, noise=noise, logits=outputs.detach(),
                          teacher_logits=ema_output, loss=loss.detach(), ce=loss_ce.detach(), dice=loss_dice.detach(),
                          cons=torch.as_tensor(float(consistency_loss)), w=consistency_weight, lr_used=lr_used,
                          student_ck=checksum(model.state_dict()), teacher_ck=checksum(ema_model.state_dict()),
                          w_out=model.decoder.out_conv.weight.detach().clone(),
                          t_out=ema_model.decoder.out_conv.weight.detach().clone(),
                          w_in=model.encoder.in_conv.conv_conv[0].weight.detach().clone()))
        iter_num = iter_num + 1                                                                 # :238
    torch.save(dict(seed=seed, init_ck=init_ck, labeled_bs=labeled_bs, steps=steps), os.path.join(HERE, "mt_step.pt"))
    print("mt_step: losses", [float(s["loss"]) for s in steps], "cons", [float(s["cons"]) for s in steps])


def vnet_fixture():
    """VNet (batchnorm, has_dropout) forward/backward + one UAMT iteration restated from
    code/train_uncertainty_aware_mean_teacher_3D.py:137-189 over the reference's own VNet/DiceLoss/softmax_mse/ramps."""
    from networks.vnet import VNet as RefVNet
    from cv_ssl_mis_b200.networks.vnet import VNet as OurVNet
    seed = 2468
    torch.manual_seed(seed)
    model = RefVNet(n_channels=1, n_classes=2, normalization='batchnorm', has_dropout=True)
    ema_model = RefVNet(n_channels=1, n_classes=2, normalization='batchnorm', has_dropout=True)
    torch.manual_seed(seed)
    ours = OurVNet(n_channels=1, n_classes=2, normalization='batchnorm', has_dropout=True)
    sd_ref, sd_ours = model.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys()), "VNet state_dict key schema differs from the reference"
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ours[k]), k
    assert [tuple(p.shape) for p in model.parameters()] == [tuple(p.shape) for p in ours.parameters()]
    for p in ema_model.parameters():
        p.detach_()
    init_ck = (checksum(sd_ref), checksum(ema_model.state_dict()))
    model.dropout.p = 0.0                 # Dropout3d off (masks cannot be shared with torch's RNG); BN stays in train mode
    ema_model.dropout.p = 0.0
    model.train(); ema_model.train()
    g = torch.Generator().manual_seed(11)
    B, Lb, P = 4, 2, 32
    volume_batch = torch.randn(B, 1, P, P, P, generator=g)
    low = torch.randint(0, 2, (B, P // 8, P // 8, P // 8), generator=g)
    label_batch = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3).long()
    base_lr, max_iterations, ema_decay, consistency, consistency_rampup = 0.01, 30000, 0.99, 0.1, 200.0
    optimizer = torch.optim.SGD(model.parameters(), lr=base_lr, momentum=0.9, weight_decay=0.0001)
    ce_loss = torch.nn.CrossEntropyLoss()
    dice_loss = ref_losses.DiceLoss(2)
    iter_num = 3000
    unlabeled_volume_batch = volume_batch[Lb:]                                                    # :139
    noises = [torch.clamp(torch.randn(unlabeled_volume_batch.shape, generator=g) * 0.1, -0.2, 0.2)]   # :141-142
    ema_inputs = unlabeled_volume_batch + noises[0]
    outputs = model(volume_batch)                                                                 # :145
    outputs_soft = torch.softmax(outputs, dim=1)
    with torch.no_grad():
        ema_output = ema_model(ema_inputs)                                                        # :148
    T = 8
    _, _, d, w, h = unlabeled_volume_batch.shape
    volume_batch_r = unlabeled_volume_batch.repeat(2, 1, 1, 1, 1)                                 # :151
    stride = volume_batch_r.shape[0] // 2
    preds = torch.zeros([stride * T, 2, d, w, h])
    for i in range(T // 2):                                                                       # :155-160
        nz = torch.clamp(torch.randn(volume_batch_r.shape, generator=g) * 0.1, -0.2, 0.2)
        noises.append(nz)
        with torch.no_grad():
            preds[2 * stride * i:2 * stride * (i + 1)] = ema_model(volume_batch_r + nz)
    preds = torch.softmax(preds, dim=1)
    preds = preds.reshape(T, stride, 2, d, w, h)
    preds = torch.mean(preds, dim=0)
    uncertainty = -1.0 * torch.sum(preds * torch.log(preds + 1e-6), dim=1, keepdim=True)          # :164-165
    loss_ce = ce_loss(outputs[:Lb], label_batch[:Lb])                                             # :167-168
    loss_dice = dice_loss(outputs_soft[:Lb], label_batch[:Lb].unsqueeze(1))
    supervised_loss = 0.5 * (loss_dice + loss_ce)
    consistency_weight = consistency * ref_ramps.sigmoid_rampup(iter_num // 150, consistency_rampup)   # :172
    consistency_dist = ref_losses.softmax_mse_loss(outputs[Lb:], ema_output)                      # :173-174
    threshold = (0.75 + 0.25 * ref_ramps.sigmoid_rampup(iter_num, max_iterations)) * np.log(2)    # :175-176
    mask = (uncertainty < threshold).float()
    consistency_loss = torch.sum(mask * consistency_dist) / (2 * torch.sum(mask) + 1e-16)         # :178-179
    loss = supervised_loss + consistency_weight * consistency_loss
    optimizer.zero_grad()
    loss.backward()
    grad_norm = {n: float(p.grad.norm()) for n, p in model.named_parameters()}
    optimizer.step()
    update_ema_variables(model, ema_model, ema_decay, iter_num)                                   # :186
    # inputs/noises are regenerated by the tests from the same generator calls (vnet_inputs below); logits are
    # stored subsampled (every 4th voxel) plus their mean/abs-mean to keep the fixture small
    sub = lambda t: t[:, :, ::4, ::4, ::4].clone()
    stat = lambda t: (float(t.mean()), float(t.abs().mean()))
    torch.save(dict(seed=seed, init_ck=init_ck, keys=list(sd_ref.keys()), labeled_bs=Lb, iter_num=iter_num, gen_seed=11,
                    B=B, P=P, logits_sub=sub(outputs.detach()), logits_stat=stat(outputs.detach()),
                    teacher_sub=sub(ema_output), teacher_stat=stat(ema_output),
                    loss=loss.detach(), ce=loss_ce.detach(), dice=loss_dice.detach(), cons=consistency_loss.detach(),
                    mask_frac=float(mask.mean()), threshold=float(threshold), w=consistency_weight, grad_norm=grad_norm,
                    w_out=model.out_conv.weight.detach().clone(), t_out=ema_model.out_conv.weight.detach().clone(),
                    student_ck=checksum(model.state_dict()), teacher_ck=checksum(ema_model.state_dict())),
               os.path.join(HERE, "vnet_uamt.pt"))
    print("vnet_uamt: loss", float(loss), "cons", float(consistency_loss), "mask frac", float(mask.mean()))


def _install_timm_shim(keep_source):
    """timm is not installed here (nor vendored by the reference).  Shim `timm.models.layers` with the three names the
    reference imports (…_sys.py:6): to_2tuple, trunc_normal_ (== torch.nn.init.trunc_normal_, same algorithm) and
    DropPath restated from timm's published drop_path(): per-sample Bernoulli(1 - p) keep, divided by the keep
    probability, identity in eval mode.  The keep draws come from `keep_source(module_index, call_index, batch)` so the
    CUDA path's Philox draws can be replayed through the reference's own block code."""
    import types
    global _TIMM_SHIM
    if _TIMM_SHIM is not None:
        # the reference's Swin module was imported against the first shim's DropPath class: keep that class and only
        # swap its draw source (a second class would never be seen by the already-imported module)
        _TIMM_SHIM.keep_source = staticmethod(keep_source)
        _TIMM_SHIM.count = 0
        return _TIMM_SHIM
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):
        count = 0
        keep_source = None

        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob, self.index, self.calls = drop_prob, DropPath.count, 0
            DropPath.count += 1

        def forward(self, x):
            call, self.calls = self.calls, self.calls + 1
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = DropPath.keep_source(self.index, call, x.shape[0], self.drop_prob)
            return x * (keep / (1.0 - self.drop_prob)).view(-1, *([1] * (x.dim() - 1)))

    DropPath.keep_source = staticmethod(keep_source)
    layers.DropPath = DropPath
    layers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    _TIMM_SHIM = DropPath
    return DropPath


_TIMM_SHIM = None

SWIN_SMALL = dict(img_size=64, patch_size=4, in_chans=3, embed_dim=32, depths=(2, 2, 2, 2), num_heads=(1, 2, 4, 8),
                  window_size=4, mlp_ratio=4.0, qkv_bias=True, drop_path_rate=0.2, patch_norm=True)


def swin_fixture():
    """Swin-UNet forward/backward and one Cross-Teaching iteration (code/train_cross_teaching_between_cnn_transformer_2D.py
    :221-259) driven through the reference's own SwinUnet / UNet / DiceLoss / ramps / torch.optim.SGD.

    Small geometry (64x64 slices, window 4 -> stages 16^2/8^2 shifted, 4^2 and 2^2 unshifted, heads 1/2/4/8 of dim 32) so
    that it runs in seconds; DropPath 0.2 with injected Philox keep draws (seed 4242, stream 3000 + 2 * block + call)."""
    from types import SimpleNamespace as NS
    from oracle import philox
    from cv_ssl_mis_b200.networks.swin_unet import SwinUnet as OurSwin, DROPPATH_STREAM
    dp_seed = 4242 + 1                      # Runtime seed 4242, one seed bump before the forward

    def keep_source(index, call, batch, p):
        w0 = philox.philox4x32_10(dp_seed, DROPPATH_STREAM + 2 * index + (call % 2), np.arange(batch, dtype=np.uint64))[0]
        u = (w0 >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
        return torch.from_numpy((u >= np.float32(p)).astype(np.float32))

    DropPath = _install_timm_shim(keep_source)
    from networks.vision_transformer import SwinUnet as RefSwin
    c = SWIN_SMALL
    config = NS(DATA=NS(IMG_SIZE=c["img_size"]),
                MODEL=NS(DROP_RATE=0.0, DROP_PATH_RATE=c["drop_path_rate"], PRETRAIN_CKPT=None,
                         SWIN=NS(PATCH_SIZE=c["patch_size"], IN_CHANS=c["in_chans"], EMBED_DIM=c["embed_dim"],
                                 DEPTHS=list(c["depths"]), NUM_HEADS=list(c["num_heads"]), WINDOW_SIZE=c["window_size"],
                                 MLP_RATIO=c["mlp_ratio"], QKV_BIAS=True, QK_SCALE=False, APE=False, PATCH_NORM=True)),
                TRAIN=NS(USE_CHECKPOINT=False))
    seed = 1357
    torch.manual_seed(seed)
    model1 = RefUNet(in_chns=1, class_num=4)
    DropPath.count = 0
    model2 = RefSwin(config, img_size=c["img_size"], num_classes=4)
    # DropPath draw ids follow the BLOCK index in construction order (rate-0 blocks hold nn.Identity, not DropPath)
    sysm = model2.swin_unet
    all_blocks = [b for l in sysm.layers for b in l.blocks] + [b for l in list(sysm.layers_up)[1:] for b in l.blocks]
    for bi, blk in enumerate(all_blocks):
        if isinstance(blk.drop_path, DropPath):
            blk.drop_path.index = bi
    torch.manual_seed(seed)
    ours1 = OurUNet(1, 4)
    ours2 = OurSwin(config, img_size=c["img_size"], num_classes=4)
    sd_ref, sd_ours = model2.state_dict(), ours2.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys()), "SwinUnet state_dict key schema differs from the reference"
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ours[k]), k
    assert [tuple(p.shape) for p in model2.parameters()] == [tuple(p.shape) for p in ours2.parameters()]
    for k in model1.state_dict():
        assert torch.equal(model1.state_dict()[k], ours1.state_dict()[k]), k
    # give LayerNorm / bias / rel-pos-bias parameters non-trivial values (their init is 1 / 0 / ~N(0, .02)):
    # seeded perturbation that the tests replay
    g = torch.Generator().manual_seed(97)
    with torch.no_grad():
        for k, p in model2.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif k.endswith("relative_position_bias_table"):
                p.add_(0.5 * torch.randn(p.shape, generator=g))
    init_ck = (checksum(model1.state_dict()), checksum(model2.state_dict()))
    no_dropout(model1)
    model1.train(); model2.train()
    g = torch.Generator().manual_seed(31)
    B, Lb, P = 4, 2, c["img_size"]
    volume_batch = torch.rand(B, 1, P, P, generator=g)
    label_batch = blocky_labels(g, B, P, P, 4)
    base_lr, max_iterations, iter_num = 0.01, 30000, 24000
    args = NS(labeled_bs=Lb, consistency=0.1, consistency_rampup=200.0)
    optimizer1 = torch.optim.SGD(model1.parameters(), lr=base_lr, momentum=0.9, weight_decay=0.0001)     # :200-203
    optimizer2 = torch.optim.SGD(model2.parameters(), lr=base_lr, momentum=0.9, weight_decay=0.0001)
    lr0 = base_lr * (1.0 - iter_num / max_iterations) ** 0.9            # installed by the previous iteration (:258-262)
    for opt in (optimizer1, optimizer2):
        for pg in opt.param_groups:
            pg['lr'] = lr0
    ce_loss = torch.nn.CrossEntropyLoss()
    dice_loss = ref_losses.DiceLoss(4)

    def get_current_consistency_weight(epoch):                                                           # :117-119
        return args.consistency * ref_ramps.sigmoid_rampup(epoch, args.consistency_rampup)

    outputs1 = model1(volume_batch)                                                                      # :224-225
    outputs_soft1 = torch.softmax(outputs1, dim=1)
    outputs2 = model2(volume_batch)                                                                      # :227-228
    outputs_soft2 = torch.softmax(outputs2, dim=1)
    consistency_weight = get_current_consistency_weight(iter_num // 150)                                 # :229-230
    loss1 = 0.5 * (ce_loss(outputs1[:args.labeled_bs], label_batch[:args.labeled_bs].long()) + dice_loss(
        outputs_soft1[:args.labeled_bs], label_batch[:args.labeled_bs].unsqueeze(1)))                    # :232-233
    loss2 = 0.5 * (ce_loss(outputs2[:args.labeled_bs], label_batch[:args.labeled_bs].long()) + dice_loss(
        outputs_soft2[:args.labeled_bs], label_batch[:args.labeled_bs].unsqueeze(1)))                    # :234-235
    pseudo_outputs1 = torch.argmax(outputs_soft1[args.labeled_bs:].detach(), dim=1, keepdim=False)       # :237-240
    pseudo_outputs2 = torch.argmax(outputs_soft2[args.labeled_bs:].detach(), dim=1, keepdim=False)
    pseudo_supervision1 = dice_loss(outputs_soft1[args.labeled_bs:], pseudo_outputs2.unsqueeze(1))       # :242-245
    pseudo_supervision2 = dice_loss(outputs_soft2[args.labeled_bs:], pseudo_outputs1.unsqueeze(1))
    model1_loss = loss1 + consistency_weight * pseudo_supervision1                                       # :247-248
    model2_loss = loss2 + consistency_weight * pseudo_supervision2
    loss = model1_loss + model2_loss                                                                     # :250
    optimizer1.zero_grad()
    optimizer2.zero_grad()
    loss.backward()                                                                                      # :255
    gn1 = {n: float(p.grad.norm()) for n, p in model1.named_parameters()}
    gn2 = {n: float(p.grad.norm()) for n, p in model2.named_parameters()}
    table_grad = model2.swin_unet.layers[0].blocks[1].attn.relative_position_bias_table.grad.clone()
    optimizer1.step()                                                                                    # :257-258
    optimizer2.step()
    stat = lambda t: (float(t.mean()), float(t.abs().mean()))
    torch.save(dict(seed=seed, perturb_seed=97, gen_seed=31, dp_seed=4242, init_ck=init_ck, keys=list(sd_ref.keys()),
                    cfg=c, B=B, labeled_bs=Lb, iter_num=iter_num, lr=lr0, w=consistency_weight,
                    logits1_sub=outputs1.detach()[:, :, ::4, ::4].clone(), logits1_stat=stat(outputs1.detach()),
                    logits2_sub=outputs2.detach()[:, :, ::2, ::2].clone(), logits2_stat=stat(outputs2.detach()),
                    loss=loss.detach(), model1_loss=model1_loss.detach(), model2_loss=model2_loss.detach(),
                    ps1=pseudo_supervision1.detach(), ps2=pseudo_supervision2.detach(), loss1=loss1.detach(),
                    loss2=loss2.detach(), grad_norm1=gn1, grad_norm2=gn2, table_grad=table_grad,
                    out_w2=model2.swin_unet.output.weight.detach().clone(),
                    qkv_w2=model2.swin_unet.layers[1].blocks[1].attn.qkv.weight.detach()[:8].clone(),
                    ck1=checksum(model1.state_dict()), ck2=checksum(model2.state_dict())),
               os.path.join(HERE, "swin_ct.pt"))
    print("swin_ct: loss", float(loss), "m1", float(model1_loss), "m2", float(model2_loss), "w", consistency_weight,
          "ps", float(pseudo_supervision1), float(pseudo_supervision2))


def cps_ict_fixture():
    """One iteration each of Cross Pseudo Supervision (code/train_cross_pseudo_supervision_2D.py:170-213) and of
    Interpolation Consistency Training (code/train_interpolation_consistency_training_2D.py:150-197), driven through the
    reference's UNet / DiceLoss / ramps / torch.optim.SGD, lines quoted below.  Dropout off (its masks cannot be shared
    with torch's RNG); BatchNorm in train mode."""
    ce_loss = torch.nn.CrossEntropyLoss()
    dice_loss = ref_losses.DiceLoss(4)
    base_lr, max_iterations, consistency, consistency_rampup = 0.01, 30000, 0.1, 200.0
    out = {}
    # ---------------------------------------------------------------- CPS
    seed = 8765
    torch.manual_seed(seed)
    model1, model2 = RefUNet(in_chns=1, class_num=4), RefUNet(in_chns=1, class_num=4)          # :118-126 (two create_model())
    no_dropout(model1), no_dropout(model2)
    model1.train(), model2.train()
    init_ck = (checksum(model1.state_dict()), checksum(model2.state_dict()))
    iter_num, labeled_bs = 4000, 2
    lr_ = base_lr * (1.0 - iter_num / max_iterations) ** 0.9                                    # rate installed by iteration 3999 (:203-207)
    optimizer1 = torch.optim.SGD(model1.parameters(), lr=lr_, momentum=0.9, weight_decay=0.0001)
    optimizer2 = torch.optim.SGD(model2.parameters(), lr=lr_, momentum=0.9, weight_decay=0.0001)
    g = torch.Generator().manual_seed(6)
    volume_batch = torch.rand(4, 1, 32, 32, generator=g)
    label_batch = blocky_labels(g, 4, 32, 32, 4)
    outputs1 = model1(volume_batch)                                                             # :176
    outputs_soft1 = torch.softmax(outputs1, dim=1)
    outputs2 = model2(volume_batch)                                                             # :179
    outputs_soft2 = torch.softmax(outputs2, dim=1)
    consistency_weight = consistency * ref_ramps.sigmoid_rampup(iter_num // 150, consistency_rampup)   # :180
    loss1 = 0.5 * (ce_loss(outputs1[:labeled_bs], label_batch[:][:labeled_bs].long()) + dice_loss(
        outputs_soft1[:labeled_bs], label_batch[:labeled_bs].unsqueeze(1)))                    # :182-183
    loss2 = 0.5 * (ce_loss(outputs2[:labeled_bs], label_batch[:][:labeled_bs].long()) + dice_loss(
        outputs_soft2[:labeled_bs], label_batch[:labeled_bs].unsqueeze(1)))                    # :184-185
    pseudo_outputs1 = torch.argmax(outputs_soft1[labeled_bs:].detach(), dim=1, keepdim=False)  # :187
    pseudo_outputs2 = torch.argmax(outputs_soft2[labeled_bs:].detach(), dim=1, keepdim=False)  # :188
    pseudo_supervision1 = ce_loss(outputs1[labeled_bs:], pseudo_outputs2)                       # :190
    pseudo_supervision2 = ce_loss(outputs2[labeled_bs:], pseudo_outputs1)                       # :191
    model1_loss = loss1 + consistency_weight * pseudo_supervision1                              # :193
    model2_loss = loss2 + consistency_weight * pseudo_supervision2                              # :194
    loss = model1_loss + model2_loss                                                            # :196
    optimizer1.zero_grad()
    optimizer2.zero_grad()
    loss.backward()
    optimizer1.step()
    optimizer2.step()
    key = "decoder.up4.conv.conv_conv.0.weight"
    out["cps"] = dict(seed=seed, init_ck=init_ck, iter_num=iter_num, labeled_bs=labeled_bs, lr=lr_, w=consistency_weight,
                      x=volume_batch, y=label_batch, model1_loss=model1_loss.detach(), model2_loss=model2_loss.detach(),
                      ps1=pseudo_supervision1.detach(), ps2=pseudo_supervision2.detach(), loss1=loss1.detach(),
                      loss2=loss2.detach(), key=key, w1=model1.state_dict()[key].clone(), w2=model2.state_dict()[key].clone())
    # ---------------------------------------------------------------- ICT
    seed = 9876
    torch.manual_seed(seed)
    model, ema_model = RefUNet(in_chns=1, class_num=4), RefUNet(in_chns=1, class_num=4)        # :112-121
    for p in ema_model.parameters():
        p.detach_()
    no_dropout(model), no_dropout(ema_model)
    model.train()
    init_ck = (checksum(model.state_dict()), checksum(ema_model.state_dict()))
    iter_num, labeled_bs, ema_decay = 700, 4, 0.99
    lr_ = base_lr * (1.0 - (iter_num - 1) / max_iterations) ** 0.9                              # installed after iteration 699 (:191-193)
    optimizer = torch.optim.SGD(model.parameters(), lr=lr_, momentum=0.9, weight_decay=0.0001)
    g = torch.Generator().manual_seed(8)
    volume_batch = torch.rand(8, 1, 32, 32, generator=g)
    label_batch = blocky_labels(g, 8, 32, 32, 4)
    unlabeled_volume_batch = volume_batch[labeled_bs:]                                          # :152
    labeled_volume_batch = volume_batch[:labeled_bs]
    ict_mix_factors = torch.tensor([0.3, 0.85], dtype=torch.float).view(labeled_bs // 2, 1, 1, 1)   # :156-159 (Beta draws, fixed here)
    unlabeled_volume_batch_0 = unlabeled_volume_batch[0:labeled_bs // 2, ...]                   # :160
    unlabeled_volume_batch_1 = unlabeled_volume_batch[labeled_bs // 2:, ...]                    # :161
    batch_ux_mixed = unlabeled_volume_batch_0 * (1.0 - ict_mix_factors) + unlabeled_volume_batch_1 * ict_mix_factors   # :164-166
    input_volume_batch = torch.cat([labeled_volume_batch, batch_ux_mixed], dim=0)               # :167-168
    outputs = model(input_volume_batch)                                                         # :169
    outputs_soft = torch.softmax(outputs, dim=1)
    with torch.no_grad():
        ema_output_ux0 = torch.softmax(ema_model(unlabeled_volume_batch_0), dim=1)             # :172-173
        ema_output_ux1 = torch.softmax(ema_model(unlabeled_volume_batch_1), dim=1)             # :174-175
        batch_pred_mixed = ema_output_ux0 * (1.0 - ict_mix_factors) + ema_output_ux1 * ict_mix_factors   # :176-177
    loss_ce = ce_loss(outputs[:labeled_bs], label_batch[:labeled_bs][:].long())                 # :179-180
    loss_dice = dice_loss(outputs_soft[:labeled_bs], label_batch[:labeled_bs].unsqueeze(1))     # :181-182
    supervised_loss = 0.5 * (loss_dice + loss_ce)
    consistency_weight = consistency * ref_ramps.sigmoid_rampup(iter_num // 150, consistency_rampup)   # :184
    consistency_loss = torch.mean((outputs_soft[labeled_bs:] - batch_pred_mixed) ** 2)          # :185-186
    loss = supervised_loss + consistency_weight * consistency_loss                              # :187
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    update_ema_variables(model, ema_model, ema_decay, iter_num)                                 # :192
    key = "decoder.up1.conv1x1.weight"
    out["ict"] = dict(seed=seed, init_ck=init_ck, iter_num=iter_num, labeled_bs=labeled_bs, lr=lr_, w=consistency_weight,
                      x=volume_batch, y=label_batch, mix=ict_mix_factors.view(-1).clone(), loss=loss.detach(),
                      ce=loss_ce.detach(), dice=loss_dice.detach(), cons=consistency_loss.detach(), key=key,
                      w_student=model.state_dict()[key].clone(), w_teacher=ema_model.state_dict()[key].clone())
    torch.save(out, os.path.join(HERE, "cps_ict.pt"))
    print("cps: m1", float(model1_loss), "m2", float(model2_loss), "ps", float(pseudo_supervision1), float(pseudo_supervision2),
          "| ict: loss", float(loss), "cons", float(consistency_loss), "w", consistency_weight)


def mt_vit_fixture():
    """One Mean-Teacher iteration over two Swin-UNets (code/train_mean_teacher_ViT.py:147-158, 201-233) through the
    reference's own SwinUnet / DiceLoss / ramps / SGD; DropPath rate 0, the input noise is OUR Philox stream (seed 7,
    epoch 1, stream 1000) so the trainer under test draws the same values."""
    from types import SimpleNamespace as NS
    from oracle import philox
    from cv_ssl_mis_b200.networks.swin_unet import SwinUnet as OurSwin
    _install_timm_shim(lambda *a: None)
    from networks.vision_transformer import SwinUnet as RefSwin
    c = dict(SWIN_SMALL, drop_path_rate=0.0)
    config = NS(DATA=NS(IMG_SIZE=c["img_size"]),
                MODEL=NS(DROP_RATE=0.0, DROP_PATH_RATE=0.0, PRETRAIN_CKPT=None,
                         SWIN=NS(PATCH_SIZE=c["patch_size"], IN_CHANS=c["in_chans"], EMBED_DIM=c["embed_dim"],
                                 DEPTHS=list(c["depths"]), NUM_HEADS=list(c["num_heads"]), WINDOW_SIZE=c["window_size"],
                                 MLP_RATIO=c["mlp_ratio"], QKV_BIAS=True, QK_SCALE=False, APE=False, PATCH_NORM=True)),
                TRAIN=NS(USE_CHECKPOINT=False))
    seed = 2468
    torch.manual_seed(seed)
    model = RefSwin(config, img_size=c["img_size"], num_classes=4)                              # :147-148
    ema_model = RefSwin(config, img_size=c["img_size"], num_classes=4)                          # :151-152
    for param in ema_model.parameters():                                                        # :155-156
        param.detach_()
    torch.manual_seed(seed)
    ours = [OurSwin(config, img_size=c["img_size"], num_classes=4) for _ in range(2)]
    for ref, our in zip((model, ema_model), ours):
        for k, v in ref.state_dict().items():
            assert torch.equal(v, our.state_dict()[k]), k
    init_ck = (checksum(model.state_dict()), checksum(ema_model.state_dict()))
    model.train()
    base_lr, max_iterations, labeled_bs, ema_decay, consistency, consistency_rampup = 0.01, 30000, 2, 0.99, 0.1, 200.0
    iter_num = 1500
    lr_ = base_lr * (1.0 - (iter_num - 1) / max_iterations) ** 0.9
    optimizer = torch.optim.SGD(model.parameters(), lr=lr_, momentum=0.9, weight_decay=0.0001)   # :186-187
    ce_loss = torch.nn.CrossEntropyLoss()
    dice_loss = ref_losses.DiceLoss(4)
    g = torch.Generator().manual_seed(12)
    P = c["img_size"]
    volume_batch = torch.rand(4, 1, P, P, generator=g)
    label_batch = blocky_labels(g, 4, P, P, 4)
    unlabeled_volume_batch = volume_batch[labeled_bs:]                                          # :205
    noise = torch.from_numpy(philox.clamp_noise(7 + 1, 1000, 2 * P * P)).reshape(2, 1, P, P)    # stands for :207-208
    ema_inputs = unlabeled_volume_batch + noise                                                 # :209
    outputs = model(volume_batch)                                                               # :211
    outputs_soft = torch.softmax(outputs, dim=1)
    with torch.no_grad():
        ema_output = ema_model(ema_inputs)                                                      # :214
        ema_output_soft = torch.softmax(ema_output, dim=1)
    loss_ce = ce_loss(outputs[:labeled_bs], label_batch[:][:labeled_bs].long())                 # :217-218
    loss_dice = dice_loss(outputs_soft[:labeled_bs], label_batch[:labeled_bs].unsqueeze(1))     # :219-220
    supervised_loss = 0.5 * (loss_dice + loss_ce)
    consistency_weight = consistency * ref_ramps.sigmoid_rampup(iter_num // 150, consistency_rampup)   # :222
    consistency_loss = torch.mean((outputs_soft[labeled_bs:] - ema_output_soft) ** 2)           # :226-227 (iter >= 1000)
    loss = supervised_loss + consistency_weight * consistency_loss
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    update_ema_variables(model, ema_model, ema_decay, iter_num)                                 # :232
    keys = ["swin_unet.layers.0.blocks.0.mlp.fc1.weight", "swin_unet.layers_up.3.blocks.1.attn.qkv.bias", "swin_unet.output.weight"]
    torch.save(dict(seed=seed, init_ck=init_ck, cfg=c, iter_num=iter_num, labeled_bs=labeled_bs, lr=lr_, w=consistency_weight,
                    x=volume_batch, y=label_batch, loss=loss.detach(), ce=loss_ce.detach(), dice=loss_dice.detach(),
                    cons=consistency_loss.detach(), keys=keys,
                    student={k: model.state_dict()[k].clone() for k in keys},
                    teacher={k: ema_model.state_dict()[k].clone() for k in keys}),
               os.path.join(HERE, "mt_vit.pt"))
    print("mt_vit: loss", float(loss), "cons", float(consistency_loss), "w", consistency_weight)


def uamt2d_fixture():
    """One iteration of code/train_uncertainty_aware_mean_teacher_2D.py:150-196 through the reference's own UNet /
    DiceLoss / softmax_mse_loss / ramps / SGD.  Dropout off; the five noise tensors are OUR Philox stream (seed 7, epochs
    1..5, stream 1000) so the trainer under test draws the same values."""
    import torch.nn.functional as F
    from oracle import philox
    seed = 1470
    torch.manual_seed(seed)
    model, ema_model = RefUNet(in_chns=1, class_num=4), RefUNet(in_chns=1, class_num=4)
    for p in ema_model.parameters():
        p.detach_()
    no_dropout(model), no_dropout(ema_model)
    model.train()
    init_ck = (checksum(model.state_dict()), checksum(ema_model.state_dict()))
    # freshly initialised networks predict near-uniform classes (entropy ~ ln 4 > threshold: an empty mask); sharpen the
    # logits so that the uncertainty mask is partial.  The test applies the same scaling.
    with torch.no_grad():
        for m in (model, ema_model):
            m.decoder.out_conv.weight.mul_(40.0)
    base_lr, max_iterations, labeled_bs, ema_decay, consistency, consistency_rampup = 0.01, 30000, 2, 0.99, 0.1, 200.0
    num_classes, iter_num = 4, 2000
    lr_ = base_lr * (1.0 - (iter_num - 1) / max_iterations) ** 0.9
    optimizer = torch.optim.SGD(model.parameters(), lr=lr_, momentum=0.9, weight_decay=0.0001)
    ce_loss = torch.nn.CrossEntropyLoss()
    dice_loss = ref_losses.DiceLoss(num_classes)
    g = torch.Generator().manual_seed(16)
    P = 32
    volume_batch = torch.rand(4, 1, P, P, generator=g)
    label_batch = blocky_labels(g, 4, P, P, 4)
    noise_of = lambda epoch, n: torch.from_numpy(philox.clamp_noise(7 + epoch, 1000, n * P * P)).reshape(n, 1, P, P)
    unlabeled_volume_batch = volume_batch[labeled_bs:]                                          # :151
    ema_inputs = unlabeled_volume_batch + noise_of(1, 2)                                        # :153-155
    outputs = model(volume_batch)                                                               # :157
    outputs_soft = torch.softmax(outputs, dim=1)
    with torch.no_grad():
        ema_output = ema_model(ema_inputs)                                                      # :160
    T = 8
    _, _, w, h = unlabeled_volume_batch.shape
    volume_batch_r = unlabeled_volume_batch.repeat(2, 1, 1, 1)                                  # :163
    stride = volume_batch_r.shape[0] // 2
    preds = torch.zeros([stride * T, num_classes, w, h])
    for i in range(T // 2):                                                                     # :166-172
        ema_inputs = volume_batch_r + noise_of(2 + i, 4)
        with torch.no_grad():
            preds[2 * stride * i:2 * stride * (i + 1)] = ema_model(ema_inputs)
    preds = F.softmax(preds, dim=1)
    preds = preds.reshape(T, stride, num_classes, w, h)
    preds = torch.mean(preds, dim=0)
    uncertainty = -1.0 * torch.sum(preds * torch.log(preds + 1e-6), dim=1, keepdim=True)       # :176-177
    loss_ce = ce_loss(outputs[:labeled_bs], label_batch[:labeled_bs][:].long())                 # :179-180
    loss_dice = dice_loss(outputs_soft[:labeled_bs], label_batch[:labeled_bs].unsqueeze(1))     # :181-182
    supervised_loss = 0.5 * (loss_dice + loss_ce)
    consistency_weight = consistency * ref_ramps.sigmoid_rampup(iter_num // 150, consistency_rampup)   # :184
    consistency_dist = ref_losses.softmax_mse_loss(outputs[labeled_bs:], ema_output)            # :185-186
    threshold = (0.75 + 0.25 * ref_ramps.sigmoid_rampup(iter_num, max_iterations)) * np.log(2)  # :187-188
    mask = (uncertainty < threshold).float()
    consistency_loss = torch.sum(mask * consistency_dist) / (2 * torch.sum(mask) + 1e-16)       # :190-191
    loss = supervised_loss + consistency_weight * consistency_loss
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    update_ema_variables(model, ema_model, ema_decay, iter_num)
    key = "encoder.down2.maxpool_conv.1.conv_conv.0.weight"
    torch.save(dict(seed=seed, init_ck=init_ck, iter_num=iter_num, labeled_bs=labeled_bs, lr=lr_, w=consistency_weight,
                    threshold=float(threshold), mask_frac=float(mask.mean()), x=volume_batch, y=label_batch,
                    loss=loss.detach(), ce=loss_ce.detach(), dice=loss_dice.detach(), cons=consistency_loss.detach(), key=key,
                    w_student=model.state_dict()[key].clone(), w_teacher=ema_model.state_dict()[key].clone()),
               os.path.join(HERE, "uamt2d.pt"))
    print("uamt2d: loss", float(loss), "cons", float(consistency_loss), "mask fraction", float(mask.mean()))


def unet3d_fixture():
    """The reference's own unet_3D (code/networks/unet_3D.py, built as net_factory_3d does): logits, supervised loss and
    gradient norms on a 2 x 32^3 batch.  The 23 MB of weights are not stored: they are regenerated from a seed
    (oracle/unet3d_oracle.py:fixture_state_dict) and loaded into the reference module with load_state_dict."""
    from networks.unet_3D import unet_3D as RefUNet3D
    from oracle import unet3d_oracle as U3
    from oracle import ssl_oracle as O
    seed, B, P = 1357, 2, 32
    model = RefUNet3D(n_classes=2, in_channels=1)
    sd = U3.fixture_state_dict(seed)
    assert list(model.state_dict().keys()) == list(sd.keys()), "unet_3D state_dict key schema differs from oracle/unet3d_oracle.py"
    model.load_state_dict(sd)
    no_dropout(model)                    # nn.Dropout(p=0.3) off: masks cannot be shared with torch's RNG
    model.train()
    x, y = U3.fixture_inputs(seed + 1, B, P)
    logits = model(x)
    loss, ce, dice = O.supervised_loss(logits, y, 2)
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    torch.save(dict(seed=seed, B=B, P=P, checksum=checksum(sd), logits_sub=logits[:, :, ::2, ::2, ::2].detach().clone(),
                    logits_stat=(float(logits.mean()), float(logits.abs().mean()), float(logits.abs().max())),
                    loss=loss.detach(), ce=ce.detach(), dice=dice.detach(),
                    grad_norm={k: float(g.norm()) for k, g in grads.items()},
                    grad_head={k: g.flatten()[:8].clone() for k, g in grads.items()}),
               os.path.join(HERE, "unet3d.pt"))
    print("unet3d: loss", float(loss), "logits |mean|", float(logits.abs().mean()))


if __name__ == "__main__":
    fixtures = dict(unet3d=unet3d_fixture, uamt2d=uamt2d_fixture, mt_vit=mt_vit_fixture, unet=unet_fixture, losses=losses_fixture,
                    losses_dropin=losses_dropin_fixture, ramps=ramps_fixture, mt_step=mt_step_fixture,
                    vnet=vnet_fixture, swin=swin_fixture, cps_ict=cps_ict_fixture)
    for name in (sys.argv[1:] or list(fixtures)):
        fixtures[name]()
