"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package (cv_ssl_mis_b200).

Plain PyTorch fp32 (CPU) restatement of the reference hot path of ziyangwang007/CV-SSL-MIS, written
functionally over a state_dict so it can be driven with injected dropout masks / noise.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this file.

Parity is PINNED: `tests/golden/make_golden.py` imports the reference's own modules from
/root/reference/code (networks/unet.py, networks/vnet.py, utils/losses.py, utils/ramps.py) and stores
their outputs on seeded inputs in tests/golden/*.pt; tests/test_oracle_golden.py checks every function
here against those fixtures.  The trainer step (`mt2d_step`) restates lines that cannot be imported
(the train_*.py files need tensorboardX/medpy/h5py and hard-code .cuda()); it is pinned against the same
lines driven through the reference's own UNet/DiceLoss/ramps modules by make_golden.py.

All paths below are relative to the reference root.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

UNET_FT = [16, 32, 64, 128, 256]          # code/networks/unet.py:312
UNET_DROPOUT = [0.05, 0.1, 0.2, 0.3, 0.5]  # code/networks/unet.py:313


# --------------------------------------------------------------------------- ramps / schedules
def sigmoid_rampup(current, rampup_length):
    """code/utils/ramps.py:20-27"""
    if rampup_length == 0:
        return 1.0
    current = np.clip(current, 0.0, rampup_length)
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def consistency_weight(iter_num, consistency=0.1, consistency_rampup=200.0):
    """code/train_mean_teacher_2D.py:119-121,223 (epoch argument is iter_num // 150)"""
    return consistency * sigmoid_rampup(iter_num // 150, consistency_rampup)


def poly_lr(base_lr, iter_num, max_iterations):
    """code/train_mean_teacher_2D.py:234 -- value installed AFTER step `iter_num`, i.e. used by step iter_num+1"""
    return base_lr * (1.0 - iter_num / max_iterations) ** 0.9


def ema_alpha(global_step, ema_decay=0.99):
    """code/train_mean_teacher_2D.py:126"""
    return min(1 - 1 / (global_step + 1), ema_decay)


# --------------------------------------------------------------------------- UNet (code/networks/unet.py)
def _bn_train(x, sd, prefix, update_running, momentum=0.1, eps=1e-5):
    rm = sd[prefix + ".running_mean"] if update_running else None
    rv = sd[prefix + ".running_var"] if update_running else None
    if update_running:
        if prefix + ".num_batches_tracked" in sd:
            sd[prefix + ".num_batches_tracked"] += 1
        return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, momentum, eps)
    return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, momentum, eps)


def _bn(x, sd, prefix, train, update_running):
    if train:
        return _bn_train(x, sd, prefix, update_running)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.1, 1e-5)


def _unet_conv_block(x, sd, prefix, p, mask, train, update_running):
    """ConvBlock, code/networks/unet.py:31-47: conv3x3-BN-LeakyReLU-Dropout(p)-conv3x3-BN-LeakyReLU"""
    x = F.conv2d(x, sd[prefix + ".conv_conv.0.weight"], sd[prefix + ".conv_conv.0.bias"], padding=1)
    x = F.leaky_relu(_bn(x, sd, prefix + ".conv_conv.1", train, update_running), 0.01)
    if train and p > 0 and mask is not None:
        x = x * mask / (1.0 - p)          # nn.Dropout with an injected keep-mask
    x = F.conv2d(x, sd[prefix + ".conv_conv.4.weight"], sd[prefix + ".conv_conv.4.bias"], padding=1)
    x = F.leaky_relu(_bn(x, sd, prefix + ".conv_conv.5", train, update_running), 0.01)
    return x


def unet_forward(sd, x, train=True, masks=None, update_running=False, return_features=False):
    """UNet.forward, code/networks/unet.py:318-321 (Encoder :110-116, Decoder :141-153, UpBlock :81-86).

    sd: state_dict with the reference keys; masks: optional list of 5 keep-masks (NCHW 0/1 floats) for the
    encoder dropouts (None => no dropout, i.e. p treated as 0)."""
    feats = []
    names = ["encoder.in_conv"] + [f"encoder.down{i}.maxpool_conv.1" for i in range(1, 5)]
    h = x
    for i, name in enumerate(names):
        if i > 0:
            h = F.max_pool2d(h, 2)
        m = masks[i] if masks is not None else None
        h = _unet_conv_block(h, sd, name, UNET_DROPOUT[i], m, train, update_running)
        feats.append(h)
    h = feats[4]
    for j in range(1, 5):
        pre = f"decoder.up{j}"
        h = F.conv2d(h, sd[pre + ".conv1x1.weight"], sd[pre + ".conv1x1.bias"])
        h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=True)
        h = torch.cat([feats[4 - j], h], dim=1)
        h = _unet_conv_block(h, sd, pre + ".conv", 0.0, None, train, update_running)
    out = F.conv2d(h, sd["decoder.out_conv.weight"], sd["decoder.out_conv.bias"], padding=1)
    return (out, feats) if return_features else out


# --------------------------------------------------------------------------- losses (code/utils/losses.py)
def dice_loss_multiclass(probs, target, n_classes):
    """DiceLoss.forward with softmax=False, weight=None: code/utils/losses.py:170-201.
    probs [B,C,...] probabilities, target [B,1,...] integer labels."""
    loss = 0.0
    smooth = 1e-5
    for i in range(n_classes):
        t = (target == i).float()[:, 0]
        s = probs[:, i]
        intersect = torch.sum(s * t)
        y_sum = torch.sum(t * t)
        z_sum = torch.sum(s * s)
        loss = loss + (1 - (2 * intersect + smooth) / (z_sum + y_sum + smooth))
    return loss / n_classes


def softmax_mse_loss(input_logits, target_logits):
    """code/utils/losses.py:74-91 (sigmoid=False): element-wise, unreduced"""
    return (F.softmax(input_logits, dim=1) - F.softmax(target_logits, dim=1)) ** 2


def supervised_loss(logits, labels, n_classes):
    """0.5 * (Dice + CE): code/train_mean_teacher_2D.py:218-222 / code/train_fully_supervised_2D.py:111-114"""
    ce = F.cross_entropy(logits, labels.long())
    dice = dice_loss_multiclass(torch.softmax(logits, dim=1), labels.unsqueeze(1), n_classes)
    return 0.5 * (dice + ce), ce, dice


def mt_loss(student_logits, teacher_logits, labels, labeled_bs, n_classes, w_cons):
    """Loss of one Mean-Teacher step, code/train_mean_teacher_2D.py:212-229 (w_cons already includes the
    iter<1000 gate: pass 0.0 there)."""
    sup, ce, dice = supervised_loss(student_logits[:labeled_bs], labels[:labeled_bs], n_classes)
    soft = torch.softmax(student_logits, dim=1)
    if teacher_logits is not None:
        cons = torch.mean((soft[labeled_bs:] - torch.softmax(teacher_logits, dim=1)) ** 2)
    else:
        cons = torch.zeros(())
    return sup + w_cons * cons, ce, dice, cons


# --------------------------------------------------------------------------- optimizer / EMA
def sgd_momentum_step(params, grads, bufs, lr, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD(momentum, weight_decay), no dampening/nesterov: code/train_mean_teacher_2D.py:189-190.
    bufs start at zero (== torch's 'first step clones the gradient')."""
    for p, g, b in zip(params, grads, bufs):
        d = g + weight_decay * p
        b.mul_(momentum).add_(d)
        p.add_(b, alpha=-lr)


def ema_update(ema_params, params, alpha):
    """update_ema_variables, code/train_mean_teacher_2D.py:124-128"""
    for e, p in zip(ema_params, params):
        e.mul_(alpha).add_(p, alpha=1 - alpha)


def clamp_noise(like, generator=None, sigma=0.1, clip=0.2):
    """code/train_mean_teacher_2D.py:208-209"""
    return torch.clamp(torch.randn(like.shape, generator=generator) * sigma, -clip, clip)


PARAM_SUFFIXES = (".weight", ".bias")


def param_keys(sd):
    """state_dict keys that are parameters (not BN buffers), in registration order."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


def mt2d_step(student_sd, teacher_sd, bufs, images, labels, noise, iter_num, *, labeled_bs, n_classes=4, base_lr=0.01,
              max_iterations=30000, ema_decay=0.99, consistency=0.1, consistency_rampup=200.0, lr=None,
              student_masks=None, teacher_masks=None):
    """One iteration of code/train_mean_teacher_2D.py:204-236 on CPU.

    student_sd / teacher_sd: state_dicts (updated in place, BN running stats included);
    bufs: dict key -> momentum buffer (updated in place); noise: the clamp(randn*0.1) tensor to add to the
    unlabeled images (injected so the CUDA path can use the same draw); lr: learning rate in effect for this
    step (default: what the reference would have installed after the previous step).
    Returns dict(loss, ce, dice, cons, logits, teacher_logits, grads)."""
    keys = param_keys(student_sd)
    leaf = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in student_sd.items()}
    ema_inputs = images[labeled_bs:] + noise                                        # :206-210
    outputs = unet_forward(leaf, images, True, student_masks, update_running=True)   # :212
    with torch.no_grad():
        ema_output = unet_forward(teacher_sd, ema_inputs, True, teacher_masks, update_running=True)   # :214-216 (train mode)
    w = consistency_weight(iter_num, consistency, consistency_rampup)               # :223
    w_eff = 0.0 if iter_num < 1000 else w                                           # :224-228
    loss, ce, dice, cons = mt_loss(outputs, ema_output, labels, labeled_bs, n_classes, w_eff)
    if iter_num < 1000:
        cons = torch.zeros(())                                                      # :224-225 consistency_loss = 0.0
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys])                      # :230-231
    if lr is None:
        lr = base_lr if iter_num == 0 else poly_lr(base_lr, iter_num - 1, max_iterations)
    with torch.no_grad():
        params = [student_sd[k] for k in keys]
        sgd_momentum_step(params, grads, [bufs[k] for k in keys], lr)               # :232
        ema_update([teacher_sd[k] for k in keys], params, ema_alpha(iter_num, ema_decay))   # :233
        for k in student_sd:                                                        # BN running stats moved in `leaf`
            if k not in keys:
                student_sd[k] = leaf[k]
    return dict(loss=loss.detach(), ce=ce.detach(), dice=dice.detach(), cons=cons.detach(), logits=outputs.detach(),
                teacher_logits=ema_output, grads=dict(zip(keys, grads)), lr=lr, w=w_eff)


# --------------------------------------------------------------------------- VNet (code/networks/vnet.py)
VNET_STAGES = [1, 2, 3, 3, 3, 3, 3, 2, 1]
VNET_NAMES = ["one", "two", "three", "four", "five", "six", "seven", "eight", "nine"]


def _vnet_stage(h, sd, name, n, train, upd):
    """ConvBlock, code/networks/vnet.py:5-31: n x [conv3x3x3 -> BatchNorm3d -> ReLU]"""
    for i in range(n):
        h = F.conv3d(h, sd[f"{name}.conv.{3 * i}.weight"], sd[f"{name}.conv.{3 * i}.bias"], padding=1)
        h = F.relu(_bn(h, sd, f"{name}.conv.{3 * i + 1}", train, upd))
    return h


def vnet_forward(sd, x, train=True, drop5=None, drop9=None, update_running=False):
    """VNet.forward with normalization='batchnorm', code/networks/vnet.py:180-239.
    drop5 / drop9: optional [B, C] keep-masks of the two Dropout3d(0.5) (None => has_dropout False / eval)."""
    feats = []
    h = x
    for s in range(5):                                                     # encoder :180-200
        h = _vnet_stage(h, sd, f"block_{VNET_NAMES[s]}", VNET_STAGES[s], train, update_running)
        if s == 4 and drop5 is not None:
            h = h * drop5[:, :, None, None, None] / 0.5                    # Dropout3d: whole channels (:195-196)
        feats.append(h)
        if s < 4:
            name = f"block_{VNET_NAMES[s]}_dw"                             # DownsamplingConvBlock :67-91
            h = F.conv3d(h, sd[name + ".conv.0.weight"], sd[name + ".conv.0.bias"], stride=2)
            h = F.relu(_bn(h, sd, name + ".conv.1", train, update_running))
    for k in range(4):                                                     # decoder :202-228
        name = f"block_{VNET_NAMES[4 + k]}_up"                             # UpsamplingDeconvBlock :94-118
        h = F.conv_transpose3d(h, sd[name + ".conv.0.weight"], sd[name + ".conv.0.bias"], stride=2)
        h = F.relu(_bn(h, sd, name + ".conv.1", train, update_running))
        h = h + feats[3 - k]                                               # additive skips :210,214,218,222
        h = _vnet_stage(h, sd, f"block_{VNET_NAMES[5 + k]}", VNET_STAGES[5 + k], train, update_running)
    if drop9 is not None:
        h = h * drop9[:, :, None, None, None] / 0.5                        # :225-226
    return F.conv3d(h, sd["out_conv.weight"], sd["out_conv.bias"])


def uamt_loss(student_logits, teacher_logits, mc_logits, labels, labeled_bs, n_classes, w_cons, threshold, T=8):
    """code/train_uncertainty_aware_mean_teacher_3D.py:161-181.
    mc_logits: list of T//2 tensors [2U, C, ...] (the stochastic teacher passes of the twice-repeated unlabeled batch)."""
    U = student_logits.shape[0] - labeled_bs
    preds = torch.cat(mc_logits, 0)                                        # [stride*T, C, ...] with stride = U  (:152-160)
    preds = torch.softmax(preds, dim=1)
    preds = preds.reshape(T, U, *preds.shape[1:]).mean(0)                  # :162-163
    uncertainty = -1.0 * torch.sum(preds * torch.log(preds + 1e-6), dim=1, keepdim=True)   # :164-165
    sup, ce, dice = supervised_loss(student_logits[:labeled_bs], labels[:labeled_bs], n_classes)
    dist = softmax_mse_loss(student_logits[labeled_bs:], teacher_logits)   # :173-174
    mask = (uncertainty < threshold).float()                               # :177
    cons = torch.sum(mask * dist) / (2 * torch.sum(mask) + 1e-16)          # :178-179
    return sup + w_cons * cons, ce, dice, cons, mask


def uamt3d_step(student_sd, teacher_sd, bufs, images, labels, noises, iter_num, *, labeled_bs, n_classes=2, base_lr=0.01,
                max_iterations=30000, ema_decay=0.99, consistency=0.1, consistency_rampup=200.0, lr=None, T=8,
                student_drops=None, teacher_drops=None, forward=vnet_forward):
    """One iteration of code/train_uncertainty_aware_mean_teacher_3D.py:137-189 on CPU.
    noises: [noise for ema_inputs (U samples)] + T//2 noises for the repeated batch (2U samples each);
    student_drops: (drop5, drop9) or None; teacher_drops: list of 1 + T//2 (drop5, drop9) pairs or None."""
    keys = param_keys(student_sd)
    leaf = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in student_sd.items()}
    unlabeled = images[labeled_bs:]
    sdrop = student_drops or (None, None)
    outputs = forward(leaf, images, True, sdrop[0], sdrop[1], update_running=True)                 # :145
    mc = []
    with torch.no_grad():
        td = teacher_drops[0] if teacher_drops else (None, None)
        ema_output = forward(teacher_sd, unlabeled + noises[0], True, td[0], td[1], update_running=True)   # :147-148
        volume_batch_r = unlabeled.repeat(2, *([1] * (images.dim() - 1)))                        # :151
        for i in range(T // 2):                                                                    # :155-160
            td = teacher_drops[1 + i] if teacher_drops else (None, None)
            mc.append(forward(teacher_sd, volume_batch_r + noises[1 + i], True, td[0], td[1], update_running=True))
    w = consistency_weight(iter_num, consistency, consistency_rampup)                              # :172
    threshold = (0.75 + 0.25 * sigmoid_rampup(iter_num, max_iterations)) * np.log(2)               # :175-176
    loss, ce, dice, cons, mask = uamt_loss(outputs, ema_output, mc, labels, labeled_bs, n_classes, w, threshold, T)
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys])
    if lr is None:
        lr = base_lr if iter_num == 0 else poly_lr(base_lr, iter_num - 1, max_iterations)
    with torch.no_grad():
        params = [student_sd[k] for k in keys]
        sgd_momentum_step(params, grads, [bufs[k] for k in keys], lr)                              # :183-185
        ema_update([teacher_sd[k] for k in keys], params, ema_alpha(iter_num, ema_decay))          # :186
    return dict(loss=loss.detach(), ce=ce.detach(), dice=dice.detach(), cons=cons.detach(), logits=outputs.detach(),
                teacher_logits=ema_output, mask_frac=float(mask.mean()), grads=dict(zip(keys, grads)), lr=lr, w=w,
                threshold=float(threshold))


def vnet_fixture_inputs(gen_seed, B, Lb, P, T=8):
    """Replays the torch.Generator calls of tests/golden/make_golden.py:vnet_fixture (inputs, labels, noises)."""
    g = torch.Generator().manual_seed(gen_seed)
    x = torch.randn(B, 1, P, P, P, generator=g)
    low = torch.randint(0, 2, (B, P // 8, P // 8, P // 8), generator=g)
    y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3).long()
    U = B - Lb
    noises = [torch.clamp(torch.randn(U, 1, P, P, P, generator=g) * 0.1, -0.2, 0.2)]
    for _ in range(T // 2):
        noises.append(torch.clamp(torch.randn(2 * U, 1, P, P, P, generator=g) * 0.1, -0.2, 0.2))
    return x, y, noises
