"""Drop-in for the reference's `utils/losses.py` (code/utils/losses.py:74-113,165-201) on the B200 kernels.

A reference trainer that replaces `from utils import losses` by `from cv_ssl_mis_b200.utils import losses` keeps its code:

    dice_loss = losses.DiceLoss(num_classes)
    loss_dice = dice_loss(outputs_soft[:labeled_bs], label_batch[:labeled_bs].unsqueeze(1))     # train_mean_teacher_2D.py:214-215
    consistency_dist = losses.softmax_mse_loss(outputs[labeled_bs:], ema_output)                # ..._uncertainty_aware_...:180

Each call is a `torch.autograd.Function` over ONE forward and ONE backward kernel of csrc/losses_dropin.cu (so
`loss.backward()` keeps working); the fused trainers (cv_ssl_mis_b200/trainers.py) do not go through here -- they run
CE + Dice + consistency in a single pass.  There is no CPU path: tensors must live on the GPU.
`update_ema_variables` is the helper every reference train script defines (code/train_mean_teacher_2D.py:124-128).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


def _bcs(t):
    """[B, C, *spatial] contiguous fp32 view of the tensor and (B, C, S)."""
    if t.dim() < 3:
        raise ValueError("expected a [B, C, *spatial] tensor")
    t = t.contiguous().float()
    B, C = t.shape[0], t.shape[1]
    return t, B, C, t.numel() // (B * C)


def _labels(target, B, S):
    """[B, 1, *spatial] or [B, *spatial] class indices -> contiguous uint8 / int64 [B, S]."""
    t = target
    if t.dtype not in (torch.uint8, torch.int64):
        t = t.long()
    t = t.contiguous()
    if t.numel() != B * S:
        raise ValueError("predict & target shape do not match")          # losses.py:192
    return t


class _Dice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, target, weight, use_softmax):
        x, B, C, S = _bcs(inputs)
        lab = _labels(target, B, S)
        out = torch.empty(1 + 4 * 8, dtype=torch.float32, device=x.device)
        ws = torch.empty(ops.loss_dropin_workspace_bytes(B, S) // 4 + 4, dtype=torch.float32, device=x.device)
        ops.dice_fwd(x, use_softmax, lab, B, C, S, weight, out, ws)
        ctx.save_for_backward(x, lab, weight if weight is not None else torch.empty(0, device=x.device), out)
        ctx.geom, ctx.use_softmax, ctx.in_shape = (B, C, S), use_softmax, inputs.shape
        return out[0], out[1:1 + C]

    @staticmethod
    def backward(ctx, grad_loss, _grad_classwise):
        x, lab, weight, out = ctx.saved_tensors
        B, C, S = ctx.geom
        dx = torch.empty_like(x)
        g = grad_loss.reshape(1).contiguous().float()
        ops.dice_bwd(x, ctx.use_softmax, lab, B, C, S, weight if weight.numel() else None, out, g, dx)
        return dx.view(ctx.in_shape), None, None, None


class DiceLoss(nn.Module):
    """code/utils/losses.py:165-201: batch-level Dice over one-hot targets, squared terms in the denominator, smooth 1e-5,
    mean over ALL classes (background included)."""

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target, weight=None, softmax=False):
        if inputs.shape[1] != self.n_classes:
            raise AssertionError("predict & target shape do not match")
        w = None
        if weight is not None:
            w = torch.as_tensor(weight, dtype=torch.float32, device=inputs.device).contiguous()
        loss, classwise = _Dice.apply(inputs, target, w, bool(softmax))
        self.class_wise_dice = classwise.detach()          # the reference keeps a python list of .item()s (a sync per class)
        return loss


class _SoftmaxMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_logits, target_logits):
        a, B, C, S = _bcs(input_logits)
        b, _, _, _ = _bcs(target_logits)
        out = torch.empty_like(a)
        ops.softmax_mse_fwd(a, b, B, C, S, out)
        ctx.save_for_backward(a, b)
        ctx.geom = (B, C, S)
        return out.view(input_logits.shape)

    @staticmethod
    def backward(ctx, grad_out):
        a, b = ctx.saved_tensors
        B, C, S = ctx.geom
        da = torch.empty_like(a)
        ops.softmax_mse_bwd(a, b, grad_out.contiguous().float(), B, C, S, da)
        return da.view(grad_out.shape), None          # no gradient to the targets (losses.py:80)


def softmax_mse_loss(input_logits, target_logits, sigmoid=False):
    """code/utils/losses.py:74-91: element-wise (softmax(input) - softmax(target))^2; gradients to the inputs only."""
    assert input_logits.size() == target_logits.size()
    if sigmoid:
        raise NotImplementedError("sigmoid=True is not used by any reference trainer")
    return _SoftmaxMSE.apply(input_logits, target_logits.detach())


class _SoftmaxKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_logits, target_logits):
        a, B, C, S = _bcs(input_logits)
        b, _, _, _ = _bcs(target_logits)
        out = torch.empty(1, dtype=torch.float32, device=a.device)
        ws = torch.empty(ops.loss_dropin_workspace_bytes(B, S) // 4 + 4, dtype=torch.float32, device=a.device)
        ops.softmax_kl_fwd(a, b, B, C, S, out, ws)
        ctx.save_for_backward(a, b)
        ctx.geom, ctx.in_shape = (B, C, S), input_logits.shape
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        a, b = ctx.saved_tensors
        B, C, S = ctx.geom
        da = torch.empty_like(a)
        ops.softmax_kl_bwd(a, b, grad_out.reshape(1).contiguous().float(), B, C, S, da)
        return da.view(ctx.in_shape), None


def softmax_kl_loss(input_logits, target_logits, sigmoid=False):
    """code/utils/losses.py:94-113: F.kl_div(log_softmax(input), softmax(target), reduction='mean') -- a scalar, the mean
    over ALL elements (the deprecated 'mean' reduction, not 'batchmean'); gradients to the inputs only."""
    assert input_logits.size() == target_logits.size()
    if sigmoid:
        raise NotImplementedError("sigmoid=True is not used by any reference trainer")
    return _SoftmaxKL.apply(input_logits, target_logits.detach())


_ema_hp = {}


def update_ema_variables(model, ema_model, alpha, global_step):
    """code/train_mean_teacher_2D.py:124-128: teacher = a * teacher + (1 - a) * student with a = min(1 - 1/(step + 1), alpha),
    parameters only, in registration order.  Networks of this package keep their parameters in one flat buffer each, so
    the whole update is ONE launch of the EMA kernel; foreign modules are updated tensor by tensor with the same kernel."""
    alpha = min(1 - 1 / (global_step + 1), alpha)
    p0 = next(model.parameters())
    hp = _ema_hp.get(p0.device)
    if hp is None:
        hp = _ema_hp[p0.device] = torch.zeros(8, dtype=torch.float32, device=p0.device)
    hp[3:5] = torch.tensor([alpha, 1 - alpha], dtype=torch.float32)
    sflat = getattr(model, "_flat", None) if hasattr(model, "materialize") else None
    tflat = getattr(ema_model, "_flat", None) if hasattr(ema_model, "materialize") else None
    if hasattr(model, "materialize") and hasattr(ema_model, "materialize"):
        sflat, tflat = model.materialize(), ema_model.materialize()
    if sflat is not None and tflat is not None and sflat.padded == tflat.padded:
        ops.ema_update(tflat.data, sflat.data, hp)
        return
    for ema_param, param in zip(ema_model.parameters(), model.parameters()):
        if ema_param.data.is_contiguous() and param.data.is_contiguous():
            ops.ema_update(ema_param.data.view(-1), param.data.view(-1), hp)
        else:
            raise ValueError("update_ema_variables needs contiguous parameters")
