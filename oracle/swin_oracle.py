"""TEST INFRASTRUCTURE -- CPU restatement of the reference's Swin-UNet forward and Cross-Teaching step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.

Restates, as plain functional torch over a state_dict (autograd supplies the backward):
  code/networks/vision_transformer.py:48-52                      (1 -> 3 channel repeat)
  code/networks/swin_transformer_unet_skip_expand_decoder_sys.py  (cited per function as `sys:LINES`)
  code/train_cross_teaching_between_cnn_transformer_2D.py:221-259 (the training iteration)

Third-party dependency absent from /root/reference: timm (unpinned in the reference; `timm.models.layers.DropPath`,
`trunc_normal_`, `to_2tuple`).  DropPath is restated from its published algorithm: in training, each SAMPLE of the
branch is kept with probability 1 - p and scaled by 1 / (1 - p); identity in eval mode.  The keep decisions are injected
(`drop_keep`) so the CUDA path's counter-based draws can be replayed here.

Parity pin: tests/golden/swin_ct.pt is produced by tests/golden/make_golden.py running the reference's own
SwinTransformerSys / SwinUnet classes (timm shimmed as described) -- tests/test_oracle_golden.py checks this
restatement against it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ssl_oracle as O


def window_partition(x, ws):
    """sys:28-40: [B, H, W, C] -> [B * nW, ws, ws, C]"""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    """sys:43-57"""
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(ws):
    """sys:91-104"""
    ch = torch.arange(ws)
    coords = torch.stack(torch.meshgrid([ch, ch], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_attn_mask(H, W, ws, shift):
    """sys:212-232: -100 (not -inf) between tokens of different cyclic-shift regions"""
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


def window_attention(x, sd, pre, ws, heads, mask):
    """WindowAttention.forward, sys:115-150.  x: [nW*B, N, C]"""
    B_, N, C = x.shape
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (C // heads) ** -0.5                                             # qk_scale None/False -> head_dim ** -0.5 (:83)
    attn = q @ k.transpose(-2, -1)
    index = relative_position_index(ws)
    bias = sd[pre + "relative_position_bias_table"][index.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(x, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def _layer_norm(x, sd, pre):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "weight"], sd[pre + "bias"], 1e-5)


def _drop_path(x, keep, p):
    """timm DropPath (training): x * keep[b] / (1 - p); keep None => identity (p == 0 or eval)."""
    if keep is None or p == 0.0:
        return x
    return x * (keep.view(-1, 1, 1) / (1.0 - p))


def swin_block(x, sd, pre, res, heads, window, shift, p_drop, keeps):
    """SwinTransformerBlock.forward, sys:239-288.  keeps: (keep_attn, keep_mlp) per-sample 0/1 vectors or None."""
    H, W = res
    if min(res) <= window:                                                   # sys:184-187
        shift, window = 0, min(res)
    B, L, C = x.shape
    shortcut = x
    x = _layer_norm(x, sd, pre + "norm1.").view(B, H, W, C)
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = window_partition(x, window).view(-1, window * window, C)
    mask = shift_attn_mask(H, W, window, shift) if shift > 0 else None
    aw = window_attention(xw, sd, pre + "attn.", window, heads, mask).view(-1, window, window, C)
    x = window_reverse(aw, window, H, W)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    x = x.view(B, H * W, C)
    x = shortcut + _drop_path(x, keeps[0] if keeps else None, p_drop)
    h = F.linear(_layer_norm(x, sd, pre + "norm2."), sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])   # Mlp sys:19-25
    h = F.linear(F.gelu(h), sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + _drop_path(h, keeps[1] if keeps else None, p_drop)


def patch_merging(x, sd, pre, res):
    """sys:325-346"""
    H, W = res
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1).view(B, -1, 4 * C)
    return F.linear(_layer_norm(x, sd, pre + "norm."), sd[pre + "reduction.weight"])


def patch_expand(x, sd, pre, res, p):
    """PatchExpand sys:367-382 (p = 2) and FinalPatchExpand_X4 sys:395-410 (p = 4):
    rearrange 'b h w (p1 p2 c) -> b (h p1) (w p2) c'"""
    H, W = res
    x = F.linear(x, sd[pre + "expand.weight"])
    B, L, C = x.shape
    c = C // (p * p)
    x = x.view(B, H, W, p, p, c).permute(0, 1, 3, 2, 4, 5).reshape(B, H * p * W * p, c)
    return _layer_norm(x, sd, pre + "norm.")


def swin_config(sd, img_size, window_size=7, drop_path_rate=0.2, patch_size=4):
    """Recover (embed_dim, depths, heads) from the state_dict keys; the rest are the yaml-lite constants."""
    E = sd["patch_embed.proj.weight"].shape[0]
    depths = [1 + max(int(k.split(".")[3]) for k in sd if k.startswith(f"layers.{i}.blocks.")) for i in range(4)]
    heads = [sd[f"layers.{i}.blocks.0.attn.relative_position_bias_table"].shape[1] for i in range(4)]
    dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]  # sys:650-651
    return dict(E=E, depths=depths, heads=heads, dpr=dpr, img=img_size, window=window_size, patch=patch_size)


def swin_unet_forward(sd, x, cfg, train=True, drop_keep=None):
    """SwinUnet.forward (vision_transformer.py:48-52) over SwinTransformerSys.forward (sys:788-793).

    sd: state_dict of `SwinUnet.swin_unet` (keys without the 'swin_unet.' prefix); x: [B, 1 or 3, img, img];
    drop_keep: None (no DropPath) or a list, in block construction order (encoder blocks then decoder blocks), of
    (keep_attn, keep_mlp) float vectors [B]."""
    if x.size(1) == 1:
        x = x.repeat(1, 3, 1, 1)
    depths, heads, dpr, window, patch = cfg["depths"], cfg["heads"], cfg["dpr"], cfg["window"], cfg["patch"]
    R = cfg["img"] // patch
    bi = 0

    def blocks(x, pre, n, res, nh, rates):
        nonlocal bi
        for j in range(n):
            keeps = drop_keep[bi] if (drop_keep is not None and train) else None
            x = swin_block(x, sd, f"{pre}.blocks.{j}.", res, nh, window, 0 if j % 2 == 0 else window // 2, rates[j], keeps)
            bi += 1
        return x

    # forward_features sys:742-757
    x = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)   # sys:585
    x = _layer_norm(x, sd, "patch_embed.norm.")
    skips = []
    for i in range(4):
        res = (R // 2 ** i, R // 2 ** i)
        skips.append(x)
        x = blocks(x, f"layers.{i}", depths[i], res, heads[i], dpr[sum(depths[:i]):sum(depths[:i + 1])])
        if i < 3:
            x = patch_merging(x, sd, f"layers.{i}.downsample.", res)
    x = _layer_norm(x, sd, "norm.")
    # forward_up_features sys:762-773
    for inx in range(4):
        j = 3 - inx
        res = (R // 2 ** j, R // 2 ** j)
        if inx == 0:
            x = patch_expand(x, sd, "layers_up.0.", res, 2)
        else:
            x = torch.cat([x, skips[3 - inx]], -1)
            x = F.linear(x, sd[f"concat_back_dim.{inx}.weight"], sd[f"concat_back_dim.{inx}.bias"])
            x = blocks(x, f"layers_up.{inx}", depths[j], res, heads[j], dpr[sum(depths[:j]):sum(depths[:j + 1])])   # sys:706-707
            if inx < 3:
                x = patch_expand(x, sd, f"layers_up.{inx}.upsample.", res, 2)
    x = _layer_norm(x, sd, "norm_up.")
    # up_x4 sys:775-786
    B = x.shape[0]
    x = patch_expand(x, sd, "up.", (R, R), 4)
    x = x.view(B, 4 * R, 4 * R, -1).permute(0, 3, 1, 2)
    return F.conv2d(x, sd["output.weight"])


def cross_teaching_loss(out1, out2, labels, labeled_bs, n_classes, w):
    """code/train_cross_teaching_between_cnn_transformer_2D.py:229-249.  Returns (loss, model1_loss, model2_loss, parts)."""
    Lb = labeled_bs
    soft1, soft2 = torch.softmax(out1, dim=1), torch.softmax(out2, dim=1)
    l1, ce1, dice1 = O.supervised_loss(out1[:Lb], labels[:Lb], n_classes)          # :232-235
    l2, ce2, dice2 = O.supervised_loss(out2[:Lb], labels[:Lb], n_classes)
    pseudo1 = torch.argmax(soft1[Lb:].detach(), dim=1, keepdim=False)              # :237-240
    pseudo2 = torch.argmax(soft2[Lb:].detach(), dim=1, keepdim=False)
    ps1 = O.dice_loss_multiclass(soft1[Lb:], pseudo2.unsqueeze(1), n_classes)      # :242-245
    ps2 = O.dice_loss_multiclass(soft2[Lb:], pseudo1.unsqueeze(1), n_classes)
    m1, m2 = l1 + w * ps1, l2 + w * ps2                                            # :247-248
    return m1 + m2, m1, m2, dict(ce1=ce1, dice1=dice1, ps1=ps1, ce2=ce2, dice2=dice2, ps2=ps2)


def ct2d_step(sd1, sd2, bufs1, bufs2, images, labels, iter_num, cfg, *, labeled_bs, n_classes=4, base_lr=0.01,
              max_iterations=30000, consistency=0.1, consistency_rampup=200.0, lr=None, masks1=None, drop_keep=None):
    """One iteration of code/train_cross_teaching_between_cnn_transformer_2D.py:221-262 on CPU.

    sd1: UNet state_dict, sd2: SwinUnet.swin_unet state_dict (both updated in place); bufs*: SGD momentum buffers.
    `lr` is the rate in effect for this step: the reference increments iter_num BEFORE recomputing it (:257-259), so step k
    (0-based) runs with poly_lr(k) -- base_lr for the first step."""
    k1, k2 = O.param_keys(sd1), [k for k, v in sd2.items() if v.dtype.is_floating_point and not k.endswith("attn_mask")]
    leaf1 = {k: (v.clone().requires_grad_(True) if k in k1 else v) for k, v in sd1.items()}
    leaf2 = {k: (v.clone().requires_grad_(True) if k in k2 else v) for k, v in sd2.items()}
    out1 = O.unet_forward(leaf1, images, True, masks1, update_running=True)         # :224
    out2 = swin_unet_forward(leaf2, images, cfg, True, drop_keep)                   # :227
    w = O.consistency_weight(iter_num, consistency, consistency_rampup)             # :229-230 (iter_num // 150 inside)
    loss, m1, m2, parts = cross_teaching_loss(out1, out2, labels, labeled_bs, n_classes, w)
    g = torch.autograd.grad(loss, [leaf1[k] for k in k1] + [leaf2[k] for k in k2])  # :252-255
    g1, g2 = g[:len(k1)], g[len(k1):]
    if lr is None:
        lr = O.poly_lr(base_lr, iter_num, max_iterations)
    with torch.no_grad():
        O.sgd_momentum_step([sd1[k] for k in k1], g1, [bufs1[k] for k in k1], lr)   # :255
        O.sgd_momentum_step([sd2[k] for k in k2], g2, [bufs2[k] for k in k2], lr)   # :256
        for k in sd1:
            if k not in k1:
                sd1[k] = leaf1[k]
    return dict(loss=loss.detach(), model1_loss=m1.detach(), model2_loss=m2.detach(), w=w, lr=lr, logits1=out1.detach(),
                logits2=out2.detach(), grads1=dict(zip(k1, g1)), grads2=dict(zip(k2, g2)),
                **{k: v.detach() for k, v in parts.items()})
