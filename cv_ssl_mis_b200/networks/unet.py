"""2D UNet of the reference (code/networks/unet.py:304-321, PyMIC flavour) on the B200 kernels.

The module tree only *owns* parameters and buffers -- same registration order and state_dict keys as the
reference (`encoder.in_conv.conv_conv.0.weight`, `decoder.up1.conv1x1.weight`, ...), same default
initialisers because the containers are the stock torch layers -- while `forward` runs `UNetPlan`, a
fixed schedule of C-ABI calls over channels-last buffers:

    encoder level i (code/networks/unet.py:110-116):  [maxpool2] -> conv3x3+BN+LeakyReLU+Dropout(p_i) -> conv3x3+BN+LeakyReLU
    decoder level  (code/networks/unet.py:81-86):     conv1x1 -> bilinear x2 (align_corners) -> cat[skip, up] (virtual)
                                                      -> conv3x3+BN+LeakyReLU -> conv3x3+BN+LeakyReLU
    head (code/networks/unet.py:138):                 conv3x3 -> logits, stored NCHW for the caller

Quirks kept: `bilinear=False` in the params dict is ignored by the reference (UpBlock's default wins), the
skip is concatenated first, decoder dropout is p=0, LeakyReLU slope is the torch default 0.01.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ._engine import ConvLayer, FlatParams, PackTable, Runtime

FT_CHNS = [16, 32, 64, 128, 256]
DROPOUT = [0.05, 0.1, 0.2, 0.3, 0.5]
LRELU_SLOPE = 0.01


def _conv_block(cin, cout, p):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU(), nn.Dropout(p),
                         nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU())


class ConvBlock(nn.Module):
    def __init__(self, cin, cout, p):
        super().__init__()
        self.conv_conv = _conv_block(cin, cout, p)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, p):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), ConvBlock(cin, cout, p))


class UpBlock(nn.Module):
    def __init__(self, c_low, c_skip, cout):
        super().__init__()
        self.conv1x1 = nn.Conv2d(c_low, c_skip, kernel_size=1)
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = ConvBlock(c_skip * 2, cout, 0.0)


class Encoder(nn.Module):
    def __init__(self, in_chns):
        super().__init__()
        self.in_conv = ConvBlock(in_chns, FT_CHNS[0], DROPOUT[0])
        for i in range(1, 5):
            setattr(self, f"down{i}", DownBlock(FT_CHNS[i - 1], FT_CHNS[i], DROPOUT[i]))


class Decoder(nn.Module):
    def __init__(self, class_num):
        super().__init__()
        for j in range(1, 5):
            setattr(self, f"up{j}", UpBlock(FT_CHNS[5 - j], FT_CHNS[4 - j], FT_CHNS[4 - j]))
        self.out_conv = nn.Conv2d(FT_CHNS[0], class_num, kernel_size=3, padding=1)


class UNetPlan:
    """All buffers and the launch schedule of one UNet instance for one input geometry (B, H, W)."""

    def __init__(self, net: "UNet", rt: Runtime, B, H, W, need_grad):
        assert H % 16 == 0 and W % 16 == 0, "UNet needs H, W divisible by 16"
        self.net, self.rt, self.B, self.H, self.W, self.need_grad = net, rt, B, H, W, need_grad
        dev = rt.device
        enc_blocks = [net.encoder.in_conv] + [getattr(net.encoder, f"down{i}").maxpool_conv[1] for i in range(1, 5)]
        self.enc = []            # (conv_a, conv_b) per level
        self.pooled = []         # pooled input of level i+1
        self.pooled_g = []
        stream = 0
        h, w, cin = H, W, net.in_chns
        for i, blk in enumerate(enc_blocks):
            seq = blk.conv_conv
            la = ConvLayer(seq[0], seq[1], LRELU_SLOPE, DROPOUT[i], 1, stream, name=f"enc{i}a").plan(rt, B, 1, h, w, cin, 0, need_grad)
            lb = ConvLayer(seq[4], seq[5], LRELU_SLOPE, name=f"enc{i}b").plan(rt, B, 1, h, w, FT_CHNS[i], 0, need_grad)
            stream += 1
            self.enc.append((la, lb))
            if i < 4:
                self.pooled.append(torch.empty((B * (h // 2) * (w // 2), FT_CHNS[i]), dtype=torch.float32, device=dev))
                self.pooled_g.append(torch.empty_like(self.pooled[-1]) if need_grad else None)
                h, w, cin = h // 2, w // 2, FT_CHNS[i]
        self.dec = []            # (conv1x1, up buffer, up grad, conv_a, conv_b) per level
        for j in range(1, 5):
            ub = getattr(net.decoder, f"up{j}")
            c_low, c_skip = FT_CHNS[5 - j], FT_CHNS[4 - j]
            l1 = ConvLayer(ub.conv1x1, name=f"dec{j}c1").plan(rt, B, 1, h, w, c_low, 0, need_grad)
            h, w = h * 2, w * 2
            up = torch.empty((B * h * w, c_skip), dtype=torch.float32, device=dev)
            up_g = torch.empty_like(up) if need_grad else None
            seq = ub.conv.conv_conv
            la = ConvLayer(seq[0], seq[1], LRELU_SLOPE, name=f"dec{j}a").plan(rt, B, 1, h, w, c_skip, c_skip, need_grad)
            lb = ConvLayer(seq[4], seq[5], LRELU_SLOPE, name=f"dec{j}b").plan(rt, B, 1, h, w, c_skip, 0, need_grad)
            self.dec.append((l1, up, up_g, la, lb))
        self.head = ConvLayer(net.decoder.out_conv, out_nchw=True, name="out").plan(rt, B, 1, H, W, FT_CHNS[0], 0, need_grad)
        self.layers = [l for pair in self.enc for l in pair] + [l for d in self.dec for l in (d[0], d[3], d[4])] + [self.head]
        for l in self.layers:                 # profiling tags: gradient-carrying (student) vs forward-only (teacher) plan
            l.name = ("S." if need_grad else "T.") + l.name
        rt.alloc_scratch()
        self.packer = PackTable(self.layers, need_grad, dev)
        self.in_flight = False

    @property
    def logits(self):            # [B, C, H*W] view of the head output (NCHW)
        return self.head.y

    @property
    def g_logits(self):          # channels-last d(loss)/d(logits), [B*H*W, C]
        return self.head.g

    def forward(self, x, train=True, repack=True):
        """x: [B,1,H,W] (== channels-last for one channel) or [B*H*W, in_chns] channels-last."""
        rt, B = self.rt, self.B
        if repack:                       # False: the weights have not changed since this plan's previous forward
            self.packer.run()
        self.x_in = x
        src = x
        h, w = self.H, self.W
        for i, (la, lb) in enumerate(self.enc):
            a = la.forward(rt, src, None, train)
            xi = lb.forward(rt, a, None, train)
            if i < 4:
                ops.maxpool2_fwd(xi, self.pooled[i], B, h, w, FT_CHNS[i])
                src = self.pooled[i]
                h, w = h // 2, w // 2
        cur = self.enc[4][1].a
        for j, (l1, up, up_g, la, lb) in enumerate(self.dec):
            c1 = l1.forward(rt, cur, None, train)
            ops.upsample2x_fwd(c1, up, B, h, w, l1.cout)
            h, w = h * 2, w * 2
            skip = self.enc[3 - j][1].a
            a = la.forward(rt, skip, up, train)
            cur = lb.forward(rt, a, None, train)
        self.head.forward(rt, cur, None, train)
        return self.head.y

    def decoder_grad_offset(self):
        """First element of the decoder's parameters in the flat parameter / gradient buffer (encoder first, then
        decoder: registration order of code/networks/unet.py:309-316)."""
        flat = self.net._flat
        first = next(p for p in self.net.decoder.parameters())
        return flat.offsets[[id(q) for q in flat.params].index(id(first))]

    def backward(self, dlogits_nhwc=None, after_decoder=None):
        """d(loss)/d(logits), channels-last [B*H*W, C]; if None it is already in self.head.g.
        after_decoder(): called once every launch that writes a decoder parameter gradient has been issued (main and
        side stream) -- the data-parallel trainer starts the all-reduce of that bucket there."""
        rt, B = self.rt, self.B
        if dlogits_nhwc is not None and dlogits_nhwc.data_ptr() != self.head.g.data_ptr():
            self.head.g.copy_(dlogits_nhwc.view_as(self.head.g))
        h, w = self.H, self.W
        last = self.dec[3][4]
        self.head.backward(rt, last.a, None, last.g)
        for j in range(3, -1, -1):
            l1, up, up_g, la, lb = self.dec[j]
            skip_l = self.enc[3 - j][1]
            lb.backward(rt, la.a, None, la.g)
            # the skip gradient is the first contribution to the encoder feature's gradient (overwrite)
            la.backward(rt, skip_l.a, up, skip_l.g, up_g)
            h, w = h // 2, w // 2
            ops.upsample2x_bwd(up_g, l1.g, B, h, w, l1.cout)
            src_l = self.dec[j - 1][4] if j > 0 else self.enc[4][1]
            l1.backward(rt, src_l.a, None, src_l.g)
        if after_decoder is not None:
            after_decoder()
        for i in range(4, -1, -1):
            la, lb = self.enc[i]
            lb.backward(rt, la.a, None, la.g)
            if i > 0:
                la.backward(rt, self.pooled[i - 1], None, self.pooled_g[i - 1])
                prev = self.enc[i - 1][1]
                ops.maxpool2_bwd(prev.a, self.pooled_g[i - 1], prev.g, B, h * 2, w * 2, FT_CHNS[i - 1], accumulate=True)
                h, w = h * 2, w * 2
            else:
                la.backward(rt, self.x_in, None, None)
        rt.join_side()


class _UNetFn(torch.autograd.Function):
    """Autograd bridge for drop-in use (`loss.backward()` in the reference trainers)."""

    @staticmethod
    def forward(ctx, net, need_grad, x, *params):
        B, _, H, W = x.shape
        plan = net._get_plan(B, H, W, need_grad)
        xin = x.contiguous().float()
        if net.in_chns != 1:
            xin_cl = torch.empty((B * H * W, net.in_chns), dtype=torch.float32, device=x.device)
            ops.nchw_to_nhwc(xin, xin_cl, B, net.in_chns, H * W)
            xin = xin_cl
        net._bump_seed()
        out = plan.forward(xin, train=net.training)
        ctx.plan = plan
        plan.in_flight = need_grad
        return out.view(B, net.class_num, H, W).clone()

    @staticmethod
    def backward(ctx, grad_out):
        plan = ctx.plan
        if not plan.in_flight:
            raise RuntimeError("UNet backward called twice or after its buffers were reused")
        net = plan.net
        B, C, H, W = grad_out.shape
        ops.nchw_to_nhwc(grad_out.contiguous().float(), plan.head.g, B, C, H * W)
        # parameter .grad tensors are scratch here: save, run, hand the results to autograd, restore
        saved = [p.grad for p in net._flat.params]
        tmp = torch.zeros_like(net._flat.grad)
        for p, o in zip(net._flat.params, net._flat.offsets):
            p.grad = tmp[o:o + p.numel()].view(p.shape)
        plan.backward(None)
        grads = [p.grad for p in net._flat.params]
        for p, g in zip(net._flat.params, saved):
            p.grad = g
        plan.in_flight = False
        return (None, None, None, *grads)


class UNet(nn.Module):
    """Drop-in for networks.unet.UNet(in_chns, class_num) (code/networks/unet.py:304)."""

    _instances = 0

    def __init__(self, in_chns, class_num, seed=None, exact=False):
        super().__init__()
        self.in_chns, self.class_num = in_chns, class_num
        self.encoder = Encoder(in_chns)
        self.decoder = Decoder(class_num)
        if seed is None:                      # distinct dropout streams for student / teacher instances
            seed = 1337 + 1000003 * UNet._instances
        UNet._instances += 1
        self._seed, self._exact = seed, exact
        self._flat = None
        self._rt = None
        self._plans = {}

    # -- device placement: parameters are re-homed into one flat buffer the first time the net is on the GPU
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat, self._plans = None, {}
        return out

    def materialize(self):
        dev = next(self.parameters()).device      # CPU tensors are rejected by the first ops.* call (no CPU path)
        if self._flat is None:
            self._flat = FlatParams(self, dev)
            self._rt = Runtime(dev, self._seed, self._exact)
            self._plans = {}
        return self._flat

    def _get_plan(self, B, H, W, need_grad) -> UNetPlan:
        self.materialize()
        key = (B, H, W, need_grad)
        pool = self._plans.setdefault(key, [])
        for pl in pool:
            if not pl.in_flight:
                return pl
        if len(pool) >= 4:
            pool[0].in_flight = False
            return pool[0]
        pl = UNetPlan(self, self._rt, B, H, W, need_grad)
        pool.append(pl)
        return pl

    def _bump_seed(self):
        self._rt.seed_off += 1

    def forward(self, x):
        self.materialize()
        # autograd turns grad mode off inside Function.forward, so decide here whether backward buffers are needed
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._flat.params)
        return _UNetFn.apply(self, need_grad, x, *self._flat.params)
