"""3D U-Net of the reference (code/networks/unet_3D.py:20-91, blocks in code/networks/utils.py:99-123,260-276) on the B200
kernels -- the default `--model unet_3D` of the reference's 3-D trainers (net_factory_3d.py:12-13: `unet_3D(n_classes,
in_channels)`, feature_scale 4 -> 16 / 32 / 64 / 128 / 256 filters).

    UnetConv3:    2 x [conv3x3x3 + InstanceNorm3d (no affine, no running statistics) + ReLU]          (utils.py:99-123)
    encoder:      UnetConv3, MaxPool3d(2)  x 4, then `center` + Dropout(0.3)                           (unet_3D.py:73-87)
    UnetUp3_CT:   trilinear x2 (align_corners=False), cat([skip, up]), UnetConv3                       (utils.py:260-276)
    head:         Dropout(0.3) on up1, conv 1x1x1 -> logits, stored NCDHW for the caller               (unet_3D.py:91-94)

InstanceNorm3d is train-mode BatchNorm over a batch of ONE, so (like the UNETR decoder) every sample runs through its own
n = 1 plan of the conv engine and the weight gradients accumulate over the samples.  The module tree owns the parameters
under the reference's state_dict keys (`conv1.conv1.0.weight`, `up_concat4.conv.conv2.0.bias`, `final.weight`, ...).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from ._engine import ConvLayer, FlatParams, PackTable, Runtime
from .unetr import _InstanceNormParams

P_DROP = 0.3                                     # nn.Dropout(p=0.3), element-wise (unet_3D.py:60-61)


class _UnetConv3(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv3d(cin, cout, 3, 1, 1), nn.InstanceNorm3d(cout), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(nn.Conv3d(cout, cout, 3, 1, 1), nn.InstanceNorm3d(cout), nn.ReLU(inplace=True))


class _UnetUp3_CT(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _UnetConv3(cin + cout, cout)
        self.up = nn.Upsample(scale_factor=(2, 2, 2), mode="trilinear")


class _Sample:
    """The convolutional pyramid of ONE sample (n = 1 plans: InstanceNorm = BatchNorm over a batch of one)."""

    def __init__(self, plan, b):
        net, rt, ng, dev = plan.net, plan.rt, plan.need_grad, plan.rt.device
        D, H, W = plan.dims3
        F_ = net.filters
        IN = plan.inorm
        self.b = b
        pre = ("S" if ng else "T") + f"{b}."

        def pair(mod: _UnetConv3, d, h, w, c0, c1, cout, name, drop=None):
            la = ConvLayer(mod.conv1[0], IN(cout), 0.0, dims=3, name=pre + name + "a").plan(rt, 1, d, h, w, c0, c1, ng)
            # element-wise dropout fused into the activation sweep; Philox stream = (sample, site)
            kw = dict(p_drop=P_DROP, drop_mode=1, rng_stream=2 * b + drop) if drop is not None else {}
            lb = ConvLayer(mod.conv2[0], IN(cout), 0.0, dims=3, name=pre + name + "b", **kw).plan(rt, 1, d, h, w, cout, 0, ng)
            return la, lb

        self.enc, self.pooled, self.pooled_g = [], [], []
        d, h, w, cin = D, H, W, net.in_channels
        blocks = [net.conv1, net.conv2, net.conv3, net.conv4, net.center]
        for i, blk in enumerate(blocks):
            self.enc.append(pair(blk, d, h, w, cin, 0, F_[i], f"enc{i}", drop=(0 if i == 4 else None)))
            if i < 4:
                self.pooled.append(torch.empty(((d // 2) * (h // 2) * (w // 2), F_[i]), dtype=torch.float32, device=dev))
                self.pooled_g.append(torch.empty_like(self.pooled[-1]) if ng else None)
                d, h, w, cin = d // 2, h // 2, w // 2, F_[i]
        self.dec = []                                # (up buffer, up grad, conv a, conv b), deepest first
        for j, ub in enumerate((net.up_concat4, net.up_concat3, net.up_concat2, net.up_concat1)):
            c_low, c_skip = F_[4 - j], F_[3 - j]
            d, h, w = 2 * d, 2 * h, 2 * w
            up = torch.empty((d * h * w, c_low), dtype=torch.float32, device=dev)
            up_g = torch.empty_like(up) if ng else None
            la, lb = pair(ub.conv, d, h, w, c_skip, c_low, c_skip, f"dec{j}", drop=(1 if j == 3 else None))
            self.dec.append((up, up_g, la, lb))
        self.head = ConvLayer(net.final, dims=3, out_nchw=True, name=pre + "final").plan(rt, 1, D, H, W, F_[0], 0, ng)
        S = D * H * W
        self.head.y = plan.logits_all[b:b + 1]       # the head writes into / reads from the plan-wide tensors
        self.head.a = self.head.y
        if ng:
            self.head.g = plan.g_logits_all[b * S:(b + 1) * S]
        self.layers = [l for pr in self.enc for l in pr] + [l for dd in self.dec for l in dd[2:]] + [self.head]

    def forward(self, plan, rt, x, train):
        F_ = plan.net.filters
        d, h, w = plan.dims3
        src = x
        for i, (la, lb) in enumerate(self.enc):
            a = la.forward(rt, src, None, True)                   # InstanceNorm uses the sample's statistics in eval mode too
            lb.p_drop = P_DROP if (train and i == 4) else 0.0
            xi = lb.forward(rt, a, None, True)
            if i < 4:
                ops.maxpool3d_fwd(xi, self.pooled[i], 1, d, h, w, F_[i])
                src = self.pooled[i]
                d, h, w = d // 2, h // 2, w // 2
        cur = self.enc[4][1].a
        for j, (up, up_g, la, lb) in enumerate(self.dec):
            ops.upsample3d2x_fwd(cur, up, 1, d, h, w, F_[4 - j])
            d, h, w = 2 * d, 2 * h, 2 * w
            a = la.forward(rt, self.enc[3 - j][1].a, up, True)    # cat([skip, up], 1)  (utils.py:276)
            lb.p_drop = P_DROP if (train and j == 3) else 0.0
            cur = lb.forward(rt, a, None, True)
        self.head.forward(rt, cur, None, True)

    def backward(self, plan, rt, x, acc_w):
        F_ = plan.net.filters
        D, H, W = plan.dims3
        last = self.dec[3][3]
        self.head.backward(rt, last.a, None, last.g, accumulate_w=acc_w)
        d, h, w = D, H, W
        for j in range(3, -1, -1):
            up, up_g, la, lb = self.dec[j]
            skip_l = self.enc[3 - j][1]
            lb.backward(rt, la.a, None, la.g, accumulate_w=acc_w)
            # the skip gradient is the first contribution to the encoder feature's gradient (overwrite)
            la.backward(rt, skip_l.a, up, skip_l.g, up_g, accumulate_w=acc_w)
            d, h, w = d // 2, h // 2, w // 2
            src_l = self.dec[j - 1][3] if j > 0 else self.enc[4][1]
            ops.upsample3d2x_bwd(up_g, src_l.g, 1, d, h, w, F_[4 - j])
        for i in range(4, -1, -1):
            la, lb = self.enc[i]
            lb.backward(rt, la.a, None, la.g, accumulate_w=acc_w)
            if i > 0:
                la.backward(rt, self.pooled[i - 1], None, self.pooled_g[i - 1], accumulate_w=acc_w)
                prev = self.enc[i - 1][1]
                ops.maxpool3d_bwd(prev.a, self.pooled_g[i - 1], prev.g, 1, 2 * d, 2 * h, 2 * w, F_[i - 1], accumulate=True)
                d, h, w = 2 * d, 2 * h, 2 * w
            else:
                la.backward(rt, x, None, None, accumulate_w=acc_w)


class UNet3DPlan:
    """Buffers + launch schedule of one unet_3D for one input geometry (B, D, H, W)."""

    def __init__(self, net: "unet_3D", rt: Runtime, B, D, H, W, need_grad):
        assert D % 16 == 0 and H % 16 == 0 and W % 16 == 0, "unet_3D needs D, H, W divisible by 16"
        self.net, self.rt, self.B, self.dims3, self.need_grad = net, rt, B, (D, H, W), need_grad
        dev = rt.device
        self._inorm = {}
        S = D * H * W
        self.logits_all = torch.empty((B, net.n_classes, S), dtype=torch.float32, device=dev)
        self.g_logits_all = torch.empty((B * S, net.n_classes), dtype=torch.float32, device=dev) if need_grad else None
        self.samples = []
        for b in range(B):
            self.samples.append(_Sample(self, b))
        self.layers = [l for s in self.samples for l in s.layers]
        rt.alloc_scratch()
        self.packer = PackTable(self.layers, need_grad, dev)
        self.in_flight = False

    def inorm(self, C):
        if C not in self._inorm:
            self._inorm[C] = _InstanceNormParams(C, self.rt.device)
        return self._inorm[C]

    @property
    def logits(self):
        return self.logits_all

    @property
    def g_logits(self):          # channels-last d(loss)/d(logits), [B*D*H*W, C]
        return self.g_logits_all

    def forward(self, x, train=True, repack=True):
        """x: [B, in_channels = 1, D, H, W] (== channels-last for one channel) or [B*D*H*W, in_channels] channels-last."""
        rt = self.rt
        if repack:
            self.packer.run()
        S = self.dims3[0] * self.dims3[1] * self.dims3[2]
        self.x_in = x.reshape(self.B, S, -1)
        for b, s in enumerate(self.samples):
            s.forward(self, rt, self.x_in[b], train)
        return self.logits_all

    def backward(self, dlogits_cl=None):
        rt = self.rt
        if dlogits_cl is not None and dlogits_cl.data_ptr() != self.g_logits_all.data_ptr():
            self.g_logits_all.copy_(dlogits_cl.view_as(self.g_logits_all))
        for b, s in enumerate(self.samples):
            s.backward(self, rt, self.x_in[b], acc_w=(b > 0))
        rt.join_side()


class _UNet3DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, need_grad, x, *params):
        B, Cin, D, H, W = x.shape
        plan = net._get_plan(B, D, H, W, need_grad)
        xin = x.contiguous().float()
        if Cin != 1:
            cl = torch.empty((B * D * H * W, Cin), dtype=torch.float32, device=x.device)
            ops.nchw_to_nhwc(xin, cl, B, Cin, D * H * W)
            xin = cl
        net._rt.seed_off += 1
        out = plan.forward(xin, train=net.training)
        ctx.plan = plan
        plan.in_flight = need_grad
        return out.view(B, net.n_classes, D, H, W).clone()

    @staticmethod
    def backward(ctx, grad_out):
        plan = ctx.plan
        if not plan.in_flight:
            raise RuntimeError("unet_3D backward called twice or after its buffers were reused")
        net = plan.net
        B, C = grad_out.shape[:2]
        S = grad_out[0, 0].numel()
        ops.nchw_to_nhwc(grad_out.contiguous().float(), plan.g_logits_all, B, C, S)
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


class unet_3D(nn.Module):
    """Drop-in for networks.unet_3D.unet_3D (code/networks/unet_3D.py:20) as net_factory_3d builds it (is_batchnorm=True)."""

    _instances = 0
    plan_key_is_batch = False

    def __init__(self, feature_scale=4, n_classes=21, is_deconv=True, in_channels=3, is_batchnorm=True, seed=None, exact=False):
        super().__init__()
        if not is_batchnorm:
            raise NotImplementedError("only is_batchnorm=True (what net_factory_3d builds) is implemented")
        self.feature_scale, self.n_classes, self.in_channels, self.is_deconv, self.is_batchnorm = \
            feature_scale, n_classes, in_channels, is_deconv, is_batchnorm
        self.filters = f = [int(x / feature_scale) for x in (64, 128, 256, 512, 1024)]
        assert all(c % 16 == 0 for c in f), "feature_scale must leave multiples of 16 channels"
        self.conv1 = _UnetConv3(in_channels, f[0])
        self.maxpool1 = nn.MaxPool3d(kernel_size=(2, 2, 2))
        self.conv2 = _UnetConv3(f[0], f[1])
        self.maxpool2 = nn.MaxPool3d(kernel_size=(2, 2, 2))
        self.conv3 = _UnetConv3(f[1], f[2])
        self.maxpool3 = nn.MaxPool3d(kernel_size=(2, 2, 2))
        self.conv4 = _UnetConv3(f[2], f[3])
        self.maxpool4 = nn.MaxPool3d(kernel_size=(2, 2, 2))
        self.center = _UnetConv3(f[3], f[4])
        self.up_concat4 = _UnetUp3_CT(f[4], f[3])
        self.up_concat3 = _UnetUp3_CT(f[3], f[2])
        self.up_concat2 = _UnetUp3_CT(f[2], f[1])
        self.up_concat1 = _UnetUp3_CT(f[1], f[0])
        self.final = nn.Conv3d(f[0], n_classes, 1)
        self.dropout1 = nn.Dropout(p=P_DROP)
        self.dropout2 = nn.Dropout(p=P_DROP)
        for m in self.modules():                     # init_weights(m, 'kaiming') (networks_other.py: kaiming_normal_, fan_in)
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
        if seed is None:
            seed = 5353 + 1000003 * unet_3D._instances
        unet_3D._instances += 1
        self._seed, self._exact = seed, exact
        self._flat, self._rt, self._plans = None, None, {}

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat, self._plans = None, {}
        return out

    def materialize(self):
        dev = next(self.parameters()).device
        if self._flat is None:
            self._flat = FlatParams(self, dev)
            self._rt = Runtime(dev, self._seed, self._exact)
            self._plans = {}
        return self._flat

    def _get_plan(self, B, D, H, W, need_grad) -> UNet3DPlan:
        self.materialize()
        pool = self._plans.setdefault((B, D, H, W, need_grad), [])
        for pl in pool:
            if not pl.in_flight:
                return pl
        if len(pool) >= 2:
            pool[0].in_flight = False
            return pool[0]
        pl = UNet3DPlan(self, self._rt, B, D, H, W, need_grad)
        pool.append(pl)
        return pl

    def forward(self, inputs):
        self.materialize()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._flat.params)
        return _UNet3DFn.apply(self, need_grad, inputs, *self._flat.params)
