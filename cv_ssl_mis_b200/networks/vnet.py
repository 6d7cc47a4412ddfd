"""3D VNet of the reference (code/networks/vnet.py:145-239) on the B200 kernels.

Same construction as `net_factory_3d("vnet")` uses (code/networks/net_factory_3d.py:18-20):
normalization='batchnorm', has_dropout=True.  The module tree owns parameters/buffers under the reference's
state_dict keys (`block_one.conv.0.weight`, `block_five_up.conv.1.running_mean`, `out_conv.bias`, ...);
`forward` runs `VNetPlan` over channels-last (NDHWC) buffers:

    stage:       n x [conv3x3x3 + BN3d + ReLU]                               (ConvBlock,               vnet.py:5-31)
    down:        conv 2x2x2 stride 2 + BN3d + ReLU                           (DownsamplingConvBlock,   vnet.py:67-91)
    up:          ConvTranspose 2x2x2 stride 2 + BN3d + ReLU, then + skip     (UpsamplingDeconvBlock,   vnet.py:94-118,210-222)
    Dropout3d(0.5) on x5 and x9 (whole channels per sample)                   (vnet.py:195-196,225-226)
    head:        conv 1x1x1 -> logits, stored NCDHW for the caller            (vnet.py:175,227)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ._engine import ConvLayer, FlatParams, PackTable, Runtime

STAGES = [1, 2, 3, 3, 3, 3, 3, 2, 1]           # block_one .. block_nine


def _conv_block(n_stages, cin, cout):
    ops_ = []
    for i in range(n_stages):
        ops_ += [nn.Conv3d(cin if i == 0 else cout, cout, 3, padding=1), nn.BatchNorm3d(cout), nn.ReLU(inplace=True)]
    return ops_


class _Seq(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.conv = nn.Sequential(*layers)


class VNetPlan:
    """Buffers + launch schedule of one VNet for one input geometry (B, D, H, W)."""

    def __init__(self, net: "VNet", rt: Runtime, B, D, H, W, need_grad):
        assert D % 16 == 0 and H % 16 == 0 and W % 16 == 0, "VNet needs D, H, W divisible by 16"
        self.net, self.rt, self.B, self.dims3, self.need_grad = net, rt, B, (D, H, W), need_grad
        dev = rt.device
        nf = net.n_filters
        names = ["one", "two", "three", "four", "five", "six", "seven", "eight", "nine"]
        self.stages = []         # list of lists of ConvLayer
        self.down, self.up, self.sums, self.sums_g = [], [], [], []
        d, h, w = D, H, W
        c = net.n_channels
        stream = 0
        # encoder
        for s in range(5):
            cout = nf * (2 ** s)
            blk = getattr(net, f"block_{names[s]}").conv
            layers = []
            for i in range(STAGES[s]):
                last = i == STAGES[s] - 1
                drop = (0.5, 2, stream) if (s == 4 and last and net.has_dropout) else (0.0, 0, 0)
                layers.append(ConvLayer(blk[3 * i], blk[3 * i + 1], 0.0, drop[0], drop[1], drop[2], dims=3,
                                        name=f"b{s + 1}c{i}").plan(rt, B, d, h, w, c if i == 0 else cout, 0, need_grad))
                c = cout if i == 0 else c
            c = cout
            self.stages.append(layers)
            if s < 4:
                dw = getattr(net, f"block_{names[s]}_dw").conv
                self.down.append(ConvLayer(dw[0], dw[1], 0.0, dims=3, name=f"b{s + 1}dw").plan(rt, B, d, h, w, c, 0, need_grad))
                d, h, w, c = d // 2, h // 2, w // 2, c * 2
        stream = 1
        # decoder
        for s in range(5, 9):
            upm = getattr(net, f"block_{names[s - 1]}_up").conv
            cout = c // 2
            self.up.append(ConvLayer(upm[0], upm[1], 0.0, kind="deconv", dims=3, name=f"b{s}up").plan(rt, B, d, h, w, c, 0, need_grad))
            d, h, w, c = d * 2, h * 2, w * 2, cout
            self.sums.append(torch.empty((B * d * h * w, c), dtype=torch.float32, device=dev))
            blk = getattr(net, f"block_{names[s]}").conv
            layers = []
            for i in range(STAGES[s]):
                last = i == STAGES[s] - 1
                drop = (0.5, 2, stream) if (s == 8 and last and net.has_dropout) else (0.0, 0, 0)
                layers.append(ConvLayer(blk[3 * i], blk[3 * i + 1], 0.0, drop[0], drop[1], drop[2], dims=3,
                                        name=f"b{s + 1}c{i}").plan(rt, B, d, h, w, c, 0, need_grad))
            self.stages.append(layers)
        self.head = ConvLayer(net.out_conv, dims=3, out_nchw=True, name="out").plan(rt, B, D, H, W, nf, 0, need_grad)
        self.layers = [l for st in self.stages for l in st] + self.down + self.up + [self.head]
        rt.alloc_scratch()
        self.packer = PackTable(self.layers, need_grad, dev)
        self.in_flight = False

    @property
    def logits(self):
        return self.head.y

    @property
    def g_logits(self):          # channels-last d(loss)/d(logits), [B*D*H*W, C]
        return self.head.g

    def forward(self, x, train=True, repack=True):
        rt = self.rt
        if repack:                       # False: the weights have not changed since this plan's previous forward
            self.packer.run()
        self.x_in = x
        cur = x
        for s in range(5):
            for l in self.stages[s]:
                cur = l.forward(rt, cur, None, train)
            if s < 4:
                cur = self.down[s].forward(rt, cur, None, train)
        for k in range(4):
            upa = self.up[k].forward(rt, cur, None, train)
            skip = self.stages[3 - k][-1].a
            ops.add(upa, skip, self.sums[k])                       # x_up + x_skip (vnet.py:210,214,218,222)
            cur = self.sums[k]
            for l in self.stages[5 + k]:
                cur = l.forward(rt, cur, None, train)
        self.head.forward(rt, cur, None, train)
        return self.head.y

    def backward(self, dlogits_cl=None):
        rt = self.rt
        if dlogits_cl is not None and dlogits_cl.data_ptr() != self.head.g.data_ptr():
            self.head.g.copy_(dlogits_cl.view_as(self.head.g))
        last = self.stages[8][-1]
        self.head.backward(rt, last.a, None, last.g)
        for k in range(3, -1, -1):
            st = self.stages[5 + k]
            skip_l = self.stages[3 - k][-1]
            up_l = self.up[k]
            for i in range(len(st) - 1, 0, -1):
                st[i].backward(rt, st[i - 1].a, None, st[i - 1].g)
            # d(sum) is the gradient of both addends: written once into the skip's gradient buffer (first
            # contribution, overwrite), and read from there by the up-conv's BN backward (g_in)
            st[0].backward(rt, self.sums[k], None, skip_l.g)
            src_l = self.stages[4 + k][-1]
            up_l.backward(rt, src_l.a, None, src_l.g, g_in=skip_l.g)
        for s in range(4, -1, -1):
            st = self.stages[s]
            for i in range(len(st) - 1, 0, -1):
                st[i].backward(rt, st[i - 1].a, None, st[i - 1].g)
            if s > 0:
                dw = self.down[s - 1]
                st[0].backward(rt, dw.a, None, dw.g)
                prev = self.stages[s - 1][-1]
                dw.backward(rt, prev.a, None, prev.g, accumulate_dx=True)     # adds to the skip gradient
            else:
                st[0].backward(rt, self.x_in, None, None)
        rt.join_side()


class _VNetFn(torch.autograd.Function):
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
            raise RuntimeError("VNet backward called twice or after its buffers were reused")
        net = plan.net
        B, C = grad_out.shape[:2]
        S = grad_out[0, 0].numel()
        ops.nchw_to_nhwc(grad_out.contiguous().float(), plan.head.g, B, C, S)
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


class VNet(nn.Module):
    """Drop-in for networks.vnet.VNet (code/networks/vnet.py:145) with normalization='batchnorm'."""

    _instances = 0

    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization="batchnorm", has_dropout=False, seed=None,
                 exact=False):
        super().__init__()
        if normalization != "batchnorm":
            raise NotImplementedError("only normalization='batchnorm' (what net_factory_3d builds) is implemented")
        self.n_channels, self.n_classes, self.n_filters, self.has_dropout = n_channels, n_classes, n_filters, has_dropout
        nf = n_filters
        names = ["one", "two", "three", "four", "five", "six", "seven", "eight", "nine"]
        chans = [nf, 2 * nf, 4 * nf, 8 * nf, 16 * nf, 8 * nf, 4 * nf, 2 * nf, nf]
        for s, name in enumerate(names):
            cin = n_channels if s == 0 else chans[s]
            setattr(self, f"block_{name}", _Seq(_conv_block(STAGES[s], cin, chans[s])))
            if s < 4:
                setattr(self, f"block_{name}_dw", _Seq([nn.Conv3d(chans[s], chans[s + 1], 2, padding=0, stride=2),
                                                        nn.BatchNorm3d(chans[s + 1]), nn.ReLU(inplace=True)]))
            elif s < 8:
                setattr(self, f"block_{name}_up", _Seq([nn.ConvTranspose3d(chans[s], chans[s + 1], 2, padding=0, stride=2),
                                                        nn.BatchNorm3d(chans[s + 1]), nn.ReLU(inplace=True)]))
        self.out_conv = nn.Conv3d(nf, n_classes, 1, padding=0)
        self.dropout = nn.Dropout3d(p=0.5, inplace=False)
        if seed is None:
            seed = 4242 + 1000003 * VNet._instances
        VNet._instances += 1
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

    def _get_plan(self, B, D, H, W, need_grad) -> VNetPlan:
        self.materialize()
        pool = self._plans.setdefault((B, D, H, W, need_grad), [])
        for pl in pool:
            if not pl.in_flight:
                return pl
        if len(pool) >= 2:
            pool[0].in_flight = False
            return pool[0]
        pl = VNetPlan(self, self._rt, B, D, H, W, need_grad)
        pool.append(pl)
        return pl

    def forward(self, input, turnoff_drop=False):
        self.materialize()
        if turnoff_drop:
            raise NotImplementedError("turnoff_drop is never used by the reference trainers")
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._flat.params)
        return _VNetFn.apply(self, need_grad, input, *self._flat.params)
