"""Host-side building blocks shared by the network engines.

`FlatParams` keeps all parameters (and their gradients) of a module in two flat fp32 device buffers so the
optimizer/EMA/all-reduce each touch one contiguous range.  `ConvLayer` is one convolution of the reference
networks together with the BatchNorm -> activation -> dropout that follows it, executed through the C ABI
(`ops`), forward and backward, on channels-last buffers that are allocated once per input geometry.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.nn as nn

from .. import ops, _lib
from .._lib import (PACK_CONV_FWD, PACK_CONV_DGRAD, PACK_CONV_DGRAD_D2S, PACK_DECONV_FWD, PACK_DECONV_DGRAD)


class FlatParams:
    """Re-homes module.parameters() (in registration order, like the reference's EMA zip) into one flat buffer."""

    def __init__(self, module: nn.Module, device):
        self.params = [p for p in module.parameters()]
        self.numel = sum(p.numel() for p in self.params)
        # every tensor starts on a 16-byte boundary (vectorised optimizer); padding is zero and stays zero
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.padded = off
        self.data = torch.zeros(self.padded, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.padded, dtype=torch.float32, device=device)
        for p, o in zip(self.params, self.offsets):
            view = self.data[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def grad_of(self, p):
        return p.grad


class Runtime:
    """Per-engine execution context: precision mode, RNG key and shared scratch."""

    def __init__(self, device, seed=1337, exact=False):
        self.device = device
        self.exact = exact
        self.seed = int(seed) & 0x7FFFFFFFFFFFFFFF
        # bumped once per forward so every pass draws fresh dropout masks, including CUDA-graph replays
        self.seed_off = torch.zeros(1, dtype=torch.int64, device=device)
        self.scratch = None
        self.scratch_bytes = 0
        # weight gradients only feed the optimizer, so they run on a side stream next to the data-gradient chain
        # (own scratch buffer; joined at the end of the plan's backward).  B200_OVERLAP=0 keeps one stream.
        self.overlap = torch.device(device).type == "cuda" and os.environ.get("B200_OVERLAP", "1") != "0"
        self.side = torch.cuda.Stream(device=device) if self.overlap else None
        self.scratch_side = None

    def need_scratch(self, nbytes):
        self.scratch_bytes = max(self.scratch_bytes, int(nbytes))

    def alloc_scratch(self):
        if self.scratch is None or self.scratch.numel() * 4 < self.scratch_bytes:
            # a captured CUDA graph has the old scratch pointers baked in: keep every superseded buffer alive so that its
            # replays never write into memory the caching allocator has handed to someone else
            if self.scratch is not None:
                self._retired = getattr(self, "_retired", []) + [self.scratch, self.scratch_side]
            self.scratch = torch.empty((self.scratch_bytes + 3) // 4 + 4, dtype=torch.float32, device=self.device)
            self.scratch_side = torch.empty_like(self.scratch) if self.overlap else self.scratch

    def side_stream(self):
        """Context: run the enclosed launches on the side stream, after everything issued so far on the current one."""
        if not self.overlap:
            return contextlib.nullcontext()
        self.side.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.side)

    def join_side(self):
        if self.overlap:
            torch.cuda.current_stream().wait_stream(self.side)


class ConvLayer:
    """conv (or k2s2 transposed conv) [+ BatchNorm(train) + LeakyReLU/ReLU + Dropout] on channels-last buffers."""

    def __init__(self, conv, bn=None, slope=None, p_drop=0.0, drop_mode=0, rng_stream=0, kind="conv", dims=2,
                 out_nchw=False, name=""):
        self.conv, self.bn, self.slope = conv, bn, slope
        self.p_drop, self.drop_mode, self.rng_stream = float(p_drop), int(drop_mode), int(rng_stream)
        self.kind, self.dims, self.out_nchw, self.name = kind, dims, out_nchw, name
        self.k = conv.kernel_size[0]
        self.stride = conv.stride[0]
        self.pad = conv.padding[0]
        if kind == "conv":
            self.cout, self.cin = conv.weight.shape[0], conv.weight.shape[1]
        else:
            self.cin, self.cout = conv.weight.shape[0], conv.weight.shape[1]
        self.T = self.k ** dims
        self.has_act = bn is not None
        assert (bn is None) == (slope is None)

    # ---- planning: geometry -> descriptor + buffers
    def plan(self, rt: Runtime, n, id_, ih, iw, c0, c1, need_grad):
        assert c0 + c1 == self.cin, (self.name, c0, c1, self.cin)
        dev = rt.device
        self.n, self.c0, self.c1 = n, c0, c1
        self.desc = ops.conv_desc(n, id_, ih, iw, c0, c1, self.cout, self.k, self.stride, self.pad, self.dims)
        if self.kind == "conv":
            self.od, self.oh, self.ow = ops.desc_out_dims(self.desc)
        else:
            self.od, self.oh, self.ow = (id_ * 2 if self.dims == 3 else id_), ih * 2, iw * 2
        self.M = n * self.od * self.oh * self.ow
        self.spatial = self.od * self.oh * self.ow
        shape = (n, self.cout, self.spatial) if self.out_nchw else (self.M, self.cout)
        self.y = torch.empty(shape, dtype=torch.float32, device=dev)
        self.a = torch.empty_like(self.y) if self.has_act else self.y
        # d/d(output) and then, in place, d/d(raw conv out); always channels-last
        self.g = torch.empty((self.M, self.cout), dtype=torch.float32, device=dev) if need_grad else None
        if self.has_act:
            self.state = torch.empty(4 * self.cout, dtype=torch.float32, device=dev)
            rt.need_scratch(ops.bn_workspace_bytes(self.M, self.cout))
        O, I = self.cout, self.cin
        # production path for 3x3 / 3x3x3 stride-1 convs: shared-memory tile kernels (TF32); the generic implicit
        # GEMM serves everything else and the `exact` (3xTF32) validation mode
        is_conv = self.kind == "conv"
        # 1x1 (1x1x1) convolutions are plain GEMMs over the channels-last rows: tcgen05/TMA GEMM, no weight packing
        self.gemm = (is_conv and self.k == 1 and self.stride == 1 and not rt.exact and not self.out_nchw
                     and self.cout >= 32 and self.cin >= 32 and ops.linear_supported(self.M, self.cout, c0, c1))
        if self.gemm:
            self.umma_fwd = self.umma_dgrad = self.use_c1_kernel = self.tile_fwd = self.tile_dgrad = self.tile_wgrad = self.row_wgrad = False
            self.row_fwd = self.row_dgrad = self.blk_fwd = self.blk_dgrad = False
            self.wp_fwd = self.wp_bwd = None
            if need_grad:
                rt.need_scratch(max(ops.linear_wgrad_workspace_bytes(self.M, self.cout, self.cin),
                                    ops.colsum_workspace_bytes(self.M, self.cout)))
            return self
        # 3D 2x2x2 stride-2 convs / transposed convs (vnet.py:73,100, the UNETR up-blocks): every input voxel is used once,
        # so the forward is a pure GEMM over the space-to-depth view (tcgen05 GEMM + one 16-byte copy kernel).
        # B200_K2S2=generic keeps the implicit-GEMM kernels for A/B runs; the backward still uses them.
        self.k2s2_gemm = False
        self.wp_gemm = None
        if (self.dims == 3 and self.k == 2 and self.stride == 2 and self.pad == 0 and c1 == 0 and not rt.exact and not self.out_nchw
                and os.environ.get("B200_K2S2", "gemm") == "gemm" and c0 % 4 == 0 and self.cout % 4 == 0):
            if is_conv and id_ % 2 == 0 and ih % 2 == 0 and iw % 2 == 0:
                self.k2s2_gemm = bool(ops.linear_supported(self.M, self.cout, 8 * c0, 0))
                shape = (self.M, 8 * c0)
            elif not is_conv:
                self.k2s2_gemm = bool(ops.linear_supported(n * id_ * ih * iw, 8 * self.cout, c0, 0))
                shape = (n * id_ * ih * iw, 8 * self.cout)
            if self.k2s2_gemm:
                self.gview = torch.empty(shape, dtype=torch.float32, device=dev)       # xs (conv) / ys, then d(ys) (transposed conv)
                if is_conv and need_grad:
                    # strided conv backward on the same GEMM: d(xs) = dy W2 scattered depth-to-space, dW2 = dy^T xs
                    self.gview_g = torch.empty(shape, dtype=torch.float32, device=dev)
                    self.dw_gemm = torch.empty((self.cout, 8 * c0), dtype=torch.float32, device=dev)
                    rt.need_scratch(ops.linear_wgrad_workspace_bytes(self.M, self.cout, 8 * c0))
                if not is_conv and need_grad:
                    # transposed conv backward on the same GEMM: d(ys) = space-to-depth view of dy, dx = d(ys) W2d,
                    # dW2d = d(ys)^T x (copied into the framework layout [cin][cout][2][2][2])
                    self.dw_gemm = torch.empty((8 * self.cout, c0), dtype=torch.float32, device=dev)
                    rt.need_scratch(ops.linear_wgrad_workspace_bytes(shape[0], 8 * self.cout, c0))
                self.gemm_mode = PACK_CONV_DGRAD_D2S if is_conv else PACK_DECONV_DGRAD  # [cout][8 cin] / [8 cout][cin]
                self.wp_gemm = torch.empty(ops.conv_packed_floats(self.gemm_mode, O, I, self.T), dtype=torch.float32, device=dev)
        # 2D 3x3: tcgen05/TMEM kernel (B200_CONV=tile falls back to the mma.sync tile kernel for A/B comparisons)
        want_umma = is_conv and not rt.exact and os.environ.get("B200_CONV", "umma") == "umma"
        # measured (tools/bench_conv.py): with fp32 operands the UMMA is bound by its shared-memory operand reads,
        # 4(M+N)/(MN) bytes per MAC, so it only beats mma.sync when both channel counts are >= 32
        wide = self.cout >= 32 and self.cin >= 32
        # wide images (W % 128 == 0: the 16 / 32-channel levels): row-ring tcgen05 kernel, weights resident in shared
        # memory, BatchNorm statistics taken in the epilogue.  B200_CONV_ROW=0 keeps the older kernels for A/B runs.
        want_row = is_conv and not rt.exact and os.environ.get("B200_CONV_ROW", "1") != "0"
        self.row_fwd = ops.conv_row_supported(self.desc, False) if (want_row and not self.out_nchw) else 0     # 8 + pack mode, or 0
        self.row_dgrad = ops.conv_row_supported(self.desc, True) if (want_row and need_grad) else 0
        # narrow images (the 64^2 / 32^2 / 16^2 levels): halo-block tcgen05 kernel (TMA halo planes, all accumulator
        # blocks of an item in TMEM, statistics in the epilogue).  B200_CONV_BLK=0 keeps conv_umma2 for A/B runs.
        want_blk = want_row and os.environ.get("B200_CONV_BLK", "1") != "0"
        self.blk_fwd = ops.conv_blk_supported(self.desc, False) if (want_blk and not self.row_fwd and not self.out_nchw) else 0    # 8 + pack mode
        self.blk_dgrad = ops.conv_blk_supported(self.desc, True) if (want_blk and need_grad and not self.row_dgrad) else 0
        self.umma_fwd = (want_umma and wide and ops.conv_umma_supported(self.desc, False) and not self.row_fwd
                         and not self.blk_fwd and (self.out_nchw or self.cout % 4 == 0))
        self.umma_dgrad = (want_umma and wide and ops.conv_umma_supported(self.desc, True) and not self.row_dgrad
                           and not self.blk_dgrad)
        self.use_c1_kernel = is_conv and not rt.exact and ops.conv_c1_supported(self.desc)       # first layer: FFMA kernels
        self.tile_fwd = is_conv and not rt.exact and ops.conv_tile_supported(self.desc, False)
        self.tile_dgrad = self.tile_fwd and self.cout % 4 == 0 and c0 % 2 == 0 and c1 % 2 == 0
        self.tile_wgrad = is_conv and not rt.exact and ops.conv_tile_supported(self.desc, True)
        # 2D 3x3 weight gradient on tcgen05 (row-ring kernel); B200_WGRAD=tile keeps the mma.sync kernel for A/B runs
        self.row_wgrad = (is_conv and not rt.exact and need_grad and os.environ.get("B200_WGRAD", "row") == "row"
                          and ops.conv_row_wgrad_supported(self.desc))
        fwd_mode = PACK_CONV_FWD if is_conv else PACK_DECONV_FWD
        if (self.row_fwd or self.blk_fwd) and self.has_act:
            self.stats_blocks = ops.conv_row_stats_blocks(self.desc) if self.row_fwd else ops.conv_blk_stats_blocks(self.desc)
            rt.need_scratch(self.stats_blocks * 2 * self.cout * 8)
        nfwd = (ops.conv_row_packed_floats(self.desc, False) if self.row_fwd else self.T * O * I if self.blk_fwd else
                ops.conv_umma_packed_floats(False, O, I, self.T) if self.umma_fwd else
                ops.conv_tile_packed_floats(False, O, I, self.T) if self.tile_fwd else ops.conv_packed_floats(fwd_mode, O, I, self.T))
        self.wp_fwd = torch.empty(nfwd, dtype=torch.float32, device=dev)
        self.wp_bwd = None
        if need_grad:
            if is_conv:
                self.bwd_mode = PACK_CONV_DGRAD if self.stride == 1 else PACK_CONV_DGRAD_D2S
                rt.need_scratch(ops.conv_c1_wgrad_workspace_bytes(self.desc) if self.use_c1_kernel else
                                max(ops.conv_row_wgrad_workspace_bytes(self.desc), ops.colsum_workspace_bytes(self.M, self.cout))
                                if self.row_wgrad else
                                ops.conv_tile_wgrad_workspace_bytes(self.desc) if self.tile_wgrad
                                else ops.conv_wgrad_workspace_bytes(self.desc))
            else:
                self.bwd_mode = PACK_DECONV_DGRAD
                rt.need_scratch(max(ops.deconv_k2s2_wgrad_workspace_bytes(self.desc),
                                    ops.colsum_workspace_bytes(self.M, self.cout)))
            nbwd = (ops.conv_row_packed_floats(self.desc, True) if self.row_dgrad else self.T * O * I if self.blk_dgrad else
                    ops.conv_umma_packed_floats(True, O, I, self.T) if self.umma_dgrad else
                    ops.conv_tile_packed_floats(True, O, I, self.T) if self.tile_dgrad else ops.conv_packed_floats(self.bwd_mode, O, I, self.T))
            self.wp_bwd = torch.empty(nbwd, dtype=torch.float32, device=dev)
        return self

    def pack_jobs(self, need_dgrad):
        """(weight, packed buffer, kind, mode, O, I, T) tuples for the one-launch packer (ops.conv_pack_batch)."""
        O, I, T, w = self.cout, self.cin, self.T, self.conv.weight
        jobs = []
        if self.gemm:
            return jobs
        if self.wp_gemm is not None:
            jobs.append((w, self.wp_gemm, 0, self.gemm_mode, O, I, T))
        if not self.use_c1_kernel:
            if self.row_fwd:
                jobs.append((w, self.wp_fwd, 3, self.row_fwd - 8, O, I, T))
            elif self.blk_fwd:
                jobs.append((w, self.wp_fwd, 3, self.blk_fwd - 8, O, I, T))
            elif self.umma_fwd:
                jobs.append((w, self.wp_fwd, 2, 0, O, I, T))
            elif self.tile_fwd:
                jobs.append((w, self.wp_fwd, 1, 0, O, I, T))
            else:
                jobs.append((w, self.wp_fwd, 0, PACK_CONV_FWD if self.kind == "conv" else PACK_DECONV_FWD, O, I, T))
        if need_dgrad and self.wp_bwd is not None:
            if self.row_dgrad:
                jobs.append((w, self.wp_bwd, 3, self.row_dgrad - 8, O, I, T))
            elif self.blk_dgrad:
                jobs.append((w, self.wp_bwd, 3, self.blk_dgrad - 8, O, I, T))
            elif self.umma_dgrad:
                jobs.append((w, self.wp_bwd, 2, 1, O, I, T))
            elif self.tile_dgrad:
                jobs.append((w, self.wp_bwd, 1, 1, O, I, T))
            else:
                jobs.append((w, self.wp_bwd, 0, self.bwd_mode, O, I, T))
        return jobs

    def pack(self, need_dgrad):
        O, I = self.cout, self.cin
        if self.gemm:
            return
        if self.wp_gemm is not None:
            ops.conv_pack_weights(self.conv.weight, self.wp_gemm, self.gemm_mode, O, I, self.T)
        if self.use_c1_kernel:
            pass
        elif self.row_fwd:
            ops.conv_row_pack_weights(self.desc, False, self.conv.weight, self.wp_fwd)
        elif self.blk_fwd:
            ops.conv_blk_pack_weights(self.conv.weight, self.wp_fwd, self.blk_fwd - 8, O, I, self.T)
        elif self.umma_fwd:
            ops.conv_umma_pack_weights(self.conv.weight, self.wp_fwd, False, O, I, self.T)
        elif self.tile_fwd:
            ops.conv_tile_pack_weights(self.conv.weight, self.wp_fwd, False, O, I, self.T)
        else:
            ops.conv_pack_weights(self.conv.weight, self.wp_fwd, PACK_CONV_FWD if self.kind == "conv" else PACK_DECONV_FWD, O, I, self.T)
        if need_dgrad and self.wp_bwd is not None:
            if self.row_dgrad:
                ops.conv_row_pack_weights(self.desc, True, self.conv.weight, self.wp_bwd)
            elif self.blk_dgrad:
                ops.conv_blk_pack_weights(self.conv.weight, self.wp_bwd, self.blk_dgrad - 8, O, I, self.T)
            elif self.umma_dgrad:
                ops.conv_umma_pack_weights(self.conv.weight, self.wp_bwd, True, O, I, self.T)
            elif self.tile_dgrad:
                ops.conv_tile_pack_weights(self.conv.weight, self.wp_bwd, True, O, I, self.T)
            else:
                ops.conv_pack_weights(self.conv.weight, self.wp_bwd, self.bwd_mode, O, I, self.T)

    # ---- forward
    def forward(self, rt: Runtime, src0, src1=None, train=True):
        _lib.tag = self.name
        self.bn_train = train
        if self.gemm:
            ops.linear_fwd(src0, src1, self.conv.weight.view(self.cout, self.cin), self.conv.bias, self.y, self.M, self.cout)
        elif self.k2s2_gemm and self.kind == "conv":
            d = self.desc
            ops.s2d_gather3d(src0, self.gview, d.n, d.id, d.ih, d.iw, self.cin)
            ops.linear_fwd(self.gview, None, self.wp_gemm.view(self.cout, 8 * self.cin), self.conv.bias, self.y, self.M, self.cout)
        elif self.k2s2_gemm:
            d = self.desc
            ops.linear_fwd(src0, None, self.wp_gemm.view(8 * self.cout, self.cin), None, self.gview, self.gview.shape[0], 8 * self.cout)
            ops.d2s_scatter3d(self.gview, self.conv.bias, self.y, d.n, d.id, d.ih, d.iw, self.cout)
        elif self.use_c1_kernel:
            ops.conv_c1_fwd(self.desc, src0, self.conv.weight, self.conv.bias, self.y)
        elif self.row_fwd:
            fused_stats = self.has_act and train
            ops.conv_row_fwd(self.desc, src0, src1, self.wp_fwd, self.conv.bias, self.y, rt.scratch if fused_stats else None)
        elif self.blk_fwd:
            fused_stats = self.has_act and train
            ops.conv_blk_fwd(self.desc, src0, src1, self.wp_fwd, self.conv.bias, self.y, rt.scratch if fused_stats else None)
        elif self.umma_fwd:
            ops.conv_umma_fwd(self.desc, src0, src1, self.wp_fwd, self.conv.bias, self.y, self.out_nchw)
        elif self.tile_fwd:
            ops.conv_tile_fwd(self.desc, src0, src1, self.wp_fwd, self.conv.bias, self.y, self.out_nchw)
        elif self.kind == "conv":
            ops.conv_fwd(self.desc, src0, src1, self.wp_fwd, self.conv.bias, self.y, self.out_nchw, rt.exact)
        else:
            ops.deconv_k2s2_fwd(self.desc, src0, self.wp_fwd, self.conv.bias, self.y, rt.exact)
        if not self.has_act:
            return self.y
        bn = self.bn
        if train and (self.row_fwd or self.blk_fwd):
            # the statistics pass happened in the convolution's epilogue: only the per-channel finalize is left
            ops.bn_finalize(rt.scratch, self.stats_blocks, self.M, self.cout, bn.weight, bn.bias, bn.eps, bn.momentum,
                            bn.running_mean, bn.running_var, self.state)
            ops.bn_act_fwd(self.y, self.state, self.a, self.M, self.cout, self.slope, self.p_drop, self.drop_mode,
                           rt.seed, rt.seed_off, self.rng_stream, self.spatial)
        elif train:
            ops.bn_stats_fwd(self.y, self.M, self.cout, bn.weight, bn.bias, bn.eps, bn.momentum, bn.running_mean,
                             bn.running_var, self.state, rt.scratch)
            ops.bn_act_fwd(self.y, self.state, self.a, self.M, self.cout, self.slope, self.p_drop, self.drop_mode,
                           rt.seed, rt.seed_off, self.rng_stream, self.spatial)
        else:
            ops.bn_eval_state(self.cout, bn.weight, bn.bias, bn.eps, bn.running_mean, bn.running_var, self.state)
            ops.bn_act_fwd(self.y, self.state, self.a, self.M, self.cout, self.slope, 0.0, 0, 0, None, 0, self.spatial)
        return self.a

    # ---- backward: self.g holds d/d(output); writes parameter grads and (optionally) input grads
    def backward(self, rt: Runtime, src0, src1=None, dx0=None, dx1=None, accumulate_dx=False, g_in=None, accumulate_w=False):
        """g_in: read d/d(output) from another buffer (left untouched) instead of self.g (needs has_act).
        accumulate_w: add to the parameter gradients (the same weights applied to several inputs, e.g. per-sample plans)."""
        conv, bn = self.conv, self.bn
        _lib.tag = self.name
        assert g_in is None or self.has_act
        if self.has_act:
            ops.bn_act_bwd(self.y, self.g if g_in is None else g_in, self.state, self.g, bn.weight.grad, bn.bias.grad, self.M, self.cout,
                           self.slope, rt.scratch, self.p_drop, self.drop_mode, rt.seed, rt.seed_off, self.rng_stream,
                           self.spatial)
        dy = self.g
        bias_grad = conv.bias.grad if conv.bias is not None else None
        if self.gemm:
            with rt.side_stream():
                ops.linear_wgrad(src0, src1, dy, conv.weight.grad.view(self.cout, self.cin), rt.scratch_side, self.M, self.cout,
                                 accumulate_w)
                if bias_grad is not None:
                    ops.colsum(dy, self.M, self.cout, bias_grad, rt.scratch_side, accumulate_w)
            if dx0 is not None:
                ops.linear_dgrad(dy, conv.weight.view(self.cout, self.cin), dx0, dx1, accumulate_dx, self.M, self.cout)
        elif self.kind == "conv" and self.k2s2_gemm:
            d, K8 = self.desc, 8 * self.cin
            with rt.side_stream():
                ops.linear_wgrad(self.gview, None, dy, self.dw_gemm, rt.scratch_side, self.M, self.cout, False)
                wg = conv.weight.grad.view(self.cout, self.cin, 8)
                part = self.dw_gemm.view(self.cout, 8, self.cin).permute(0, 2, 1)        # [co][(tap, ci)] -> [co][ci][tap]
                if accumulate_w:
                    wg.add_(part)
                else:
                    wg.copy_(part)
                if bias_grad is not None:
                    zero_db = self.has_act and getattr(self, "bn_train", True)           # bias in front of a train-mode BatchNorm
                    if zero_db and not accumulate_w:
                        bias_grad.zero_()
                    elif not zero_db:
                        ops.colsum(dy, self.M, self.cout, bias_grad, rt.scratch_side, accumulate_w)
            if dx0 is not None:
                ops.linear_dgrad(dy, self.wp_gemm.view(self.cout, K8), self.gview_g, None, False, self.M, self.cout)
                ops.d2s_scatter3d(self.gview_g, None, dx0, d.n, d.id // 2, d.ih // 2, d.iw // 2, self.cin, accumulate_dx)
        elif self.kind == "conv":
            with rt.side_stream():
                ws = rt.scratch_side
                if self.use_c1_kernel:
                    ops.conv_c1_wgrad(self.desc, src0, dy, ws, conv.weight.grad, bias_grad, accumulate_w)
                elif self.row_wgrad:
                    # a bias in front of a train-mode BatchNorm has an identically zero gradient (the BatchNorm backward
                    # removes the per-channel mean of dy); otherwise it is the column sum of dy
                    zero_db = self.has_act and getattr(self, "bn_train", True)
                    ops.conv_row_wgrad(self.desc, src0, src1, dy, ws, conv.weight.grad, accumulate_w,
                                       bias_grad if zero_db else None)
                    if bias_grad is not None and not zero_db:
                        ops.colsum(dy, self.M, self.cout, bias_grad, ws, accumulate_w)
                elif self.tile_wgrad:
                    ops.conv_tile_wgrad(self.desc, src0, src1, dy, ws, conv.weight.grad, bias_grad, accumulate_w)
                else:
                    ops.conv_wgrad(self.desc, src0, src1, dy, ws, conv.weight.grad, bias_grad, accumulate_w, rt.exact)
            if dx0 is not None:
                if self.row_dgrad:
                    ops.conv_row_dgrad(self.desc, dy, self.wp_bwd, dx0, dx1, accumulate_dx)
                elif self.blk_dgrad:
                    ops.conv_blk_dgrad(self.desc, dy, self.wp_bwd, dx0, dx1, accumulate_dx)
                elif self.umma_dgrad:
                    ops.conv_umma_dgrad(self.desc, dy, self.wp_bwd, dx0, dx1, accumulate_dx)
                elif self.tile_dgrad:
                    ops.conv_tile_dgrad(self.desc, dy, self.wp_bwd, dx0, dx1, accumulate_dx)
                elif self.stride == 1:
                    ops.conv_dgrad(self.desc, dy, self.wp_bwd, dx0, dx1, accumulate_dx, rt.exact)
                else:
                    ops.conv_k2s2_dgrad(self.desc, dy, self.wp_bwd, dx0, accumulate_dx, rt.exact)
        elif self.k2s2_gemm:
            d, M_in, O8 = self.desc, self.gview.shape[0], 8 * self.cout
            ops.s2d_gather3d(dy, self.gview, d.n, 2 * d.id, 2 * d.ih, 2 * d.iw, self.cout)
            with rt.side_stream():
                ops.linear_wgrad(src0, None, self.gview, self.dw_gemm, rt.scratch_side, M_in, O8, False)
                wg = conv.weight.grad.view(self.cin, self.cout, 8)
                part = self.dw_gemm.view(8, self.cout, self.cin).permute(2, 1, 0)          # [(tap, co)][ci] -> [ci][co][tap]
                if accumulate_w:
                    wg.add_(part)
                else:
                    wg.copy_(part)
                if bias_grad is not None:
                    ops.colsum(dy, self.M, self.cout, bias_grad, rt.scratch_side, accumulate_w)
            if dx0 is not None:
                ops.linear_dgrad(self.gview, self.wp_gemm.view(O8, self.cin), dx0, None, accumulate_dx, M_in, O8)
        else:
            with rt.side_stream():
                ops.deconv_k2s2_wgrad(self.desc, src0, dy, rt.scratch_side, conv.weight.grad, accumulate_w, rt.exact)
                if bias_grad is not None:
                    ops.colsum(dy, self.M, self.cout, bias_grad, rt.scratch_side, accumulate_w)
            if dx0 is not None:
                ops.deconv_k2s2_dgrad(self.desc, dy, self.wp_bwd, dx0, accumulate_dx, rt.exact)


class PackTable:
    """Device-side job table so that all weight packings of a plan run as ONE kernel launch per forward."""

    def __init__(self, layers, need_dgrad, device):
        self.jobs_py = [j for l in layers for j in l.pack_jobs(need_dgrad)]
        rows = [[w.data_ptr(), out.data_ptr(), kind, mode, O, I, T, out.numel()] for (w, out, kind, mode, O, I, T) in self.jobs_py]
        self.table = torch.tensor(rows, dtype=torch.int64, device=device)
        self.ptrs = [w.data_ptr() for (w, *_rest) in self.jobs_py]

    def run(self):
        # parameters may have been re-homed (e.g. .cuda() after planning): the table stores raw pointers
        assert self.ptrs == [w.data_ptr() for (w, *_r) in self.jobs_py], "parameters moved after planning"
        _lib.tag = "pack"
        ops.conv_pack_batch(self.table, len(self.jobs_py), 48, self.jobs_py)
