"""Fused training steps: the inner loops of the reference trainers as fixed launch schedules.

`CrossTeachingTrainer` is the CNN <-> Transformer iteration of
code/train_cross_teaching_between_cnn_transformer_2D.py:221-262 (two students, no EMA).
`MeanTeacherTrainer` covers, with one schedule,
  * Mean Teacher 2D/3D            code/train_mean_teacher_2D.py:201-238, code/train_mean_teacher_3D.py:134-166
  * Uncertainty-Aware Mean Teacher code/train_uncertainty_aware_mean_teacher_3D.py:135-189 (2D twin: ..._2D.py:147-201)
    (`uncertainty_T=8`: T/2 extra stochastic teacher passes on the twice-repeated unlabeled batch, entropy mask)
  * fully supervised               code/train_fully_supervised_2D.py:104-123, ..._3D_ViT.py   (`ema_model=None`)
  * Mean Teacher with Swin-UNets   code/train_mean_teacher_ViT.py:201-233 (the same loop over two ViT_seg models)
noise -> student forward -> teacher forward(s) (train mode, no grad) -> CE + Dice + (masked) consistency -> student
backward -> [grad all-reduce] -> SGD + EMA -> poly LR.  Everything between the host->device copy of the batch and
the loss read-back is asynchronous on one stream, allocates nothing, reads its per-step scalars from a small device
array and can therefore be replayed as a CUDA graph.
"""
from __future__ import annotations

import math
import os

import torch

from . import ops
from .utils import ramps

HP_LR, HP_MOMENTUM, HP_WD, HP_ALPHA, HP_ONE_MINUS_ALPHA, HP_GRAD_SCALE, HP_WCONS, HP_THRESHOLD = range(8)
NOISE_STREAM = 1000


class PendingLoss:
    """Handle returned by `submit`: the step's losses, available once ITS read-back (not the whole queue) has landed."""

    def __init__(self, host, event):
        self._host, self._event = host, event

    def result(self):
        if self._event is not None:
            self._event.synchronize()
        return self._host.tolist()


class _Resumable:
    """True resume, which the reference lacks (its checkpoints hold the student weights only): parameters, momentum buffers,
    the teacher, BatchNorm running statistics, the dropout / noise RNG epochs, the iteration counter and the learning rate."""

    def _networks(self):
        return [m for m in (getattr(self, "models", None) or (self.model, self.ema_model)) if m is not None]

    def state_dict(self):
        return {"iter_num": self.iter_num, "lr": self.lr, "tensors": [t.detach().clone().cpu() for t in self._state_tensors()],
                "rng_seeds": [m._rt.seed for m in self._networks()]}          # base seeds of the dropout / DropPath streams

    def load_state_dict(self, sd):
        """Call before the first step (a captured CUDA graph has the RNG seeds baked into its kernel arguments)."""
        assert getattr(self, "graph", None) is None, "load_state_dict after the step graph was captured"
        for m, seed in zip(self._networks(), sd["rng_seeds"]):
            m._rt.seed = seed
        ts = self._state_tensors()
        assert len(ts) == len(sd["tensors"]), "trainer state does not match this trainer (different models?)"
        for t, v in zip(ts, sd["tensors"]):
            assert t.shape == v.shape, (tuple(t.shape), tuple(v.shape))
            t.copy_(v)
        self.iter_num, self.lr = int(sd["iter_num"]), float(sd["lr"])


class _Pipelined(_Resumable):
    """`submit(images, labels)` is `step(..., read_loss=True)` without the stall: the pinned host batch goes up on a copy
    stream into one of two device staging pairs while the previous step is still running, the step is enqueued behind
    it, and the 16/32-byte loss read-back gets its own event.  A loop that calls `result()` of step i after `submit` of
    step i+1 keeps the GPU busy back to back; the host batch must stay untouched until its `result()` has returned."""
    _pipe = None

    def _loss_sources(self):
        return [self.lossbuf]

    def submit(self, images, labels, **kw):
        cuda = self.dev.type == "cuda"
        if self._pipe is None:
            n = 4 * len(self._loss_sources())
            self._pipe = {"i": 0, "copy": torch.cuda.Stream(device=self.dev) if cuda else None, "slots": [
                {"x": None, "y": None, "free": None, "host": torch.zeros(n).pin_memory() if cuda else torch.zeros(n)} for _ in range(2)]}
        pipe = self._pipe
        slot = pipe["slots"][pipe["i"] % 2]
        pipe["i"] += 1
        if cuda and not images.is_cuda:
            if slot["x"] is None:
                slot["x"] = torch.empty(images.shape, dtype=images.dtype, device=self.dev)
                slot["y"] = torch.empty(labels.shape, dtype=labels.dtype, device=self.dev)
            main, copy = torch.cuda.current_stream(), pipe["copy"]
            if slot["free"] is not None:
                copy.wait_event(slot["free"])              # the step that last read this staging pair has been enqueued past it
            with torch.cuda.stream(copy):
                slot["x"].copy_(images, non_blocking=True)
                slot["y"].copy_(labels, non_blocking=True)
            main.wait_stream(copy)
            images, labels = slot["x"], slot["y"]
        self.step(images, labels, **kw)
        ev = None
        if cuda:
            slot["free"] = torch.cuda.Event()
            slot["free"].record()
        host = slot["host"]
        for j, src in enumerate(self._loss_sources()):
            host[4 * j:4 * j + 4].copy_(src[:4], non_blocking=True)
        if cuda:
            ev = torch.cuda.Event()
            ev.record()
        return PendingLoss(host, ev)


class MeanTeacherTrainer(_Pipelined):
    def __init__(self, model, ema_model=None, *, batch_size=24, labeled_bs=12, patch_size=(256, 256), num_classes=4,
                 base_lr=0.01, max_iterations=30000, ema_decay=0.99, consistency=0.1, consistency_rampup=200.0,
                 momentum=0.9, weight_decay=1e-4, start_iter=0, consistency_gate_iters=1000, uncertainty_T=0,
                 label_dtype=None, noise_seed=2024, process_group=None, use_cuda_graph=False):
        self.model, self.ema_model = model, ema_model
        self.B = batch_size
        self.Lb = labeled_bs if ema_model is not None else batch_size
        self.U = self.B - self.Lb
        self.patch = tuple(patch_size)
        self.C = num_classes
        self.base_lr, self.max_iterations, self.ema_decay = base_lr, max_iterations, ema_decay
        self.consistency, self.consistency_rampup = consistency, consistency_rampup
        self.momentum, self.weight_decay = momentum, weight_decay
        self.gate = consistency_gate_iters
        self.T = uncertainty_T if ema_model is not None else 0
        assert self.T % 2 == 0
        self.iter_num = start_iter
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.noise_seed = noise_seed

        self.flat = model.materialize()
        dev = self.flat.data.device
        self.dev = dev
        self.ema_flat = ema_model.materialize() if ema_model is not None else None
        if self.ema_flat is not None:
            assert self.ema_flat.padded == self.flat.padded
        self.momentum_buf = torch.zeros_like(self.flat.data)
        # RNG epochs: the student's is bumped once per step (its backward must regenerate the forward's masks), the
        # teacher's before every teacher forward (each stochastic pass draws fresh noise and dropout)
        self.s_off = model._rt.seed_off
        self.t_off = ema_model._rt.seed_off if ema_model is not None else None
        pin = dev.type == "cuda"
        self.hp_ring = PinnedRing(8, dev)
        self.hp = torch.zeros(8, dtype=torch.float32, device=dev)
        S = 1
        for v in self.patch:
            S *= v
        self.S = S
        if label_dtype is None:
            label_dtype = torch.uint8 if len(self.patch) == 2 else torch.int64      # dataset.py:425 / brats2019.py:188
        self.x = torch.empty((self.B, 1, *self.patch), dtype=torch.float32, device=dev)
        self.y = torch.empty((self.B, *self.patch), dtype=label_dtype, device=dev)
        self.ema_in = torch.empty((self.U, 1, *self.patch), dtype=torch.float32, device=dev) if self.U else None
        self.lossbuf = torch.zeros(32, dtype=torch.float32, device=dev)
        self.loss_host = torch.zeros(4, dtype=torch.float32).pin_memory() if pin else torch.zeros(4)
        self.loss_ws = torch.empty(ops.ssl_loss_workspace_bytes(self.B, S) // 4 + 4, dtype=torch.float32, device=dev)
        self.s_plan = _plan_for(model, self.B, self.patch, True)
        self.t_plan = _plan_for(ema_model, self.U, self.patch, False) if self.U else None
        if self.T:
            self.x_rep = torch.empty((2 * self.U, 1, *self.patch), dtype=torch.float32, device=dev)
            self.ema_in2 = torch.empty_like(self.x_rep)
            self.t_plan2 = _plan_for(ema_model, 2 * self.U, self.patch, False)
            self.psum = torch.empty((self.U, self.C, S), dtype=torch.float32, device=dev)
        self.lr = base_lr                     # the reference installs the poly LR *after* each step (:234-236)
        self.use_graph = use_cuda_graph and dev.type == "cuda"
        self.graph = None
        self.comm = None
        self.kernel_launches_per_step = None

    # ---- host scalars (code/train_mean_teacher_2D.py:119-128,223-236)
    def consistency_weight(self, iter_num):
        if self.ema_model is None or iter_num < self.gate:
            return 0.0
        return self.consistency * ramps.sigmoid_rampup(iter_num // 150, self.consistency_rampup)

    def uncertainty_threshold(self, iter_num):
        """code/train_uncertainty_aware_mean_teacher_3D.py:175-176 (ln 2 whatever the class count)"""
        return (0.75 + 0.25 * ramps.sigmoid_rampup(iter_num, self.max_iterations)) * math.log(2)

    def _set_hparams(self):
        it = self.iter_num
        alpha = min(1 - 1 / (it + 1), self.ema_decay)
        h = [0.0] * 8
        h[HP_LR] = self.lr
        h[HP_MOMENTUM] = self.momentum
        h[HP_WD] = self.weight_decay
        h[HP_ALPHA] = alpha
        h[HP_ONE_MINUS_ALPHA] = 1 - alpha
        h[HP_GRAD_SCALE] = 1.0 / self.world
        h[HP_WCONS] = self.consistency_weight(it)
        h[HP_THRESHOLD] = self.uncertainty_threshold(it)
        self.hp_ring.push(h, self.hp)

    # ---- the device-side schedule (graph-capturable)
    def _device_step(self):
        self.s_off += 1
        self.model.train()
        xu = self.x[self.Lb:]
        teacher_logits = psum = thr = None
        # the teacher passes (no grad, own buffers and RNG epoch) run on a side stream next to the student forward
        t_rt = self.ema_model._rt if self.U else None
        if self.U:
            with t_rt.side_stream():
                self._teacher_passes(xu)
            teacher_logits = self.t_plan.logits
            if self.T:
                psum, thr = self.psum, self.hp[HP_THRESHOLD:HP_THRESHOLD + 1]
        self.s_plan.forward(self.x, train=True)
        if self.U:
            t_rt.join_side()
        w = self.hp[HP_WCONS:HP_WCONS + 1]
        ops.ssl_loss_fwd(self.s_plan.logits, teacher_logits, self.y, False, self.B, self.Lb, self.C, self.S, w,
                         self.lossbuf, self.loss_ws, psum, float(self.T), thr)
        ops.ssl_loss_bwd(self.s_plan.logits, teacher_logits, self.y, False, self.B, self.Lb, self.C, self.S, w,
                         self.lossbuf, 1.0, self.s_plan.g_logits, True, psum, float(self.T), thr)
        if self.world > 1 and hasattr(self.s_plan, "decoder_grad_offset") and os.environ.get("B200_DP_BUCKETS", "0") == "1":
            # opt-in: two gradient buckets -- the decoder's (produced first) is all-reduced on a communication stream while
            # the encoder's backward is still running; the encoder's follows on the main stream.  Measured with the 4.3 ms
            # step (profiles/r2_bench_config2_{2,8}gpu.json vs r2_dp_config2_{2,8}gpu_bucketed.json): 4.375 vs 4.395 ms at 2 GPUs,
            # 4.428 vs 4.444 ms at 8 -- the 7.3 MB exchange costs a fixed NCCL latency that is already hidden behind the
            # weight-gradient side stream, and the extra stream joins cost more than the overlap gains; the single
            # all-reduce below stays the default.
            off = self.s_plan.decoder_grad_offset()
            cuda = self.dev.type == "cuda"
            if cuda and self.comm is None:
                self.comm = torch.cuda.Stream(device=self.dev)
            main = torch.cuda.current_stream() if cuda else None

            def reduce_decoder():
                if not cuda:
                    torch.distributed.all_reduce(self.flat.grad[off:], group=self.pg)
                    return
                self.comm.wait_stream(main)
                if self.model._rt.side is not None:
                    self.comm.wait_stream(self.model._rt.side)
                with torch.cuda.stream(self.comm):
                    torch.distributed.all_reduce(self.flat.grad[off:], group=self.pg)

            self.s_plan.backward(None, after_decoder=reduce_decoder)
            torch.distributed.all_reduce(self.flat.grad[:off], group=self.pg)
            if cuda:
                main.wait_stream(self.comm)
        else:
            self.s_plan.backward(None)
            if self.world > 1:
                torch.distributed.all_reduce(self.flat.grad, group=self.pg)
        ops.sgd_ema_step(self.flat.data, self.flat.grad, self.momentum_buf,
                         self.ema_flat.data if self.ema_flat is not None else None, self.hp)

    def _teacher_passes(self, xu):
        self.t_off += 1
        ops.noise_add(xu, self.ema_in, 0.1, 0.2, self.noise_seed, self.t_off, NOISE_STREAM)
        self.t_plan.forward(self.ema_in, train=True)          # teacher stays in train mode (Appendix A.1)
        if self.T:
            # volume_batch_r = unlabeled.repeat(2, ...); T//2 noisy passes of the 2U batch (:153-160)
            self.x_rep[:self.U].copy_(xu)
            self.x_rep[self.U:].copy_(xu)
            for i in range(self.T // 2):
                self.t_off += 1
                ops.noise_add(self.x_rep, self.ema_in2, 0.1, 0.2, self.noise_seed, self.t_off, NOISE_STREAM)
                self.t_plan2.forward(self.ema_in2, train=True, repack=(i == 0))    # same EMA weights in all T//2 passes
                ops.mc_softmax_accumulate(self.t_plan2.logits, self.psum, 2, self.U, self.C, self.S, False, i == 0)

    def step(self, images, labels, read_loss=False):
        """images [B,1,*patch] float32, labels [B,*patch] uint8 (2D) / int64 (3D) -- host (ideally pinned) or device.
        Returns the device loss buffer [ce, dice, consistency, total, ...] (or host floats if read_loss)."""
        self._set_hparams()
        self.x.copy_(images, non_blocking=True)
        self.y.copy_(labels, non_blocking=True)
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._device_step()
        self.lr = self.base_lr * (1.0 - self.iter_num / self.max_iterations) ** 0.9
        self.iter_num += 1
        if read_loss:
            self.loss_host.copy_(self.lossbuf[:4], non_blocking=True)
            torch.cuda.current_stream().synchronize() if self.dev.type == "cuda" else None
            return self.loss_host.tolist()
        return self.lossbuf

    def _state_tensors(self):
        ts = [self.flat.data, self.momentum_buf, self.s_off]
        if self.ema_flat is not None:
            ts += [self.ema_flat.data, self.t_off]
        return ts + self._all_buffers()

    def _capture(self):
        # warm up once eagerly on a side stream (lazy module loading, allocator), restoring all state after
        keep = [t.clone() for t in self._state_tensors()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._device_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for t, k in zip(self._state_tensors(), keep):
            t.copy_(k)
        from . import _lib
        n0 = _lib.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._device_step()
        self.kernel_launches_per_step = _lib.launch_count - n0

    def _all_buffers(self):
        mods = [self.model] + ([self.ema_model] if self.ema_model is not None else [])
        return [b for m in mods for b in m.buffers() if b.dtype.is_floating_point]


class ICTTrainer(MeanTeacherTrainer):
    """Interpolation Consistency Training (code/train_interpolation_consistency_training_2D.py:150-193).

    Per iteration: mix factors f ~ Beta(alpha, alpha) for `labeled_bs // 2` pairs of unlabeled samples; the student sees
    [labeled | u0 (1 - f) + u1 f]; the teacher (train mode, no grad) sees u0 and u1 separately and its softmax outputs
    are mixed with the same f; consistency = mean((softmax(student_mixed) - mixed teacher probabilities)^2), no
    `iter < 1000` gate.  The mixed probabilities enter the fused loss kernel as pseudo-logits log(p) (softmax(log p) == p
    because p sums to one), so no new kernel is needed; the mixing itself is a handful of elementwise torch ops on
    [h, C, H, W] tensors inside the same CUDA graph.  Validated on CPU against the oracle (tests/test_host_logic.py)."""

    def __init__(self, model, ema_model, *, batch_size=24, labeled_bs=12, ict_alpha=0.2, mix_seed=None, **kw):
        self.h = labeled_bs // 2
        assert batch_size - labeled_bs == 2 * self.h and self.h > 0, "ICT pairs the unlabeled half: U == 2 * (labeled_bs // 2)"
        assert not kw.get("uncertainty_T"), "ICT has no MC-dropout branch"
        kw.setdefault("consistency_gate_iters", 0)
        self.B_in, self.ict_alpha = batch_size, ict_alpha
        # the schedule below runs the student on Lb + h samples and the teacher on h samples (twice)
        super().__init__(model, ema_model, batch_size=labeled_bs + self.h, labeled_bs=labeled_bs, **kw)
        import numpy as np
        self.mix_rng = np.random.RandomState(mix_seed) if mix_seed is not None else np.random
        dev, h = self.dev, self.h
        self.x_in = torch.empty((self.B_in, 1, *self.patch), dtype=torch.float32, device=dev)
        self.y_in = torch.empty((self.B_in, *self.patch), dtype=self.y.dtype, device=dev)
        self.mix_ring = PinnedRing(h, dev)
        self.mix = torch.zeros((h, 1) + (1,) * len(self.patch), dtype=torch.float32, device=dev)     # 2-D slices or 3-D patches
        self.p0 = torch.empty((h, self.C, self.S), dtype=torch.float32, device=dev)
        self.p1 = torch.empty_like(self.p0)
        self.pseudo_logits = torch.empty_like(self.p0)

    def _device_step(self):
        Lb, h = self.Lb, self.h
        self.s_off += 1
        self.model.train()
        f = self.mix
        u0, u1 = self.x_in[Lb:Lb + h], self.x_in[Lb + h:]
        self.x[:Lb].copy_(self.x_in[:Lb])
        torch.add(u0 * (1.0 - f), u1 * f, out=self.x[Lb:])                         # :163-165
        self.y[:Lb].copy_(self.y_in[:Lb])
        t_rt = self.ema_model._rt
        with t_rt.side_stream():                                                    # :170-176
            for src, dst in ((u0, self.p0), (u1, self.p1)):
                self.t_off += 1
                self.ema_in.copy_(src)
                self.t_plan.forward(self.ema_in, train=True)
                torch.softmax(self.t_plan.logits.view(h, self.C, self.S), dim=1, out=dst)
            f3 = f.view(h, 1, 1)
            torch.log(self.p0 * (1.0 - f3) + self.p1 * f3, out=self.pseudo_logits)
        self.s_plan.forward(self.x, train=True)
        t_rt.join_side()
        w = self.hp[HP_WCONS:HP_WCONS + 1]
        ops.ssl_loss_fwd(self.s_plan.logits, self.pseudo_logits, self.y, False, self.B, Lb, self.C, self.S, w, self.lossbuf,
                         self.loss_ws, None, 0.0, None)
        ops.ssl_loss_bwd(self.s_plan.logits, self.pseudo_logits, self.y, False, self.B, Lb, self.C, self.S, w, self.lossbuf,
                         1.0, self.s_plan.g_logits, True, None, 0.0, None)
        self.s_plan.backward(None)
        if self.world > 1:
            torch.distributed.all_reduce(self.flat.grad, group=self.pg)
        ops.sgd_ema_step(self.flat.data, self.flat.grad, self.momentum_buf, self.ema_flat.data, self.hp)

    def step(self, images, labels, read_loss=False, mix_factors=None):
        """images [batch_size, 1, *patch], labels [batch_size, *patch]; mix_factors: optional [labeled_bs // 2] values
        (default: numpy Beta(ict_alpha, ict_alpha) draws, :155-158)."""
        if mix_factors is None:
            mix_factors = self.mix_rng.beta(self.ict_alpha, self.ict_alpha, size=(self.h,))
        self.mix_ring.push(mix_factors, self.mix.view(self.h))
        self.x_in.copy_(images, non_blocking=True)
        self.y_in.copy_(labels, non_blocking=True)
        self._set_hparams()
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._device_step()
        self.lr = self.base_lr * (1.0 - self.iter_num / self.max_iterations) ** 0.9
        self.iter_num += 1
        if read_loss:
            self.loss_host.copy_(self.lossbuf[:4], non_blocking=True)
            torch.cuda.current_stream().synchronize() if self.dev.type == "cuda" else None
            return self.loss_host.tolist()
        return self.lossbuf


class PinnedRing:
    """Per-step host scalars go to the device through a ring of pinned buffers, each guarded by a CUDA event: the CPU
    runs many graph replays ahead of the GPU, so ONE reused pinned buffer could be overwritten with the values of a
    later iteration before the DMA of an earlier one has read it (lr / EMA alpha / consistency weight of the wrong step)."""

    def __init__(self, n, device, slots=8):
        self.cuda = torch.device(device).type == "cuda"
        self.bufs = [torch.zeros(n, dtype=torch.float32).pin_memory() if self.cuda else torch.zeros(n) for _ in range(slots)]
        self.events = [None] * slots
        self.i = 0

    def push(self, values, dst):
        """values: sequence / tensor of floats -> dst (device tensor) with an async copy on the current stream."""
        k = self.i
        self.i = (k + 1) % len(self.bufs)
        if self.events[k] is not None:
            self.events[k].synchronize()              # the copy that last used this slot has completed
        buf = self.bufs[k]
        buf.copy_(torch.as_tensor(values, dtype=torch.float32).reshape(buf.shape))
        dst.copy_(buf, non_blocking=True)
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record()
            self.events[k] = ev


def _plan_for(model, B, patch, need_grad):
    """UNet-style plans are keyed by (B, H, W), the Swin-UNet's by B alone (its token grid is fixed by its config)."""
    if getattr(model, "plan_key_is_batch", False):
        return model._get_plan(B, need_grad)
    return model._get_plan(B, *patch, need_grad)


class CrossTeachingTrainer(_Pipelined):
    """Cross Teaching between CNN and Transformer (code/train_cross_teaching_between_cnn_transformer_2D.py:221-262).

    Per iteration: both models see the whole batch; each is trained with 0.5 (CE + Dice) on the labeled half plus
    w * Dice against the argmax pseudo labels of the OTHER model on the unlabeled half; two independent SGD steps.
    The reference bumps iter_num before recomputing the poly LR (:257-259), so iteration k runs with poly_lr(k)."""

    def __init__(self, model1, model2, *, batch_size=16, labeled_bs=8, patch_size=(224, 224), num_classes=4, base_lr=0.01,
                 max_iterations=30000, consistency=0.1, consistency_rampup=200.0, momentum=0.9, weight_decay=1e-4,
                 start_iter=0, label_dtype=torch.uint8, process_group=None, use_cuda_graph=False, pseudo_loss="dice"):
        assert pseudo_loss in ("dice", "ce")
        # "dice": Cross Teaching (train_cross_teaching_between_cnn_transformer_2D.py:242-245);
        # "ce":   Cross Pseudo Supervision (train_cross_pseudo_supervision_2D.py:193-196) -- same loop, two CNNs
        self.loss_fwd, self.loss_bwd = (ops.ct_loss_fwd, ops.ct_loss_bwd) if pseudo_loss == "dice" else (ops.cps_loss_fwd, ops.cps_loss_bwd)
        self.models = (model1, model2)
        self.aux = None
        self.B, self.Lb, self.patch, self.C = batch_size, labeled_bs, tuple(patch_size), num_classes
        self.base_lr, self.max_iterations = base_lr, max_iterations
        self.consistency, self.consistency_rampup = consistency, consistency_rampup
        self.momentum, self.weight_decay = momentum, weight_decay
        self.iter_num = start_iter
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.flats = [m.materialize() for m in self.models]
        dev = self.dev = self.flats[0].data.device
        self.momentum_bufs = [torch.zeros_like(f.data) for f in self.flats]
        self.offs = [m._rt.seed_off for m in self.models]
        pin = dev.type == "cuda"
        self.hp_ring = PinnedRing(8, dev)
        self.hp = torch.zeros(8, dtype=torch.float32, device=dev)
        self.S = math.prod(self.patch)             # 2-D slices or 3-D patches (train_cross_pseudo_supervision_3D.py)
        self.x = torch.empty((self.B, 1, *self.patch), dtype=torch.float32, device=dev)
        self.y = torch.empty((self.B, *self.patch), dtype=label_dtype, device=dev)
        # per model: [ce, dice, pseudo-label dice, total, coefficients...]
        self.lossbufs = [torch.zeros(40, dtype=torch.float32, device=dev) for _ in range(2)]
        self.loss_host = torch.zeros(8, dtype=torch.float32).pin_memory() if pin else torch.zeros(8)
        self.loss_ws = torch.empty(ops.ssl_loss_workspace_bytes(self.B, self.S) // 4 + 4, dtype=torch.float32, device=dev)
        self.plans = [_plan_for(m, self.B, self.patch, True) for m in self.models]
        self.lr = base_lr * (1.0 - start_iter / max_iterations) ** 0.9
        self.use_graph = use_cuda_graph and dev.type == "cuda"
        self.graph = None
        self.kernel_launches_per_step = None

    def consistency_weight(self, iter_num):
        """:117-119 with the iter_num // 150 argument of :229-230 (no warm-up gate in this trainer)"""
        return self.consistency * ramps.sigmoid_rampup(iter_num // 150, self.consistency_rampup)

    def _set_hparams(self):
        h = [0.0] * 8
        h[HP_LR] = self.lr
        h[HP_MOMENTUM] = self.momentum
        h[HP_WD] = self.weight_decay
        h[HP_ALPHA], h[HP_ONE_MINUS_ALPHA] = 1.0, 0.0
        h[HP_GRAD_SCALE] = 1.0 / self.world
        h[HP_WCONS] = self.consistency_weight(self.iter_num)
        self.hp_ring.push(h, self.hp)

    def _device_step(self):
        for m, off in zip(self.models, self.offs):
            off += 1
            m.train()
        p1, p2 = self.plans
        # the two networks are independent except at the loss: model 1 runs on an auxiliary stream next to model 2
        overlap = self.dev.type == "cuda"
        if overlap and self.aux is None:
            self.aux = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream() if overlap else None

        def on_aux(fn):
            if not overlap:
                return fn()
            self.aux.wait_stream(main)
            with torch.cuda.stream(self.aux):
                fn()

        on_aux(lambda: p1.forward(self.x, train=True))                     # :224-228
        p2.forward(self.x, train=True)
        if overlap:
            main.wait_stream(self.aux)
        w = self.hp[HP_WCONS:HP_WCONS + 1]
        for mine, other, lb in ((p1, p2, self.lossbufs[0]), (p2, p1, self.lossbufs[1])):
            self.loss_fwd(mine.logits, False, other.logits, False, self.y, self.B, self.Lb, self.C, self.S, w, lb, self.loss_ws)
            self.loss_bwd(mine.logits, False, other.logits, False, self.y, self.B, self.Lb, self.C, self.S, lb, 1.0,
                          mine.g_logits, True)

        def update(i):
            self.plans[i].backward(None)                                   # :252-255 (loss = model1_loss + model2_loss)
            if self.world == 1:
                ops.sgd_ema_step(self.flats[i].data, self.flats[i].grad, self.momentum_bufs[i], None, self.hp)     # :257-258

        on_aux(lambda: update(0))
        update(1)
        if self.world > 1:
            # one communicator: its collectives stay on ONE stream, in the same order on every rank -- model 2's
            # gradient goes first while model 1's backward is still running on the auxiliary stream
            for i in (1, 0):
                if i == 0 and overlap:
                    main.wait_stream(self.aux)
                torch.distributed.all_reduce(self.flats[i].grad, group=self.pg)
                ops.sgd_ema_step(self.flats[i].data, self.flats[i].grad, self.momentum_bufs[i], None, self.hp)
        elif overlap:
            main.wait_stream(self.aux)

    def step(self, images, labels, read_loss=False):
        """images [B,1,H,W] float32, labels [B,H,W] uint8.  Returns the two device loss buffers, or with read_loss the
        host floats [ce1, dice1, pseudo1, model1_loss, ce2, dice2, pseudo2, model2_loss]."""
        self._set_hparams()
        self.x.copy_(images, non_blocking=True)
        self.y.copy_(labels, non_blocking=True)
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._device_step()
        self.iter_num += 1                                                 # :257
        self.lr = self.base_lr * (1.0 - self.iter_num / self.max_iterations) ** 0.9    # :259
        if read_loss:
            self.loss_host[:4].copy_(self.lossbufs[0][:4], non_blocking=True)
            self.loss_host[4:].copy_(self.lossbufs[1][:4], non_blocking=True)
            torch.cuda.current_stream().synchronize() if self.dev.type == "cuda" else None
            return self.loss_host.tolist()
        return self.lossbufs

    def _loss_sources(self):
        return self.lossbufs

    def _state_tensors(self):
        ts = [f.data for f in self.flats] + self.momentum_bufs + self.offs
        return ts + [b for m in self.models for b in m.buffers() if b.dtype.is_floating_point]

    def _capture(self):
        keep = [t.clone() for t in self._state_tensors()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._device_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for t, k in zip(self._state_tensors(), keep):
            t.copy_(k)
        from . import _lib
        n0 = _lib.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._device_step()
        self.kernel_launches_per_step = _lib.launch_count - n0
