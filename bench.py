"""Benchmark of the hot path: the semi-supervised training step of BASELINE.json's configs on synthetic tensors.

    python bench.py --gpus N --steps K --warmup W [--config 2|3|4|5]   # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W [--config C] # the reference's CPU path (oracle port) on host cores

--config 2 (default, the configuration BASELINE.json's metric is quoted on): 2D UNet Mean-Teacher, 256x256, bs24 (12/12)
--config 3: Cross-Teaching UNet <-> Swin-UNet, 224x224, bs16 (8/8)
--config 4: 3D VNet uncertainty-aware Mean-Teacher, 96^3, bs4 (2/2), T = 8 MC-dropout passes
--config 5: 3D UNETR (ViT-B encoder) fully supervised, 96^3, bs2

Prints ONE JSON line (rank 0).  `value`: units/s with the batch already resident in HBM (CUDA-graph replay of the whole
step); `e2e`: the same step driven through the public API with pinned HOST batches (H2D copy and loss read-back inside
the timed region); `roofline`: the dominant C-ABI call of the step, timed live with CUDA events in an eager profiling
pass (algorithmic bytes / flops from cv_ssl_mis_b200/_costs.py); `kernels`: the same for every entry point;
`cpu_baseline`: the oracle port of the reference step timed on this box's host cores; `gpu_eager_baseline`: the same
oracle port moved to the GPU (functional torch eager over cuDNN / cuBLAS, cuDNN TF32 convolutions as PyTorch defaults) --
what a user of the reference gets on this B200 today.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def blocky(g, B, C, shape, blk):
    """Blocky labels (nearest-upsampled random map): non-degenerate Dice terms."""
    low = torch.randint(0, C, (B, *[max(1, s // blk) for s in shape]), generator=g, device="cpu")
    for d in range(len(shape)):
        low = low.repeat_interleave(blk, d + 1)
    return low[(slice(None),) + tuple(slice(0, s) for s in shape)].contiguous()


# ============================================================================================== workloads
class Workload:
    key = 0
    unit = "slices/s"
    B = Lb = 0
    patch = ()
    ncls = 4

    def batch(self, seed, pinned):
        g = torch.Generator().manual_seed(seed)
        if len(self.patch) == 2:
            x = torch.rand(self.B, 1, *self.patch, generator=g)                  # min-max normalised MRI slice in [0,1]
            y = blocky(g, self.B, self.ncls, self.patch, 16).to(torch.uint8)
        else:
            x = torch.randn(self.B, 1, *self.patch, generator=g)                 # z-scored volume
            y = blocky(g, self.B, self.ncls, self.patch, 8).to(torch.int64)
        return (x.pin_memory(), y.pin_memory()) if pinned else (x, y)

    def h2d_bytes(self):
        n = self.B
        for s in self.patch:
            n *= s
        return n * 4 + n * (1 if len(self.patch) == 2 else 8) + 32

    def runtimes(self, tr):
        ms = [m for m in (getattr(tr, "models", None) or [tr.model, getattr(tr, "ema_model", None)]) if m is not None]
        return [m._rt for m in ms]

    def losses(self, tr):
        return tr.lossbuf[:4].tolist()


class MT2D(Workload):
    key = 2
    B, Lb, patch, ncls = 24, 12, (256, 256), 4
    workload = "configs[1]: 2D UNet Mean-Teacher, ACDC-shape 256x256, 4 classes, bs24 per GPU (12 lab/12 unlab)"
    metric = "train-step slices/sec (ACDC 256x256 bs24 MT-UNet)"
    notes = {"iter_num": "1000+ (consistency term live)",
             "l2": "per-step working set (~6 GB of activations) >> 126 MB L2; no explicit flush needed"}

    def build(self, rank, pg, graph):
        from cv_ssl_mis_b200.networks.unet import UNet
        from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
        torch.manual_seed(1337)                      # identical student/teacher init on every rank (reference seed)
        student, teacher = UNet(1, self.ncls, seed=1337 + rank).cuda(), UNet(1, self.ncls, seed=7331 + rank).cuda()
        for p in teacher.parameters():
            p.detach_()
        return MeanTeacherTrainer(student, teacher, batch_size=self.B, labeled_bs=self.Lb, patch_size=self.patch,
                                  num_classes=self.ncls, start_iter=1000, noise_seed=99 + rank, process_group=pg, use_cuda_graph=graph)

    def flats(self, tr):
        return [tr.flat.data, tr.ema_flat.data]

    def oracle_step(self, batch, dev, g):
        """The reference step (code/train_mean_teacher_2D.py:204-236) through the oracle port on `dev`."""
        from oracle import ssl_oracle as O
        from cv_ssl_mis_b200.networks.unet import UNet
        H, W = self.patch
        torch.manual_seed(1337)
        s_sd = {k: v.clone().to(dev) for k, v in UNet(1, self.ncls).state_dict().items()}
        t_sd = {k: v.clone().to(dev) for k, v in UNet(1, self.ncls).state_dict().items()}
        bufs = {k: torch.zeros_like(s_sd[k]) for k in O.param_keys(s_sd)}
        lb = batch // 2
        x = torch.rand(batch, 1, H, W, generator=g, device="cpu").to(dev)
        y = blocky(g, batch, self.ncls, self.patch, 16).to(torch.uint8).to(dev)
        state = {"it": 1000}

        def masks(b):      # nn.Dropout is active in the reference (train mode): fresh masks every step like it draws them
            return [(torch.rand(b, c, H >> i, W >> i, device=dev) >= p).float() for i, (c, p) in enumerate(zip(O.UNET_FT, O.UNET_DROPOUT))]

        def step():
            noise = torch.clamp(torch.randn(batch - lb, 1, H, W, device=dev) * 0.1, -0.2, 0.2)
            r = O.mt2d_step(s_sd, t_sd, bufs, x, y, noise, state["it"], labeled_bs=lb, n_classes=self.ncls,
                            student_masks=masks(batch), teacher_masks=masks(batch - lb))
            state["it"] += 1
            return r["loss"]
        return step


class CT2D(Workload):
    key = 3
    B, Lb, patch, ncls = 16, 8, (224, 224), 4
    workload = "configs[2]: 2D Cross-Teaching UNet<->Swin-UNet (tiny-lite), 224x224, 4 classes, bs16 per GPU (8 lab/8 unlab)"
    metric = "train-step slices/sec (ACDC 224x224 bs16 Cross-Teaching UNet<->SwinUNet)"
    notes = {"iter_num": "3000+", "l2": "per-step working set >> 126 MB L2; no explicit flush needed"}

    def build(self, rank, pg, graph):
        from cv_ssl_mis_b200.networks.net_factory import net_factory
        from cv_ssl_mis_b200.trainers import CrossTeachingTrainer
        torch.manual_seed(1337)
        m1, m2 = net_factory("unet", 1, self.ncls), net_factory("ViT_Seg", 1, self.ncls)
        return CrossTeachingTrainer(m1, m2, batch_size=self.B, labeled_bs=self.Lb, patch_size=self.patch, num_classes=self.ncls,
                                    start_iter=3000, process_group=pg, use_cuda_graph=graph)

    def flats(self, tr):
        return [f.data for f in tr.flats]

    def losses(self, tr):
        return [b[:4].tolist() for b in tr.lossbufs]

    def oracle_step(self, batch, dev, g):
        from oracle import ssl_oracle as O, swin_oracle as SO
        from cv_ssl_mis_b200.networks.unet import UNet
        from cv_ssl_mis_b200.networks.swin_unet import SwinUnet
        torch.manual_seed(1337)
        sd1 = {k: v.clone().to(dev) for k, v in UNet(1, self.ncls).state_dict().items()}
        sd2 = {k[len("swin_unet."):]: v.clone().to(dev) for k, v in SwinUnet(None, num_classes=self.ncls).state_dict().items()}
        cfg = SO.swin_config(sd2, 224, 7, 0.2)
        bufs1 = {k: torch.zeros_like(sd1[k]) for k in O.param_keys(sd1)}
        bufs2 = {k: torch.zeros_like(v) for k, v in sd2.items() if v.dtype.is_floating_point}
        x = torch.rand(batch, 1, *self.patch, generator=g, device="cpu").to(dev)
        y = blocky(g, batch, self.ncls, self.patch, 8).to(torch.uint8).to(dev)
        state = {"it": 3000}

        def step():
            keeps = [tuple((torch.rand(batch, device=dev) >= p).float() for _ in range(2)) for p in cfg["dpr"] + cfg["dpr"][:6]]
            r = SO.ct2d_step(sd1, sd2, bufs1, bufs2, x.clone(), y.clone(), state["it"], cfg, labeled_bs=batch // 2, drop_keep=keeps)
            state["it"] += 1
            return r["loss"] if isinstance(r, dict) and "loss" in r else None
        return step


class UAMT3D(Workload):
    key = 4
    unit = "patches/s"
    B, Lb, patch, ncls, T = 4, 2, (96, 96, 96), 2, 8
    workload = "configs[3]: 3D VNet uncertainty-aware Mean-Teacher, BraTS-shape 96^3, 2 classes, bs4 per GPU (2 lab/2 unlab), T=8"
    metric = "train-step patches/sec (BraTS 96^3 bs4 UAMT-VNet, T=8)"
    notes = {"iter_num": "3000+", "l2": "per-step working set (~10 GB) >> 126 MB L2; no explicit flush needed"}

    def build(self, rank, pg, graph):
        from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
        from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
        torch.manual_seed(1337)
        s, t = net_factory_3d("vnet", 1, self.ncls), net_factory_3d("vnet", 1, self.ncls)
        return MeanTeacherTrainer(s, t, batch_size=self.B, labeled_bs=self.Lb, patch_size=self.patch, num_classes=self.ncls,
                                  start_iter=3000, consistency_gate_iters=0, uncertainty_T=self.T, noise_seed=99 + rank,
                                  process_group=pg, use_cuda_graph=graph)

    def flats(self, tr):
        return [tr.flat.data, tr.ema_flat.data]

    def oracle_step(self, batch, dev, g):
        from oracle import ssl_oracle as O
        from cv_ssl_mis_b200.networks.vnet import VNet
        P, T = self.patch[0], self.T
        torch.manual_seed(1337)
        s_sd = {k: v.clone().to(dev) for k, v in VNet(1, self.ncls, has_dropout=True).state_dict().items()}
        t_sd = {k: v.clone().to(dev) for k, v in VNet(1, self.ncls, has_dropout=True).state_dict().items()}
        bufs = {k: torch.zeros_like(s_sd[k]) for k in O.param_keys(s_sd)}
        x = torch.randn(batch, 1, P, P, P, generator=g, device="cpu").to(dev)
        y = blocky(g, batch, self.ncls, self.patch, 8).to(torch.int64).to(dev)
        Lb = batch // 2
        U = batch - Lb
        state = {"it": 3000}

        def step():
            noises = [torch.clamp(torch.randn(U if k == 0 else 2 * U, 1, P, P, P, device=dev) * 0.1, -0.2, 0.2) for k in range(1 + T // 2)]
            drops = lambda n: ((torch.rand(n, 256, device=dev) >= 0.5).float(), (torch.rand(n, 16, device=dev) >= 0.5).float())
            r = O.uamt3d_step(s_sd, t_sd, bufs, x, y, noises, state["it"], labeled_bs=Lb, T=T, student_drops=drops(batch),
                              teacher_drops=[drops(U if k == 0 else 2 * U) for k in range(1 + T // 2)])
            state["it"] += 1
            return r["loss"]
        return step


class UNETR3D(Workload):
    key = 5
    unit = "patches/s"
    B, Lb, patch, ncls = 2, 2, (96, 96, 96), 2
    workload = "configs[4]: 3D UNETR (ViT-B encoder, 92.8 M parameters) fully supervised, 96^3, 2 classes, bs2 per GPU"
    metric = "train-step patches/sec (BraTS 96^3 bs2 fully-supervised UNETR)"
    notes = {"l2": "weights (371 MB) + activations >> 126 MB L2; no explicit flush needed",
             "oracle": "parity unpinned against MONAI itself (absent here); the restated oracle is cross-checked against torch.nn blocks"}

    def build(self, rank, pg, graph):
        from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
        from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
        torch.manual_seed(1337)
        net = net_factory_3d("unetr", 1, self.ncls)
        return MeanTeacherTrainer(net, None, batch_size=self.B, labeled_bs=self.B, patch_size=self.patch, num_classes=self.ncls,
                                  process_group=pg, use_cuda_graph=graph)

    def flats(self, tr):
        return [tr.flat.data]

    def oracle_step(self, batch, dev, g):
        from oracle import unetr_oracle as UO
        from cv_ssl_mis_b200.networks.unetr import UNETR
        P = self.patch[0]
        torch.manual_seed(1337)
        net = UNETR(in_channels=1, out_channels=self.ncls, img_size=(P, P, P), feature_size=16, hidden_size=768, mlp_dim=3072,
                    num_heads=12, pos_embed="perceptron", norm_name="instance", conv_block=True, res_block=True, dropout_rate=0.0)
        sd = {k: v.detach().clone().to(dev).requires_grad_(v.dtype.is_floating_point) for k, v in net.state_dict().items()}
        params = [v for v in sd.values() if v.requires_grad]
        mom = [torch.zeros_like(p) for p in params]
        x = torch.randn(batch, 1, P, P, P, generator=g, device="cpu").to(dev)
        y = blocky(g, batch, self.ncls, self.patch, 8).to(torch.int64).to(dev)

        def step():
            loss, _ = UO.fully_supervised_loss(sd, x, y, 12, self.ncls)
            grads = torch.autograd.grad(loss, params, allow_unused=True)      # (the last ViT block's output is not tapped)
            with torch.no_grad():                                 # SGD(momentum 0.9, weight decay 1e-4), code/train_fully_supervised_3D_ViT.py
                for p, gr, m in zip(params, grads, mom):
                    if gr is None:
                        continue
                    m.mul_(0.9).add_(gr + 1e-4 * p)
                    p.sub_(0.01 * m)
            return loss.detach()
        return step


WORKLOADS = {w.key: w for w in (MT2D(), CT2D(), UAMT3D(), UNETR3D())}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def rows(self):
        try:
            return sum(1 for _ in open(self.f.name))
        except OSError:
            return 0

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
def time_cpu(wl, batch, steps, warmup, threads):
    torch.set_num_threads(threads)
    step = wl.oracle_step(batch, torch.device("cpu"), torch.Generator().manual_seed(1337))
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


# seconds per full-batch CPU step on ~16 host cores (measured on this pool): sizes the bounded sample of the reference arm
CPU_SEC_PER_STEP = {2: 1.3, 3: 1.2, 4: 3.6, 5: 1.2}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.config]
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    est = CPU_SEC_PER_STEP[wl.key] * 16.0 / max(cores, 1)
    batch = wl.B if total * est <= 300 else max(2, int(wl.B * 300 / (total * est)) // 2 * 2)
    ts = time_cpu(wl, batch, args.steps, args.warmup, cores)
    sec = sum(ts) / len(ts)
    value = batch / sec
    sample = (f"bs{batch} {'x'.join(map(str, wl.patch))} per step, oracle port of the reference step "
              f"(torch CPU fp32, {cores} threads)")
    print(json.dumps({
        "impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": wl.workload, "sample": sample},
        "cpu_baseline": {"value": value, "unit": wl.unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def time_gpu_eager(wl, steps=5, warmup=2):
    """The oracle port of the reference step on the GPU: functional torch eager over cuDNN / cuBLAS with PyTorch's default
    numerics for it (cuDNN convolutions in TF32, fp32 matmuls), cudnn.benchmark on (its best algorithms), CUDA-event timed
    with the batch resident in HBM -- the number a user of the reference gets on this GPU today."""
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda")
    with torch.device("cuda"):              # the oracle's index / mask helpers allocate on the default device
        step = wl.oracle_step(wl.B, dev, torch.Generator().manual_seed(1337))
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": wl.B / ms * 1e3, "unit": wl.unit, "ms_per_step": ms, "steps": steps,
            "what": "oracle port of the reference step, torch eager on this GPU (cuDNN TF32 convolutions, fp32 matmuls: PyTorch "
                    "defaults; cudnn.benchmark=True), batch resident in HBM"}


# ----------------------------------------------------------------------------------------------- our arm
def profile_pass(wl, tr, x, y, reps=3):
    from cv_ssl_mis_b200 import _lib
    graph, tr.use_graph = tr.use_graph, False
    # kernels are timed alone: the side streams (weight gradients, teacher passes) are folded back into the main one
    rts = wl.runtimes(tr)
    overlap = [rt.overlap for rt in rts]
    for rt in rts:
        rt.overlap = False
    tr.step(x, y)
    torch.cuda.synchronize()
    _lib.profile = []
    for _ in range(reps):
        _lib.tag = ""
        tr.step(x, y)
    torch.cuda.synchronize()
    rec, _lib.profile = _lib.profile, None
    tr.use_graph = graph
    for rt, ov in zip(rts, overlap):
        rt.overlap = ov
    table = {}                                # entry -> [ms, bytes, flops, calls, calls with a cost model]
    for name, tag, e0, e1, cost in rec:
        e = table.setdefault(name, [0.0, 0.0, 0.0, 0, 0])
        e[0] += e0.elapsed_time(e1) / reps
        e[3] += 1
        if cost is not None:
            e[1] += cost[0] / reps; e[2] += cost[1] / reps; e[4] += 1
    for e in table.values():
        e[3] /= reps; e[4] /= reps
    return table, sum(e[0] for e in table.values())


def run_ours(args):
    import torch.distributed as dist
    from cv_ssl_mis_b200 import _lib

    wl = WORKLOADS[args.config]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the single JSON line
    pg = None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
        pg = dist.group.WORLD
    tr = wl.build(rank, pg, not args.no_graph)
    if world > 1:
        for f in wl.flats(tr):
            dist.broadcast(f, 0)
    host = [wl.batch(1337 + rank * 100 + i, True) for i in range(4)]
    xd, yd = host[0][0].cuda(), host[0][1].cuda()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # nvidia-smi needs ~100 ms before its first sample: start it before the warm-up and keep it running through both
    # timed regions (the GPU is under the same load from here to sampler.stop())
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        tr.step(xd, yd)
    n0 = _lib.launch_count
    ms = timed(lambda i: tr.step(xd, yd), args.steps)
    launches = tr.kernel_launches_per_step * args.steps if tr.kernel_launches_per_step else _lib.launch_count - n0
    loss_dev = wl.losses(tr)

    # end to end through the public API: every step uploads ITS pinned host batch and reads ITS losses back.
    # `submit` is the pipelined form (upload of step i+1 on a copy stream under step i; the loss of step i is read after
    # step i+1 has been enqueued); the blocking `step(..., read_loss=True)` form is timed next to it.
    for i in range(2):
        tr.step(*host[i % 4], read_loss=True)
    ms_e2e_sync = timed(lambda i: tr.step(*host[i % 4], read_loss=True), args.steps)
    pending, e2e_losses = [], []

    def e2e_step(i):
        pending.append(tr.submit(*host[i % 4]))
        if len(pending) > 1:
            e2e_losses.append(pending.pop(0).result())

    def e2e_drain():
        while pending:
            e2e_losses.append(pending.pop(0).result())

    for i in range(2):
        e2e_step(i)
    e2e_drain()
    e2e_losses.clear()
    ms_e2e = timed(e2e_step, args.steps, e2e_drain)
    assert len(e2e_losses) == args.steps and all(math.isfinite(v) for l in e2e_losses for v in l)
    clocks = None
    if sampler:
        # short runs end before nvidia-smi has printed enough rows: hold the same load (untimed) until it has
        t_end = time.time() + 3.0
        while world == 1 and sampler.rows() < 5 and time.time() < t_end:      # (collective inside the step: 1 rank only)
            tr.step(xd, yd)
            torch.cuda.synchronize()
        clocks = sampler.stop()

    out = None
    # every rank runs the eager profiling pass (its steps contain the gradient all-reduce); rank 0 reports
    table, total = profile_pass(wl, tr, xd, yd)
    if rank == 0:
        pk = peaks()
        ridge = pk["tf_sustained"] * 1e12 / (pk["hbm"] * 1e9)
        kernels = []
        for name, (ms_k, by_k, fl_k, calls, modelled) in sorted(table.items(), key=lambda kv: -kv[1][0]):
            row = {"entry": name, "launches_per_step": round(calls, 1), "ms_per_step": round(ms_k, 4)}
            if modelled and modelled == calls and by_k > 0:
                gbs, tfs = by_k / (ms_k * 1e-3) / 1e9, fl_k / (ms_k * 1e-3) / 1e12
                bound = "tensor" if (fl_k and fl_k / by_k > ridge) else "hbm"
                row.update({"GB/s": round(gbs, 1), "TFLOP/s": round(tfs, 1), "bound": bound,
                            "frac": round(tfs / pk["tf_sustained"] if bound == "tensor" else gbs / pk["hbm"], 3)})
                if bound == "tensor":  # the path computes in TF32, whose tensor-pipe rate is half the dense bf16 peak used above
                    row["frac_of_tf32_rate"] = round(2 * tfs / pk["tf_sustained"], 3)
            kernels.append(row)
        # dominant kernel = the C-ABI entry point with the largest share of the step; its launches are summed:
        # achieved = (algorithmic bytes or flops over all its launches in one step) / (their summed device time)
        roof = None
        top = next((k for k in kernels if "bound" in k), None)
        if top is not None:
            ms_k, by_k, fl_k, calls, _ = table[top["entry"]]
            if top["bound"] == "tensor":
                ach = fl_k / (ms_k * 1e-3) / 1e12
                roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                        "frac_of_tf32_rate": 2 * ach / pk["tf_sustained"], "traffic": None}
            else:
                ach = by_k / (ms_k * 1e-3) / 1e9
                roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None}
            roof.update({"kernel": top["entry"], "launches_per_step": calls, "ms_per_launch": ms_k / calls,
                         "algorithmic_bytes_per_launch": by_k / calls, "algorithmic_flops_per_launch": fl_k / calls,
                         "achieved_tflops": fl_k / (ms_k * 1e-3) / 1e12,
                         "peak_source": pk["src"] + " (MEASURED_PEAKS.json: HBM copy GB/s; dense bf16 sustained TF/s for tensor)",
                         "share_of_step": ms_k / total})
            # dram__bytes_read + dram__bytes_write per launch of that kernel, from the committed ncu --set full capture
            tfile = os.path.join(ROOT, "profiles", f"r2_traffic_config{wl.key}.json")
            if os.path.exists(tfile):
                with open(tfile) as f:
                    tr_ = json.load(f)
                if tr_.get("kernel") == top["entry"]:
                    roof["traffic"] = tr_["dram_bytes_per_launch"]
                    roof["traffic_source"] = tr_["source"]
        shares = {k["entry"]: round(k["ms_per_step"] / total, 4) for k in kernels[:8]}
        value = wl.B * world * args.steps / (ms * 1e-3)
        e2e_v = wl.B * world * args.steps / (ms_e2e * 1e-3)
        cfg = {"workload": wl.workload, "cuda_graph": tr.use_graph,
               "numerics": "fp32 storage, TF32 tensor-core products, fp32 accumulation (the class cuDNN runs for the reference)"}
        cfg.update(wl.notes)
        out = {
            "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32", "data": "synthetic", "config": cfg,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_v, "unit": wl.unit, "ms_per_step": ms_e2e / args.steps, "api": "trainer.submit(pinned batch) -> PendingLoss.result(), one step in flight",
                    "blocking_api_ms_per_step": ms_e2e_sync / args.steps,
                    "h2d_bytes_per_step": wl.h2d_bytes(), "d2h_bytes_per_step": 16 * (2 if wl.key == 3 else 1)},
            "roofline": roof, "kernels": kernels, "step_time_shares": shares, "profiled_eager_ms_per_step": total, "loss": loss_dev,
        }
    if world == 1 and not args.no_cpu:
        # free our buffers first: the eager baseline of the 3-D configs needs several GB of its own
        del tr
        torch.cuda.empty_cache()
        if not args.no_eager:
            try:
                out["gpu_eager_baseline"] = time_gpu_eager(wl)
                out["gpu_eager_baseline"]["ours_over_eager"] = out["value"] / out["gpu_eager_baseline"]["value"]
            except Exception as e:          # a baseline, not the product: report instead of failing the bench line
                out["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
        cores = os.cpu_count() or 1
        ts = time_cpu(wl, wl.B, 2, 1, cores)
        sec = sum(ts) / len(ts)
        out["cpu_baseline"] = {"value": wl.B / sec, "unit": wl.unit, "cores": cores, "kind": "port",
                               "sample": f"2 timed + 1 warm-up full bs{wl.B} steps of the oracle port (torch CPU fp32)"}
        tr = None
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        # a captured graph keeps NCCL work alive and destroy_process_group() then blocks (observed on this pool):
        # drop the graph, drain the device, synchronise the ranks and leave without the communicator teardown
        tr.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS), help="BASELINE.json configs[N-1]")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and gpu_eager_baseline legs")
    ap.add_argument("--no-eager", action="store_true", help="skip the gpu_eager_baseline leg")
    args = ap.parse_args()
    if os.environ.get("B200_FAULT"):          # debugging aid: dump all Python stacks if the run is still alive after N s
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["B200_FAULT"]), exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
