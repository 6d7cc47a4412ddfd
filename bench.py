"""Benchmark of the hot path: Mean-Teacher 2D UNet training step, synthetic ACDC-shape 256x256, bs 24 (12/12).

    python bench.py --gpus N --steps K --warmup W           # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W   # the reference's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  `value`: slices/s with the batch already resident in HBM (CUDA-graph replay of
the whole step); `e2e`: the same step driven through the public API with pinned HOST batches (H2D copy and
loss read-back inside the timed region); `roofline`: the dominant C-ABI call of the step, timed live with CUDA
events in an eager profiling pass; `cpu_baseline`: the oracle port of the reference step timed on this box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B, LB, H, W, NCLS = 24, 12, 256, 256, 4
WORKLOAD = "configs[1]: 2D UNet Mean-Teacher, ACDC-shape 256x256, 4 classes, bs24 per GPU (12 lab/12 unlab)"
METRIC = "train-step slices/sec (ACDC 256x256 bs24 MT-UNet)"
UNIT = "slices/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def synth_batch(seed, pinned):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 1, H, W, generator=g)                               # min-max normalised MRI slice in [0,1]
    low = torch.randint(0, NCLS, (B, H // 16, W // 16), generator=g)      # blocky labels: non-degenerate Dice terms
    y = low.repeat_interleave(16, 1).repeat_interleave(16, 2).to(torch.uint8)
    return (x.pin_memory(), y.pin_memory()) if pinned else (x, y)


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
def oracle_step_runner(batch, threads):
    """The reference step (code/train_mean_teacher_2D.py:204-236) through the oracle port on CPU."""
    from oracle import ssl_oracle as O
    from cv_ssl_mis_b200.networks.unet import UNet
    torch.set_num_threads(threads)
    torch.manual_seed(1337)
    s_sd = {k: v.clone() for k, v in UNet(1, NCLS).state_dict().items()}
    t_sd = {k: v.clone() for k, v in UNet(1, NCLS).state_dict().items()}
    bufs = {k: torch.zeros_like(s_sd[k]) for k in O.param_keys(s_sd)}
    lb = batch // 2
    g = torch.Generator().manual_seed(1337)
    x = torch.rand(batch, 1, H, W, generator=g)
    y = torch.randint(0, NCLS, (batch, H // 16, W // 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2).to(torch.uint8)
    state = {"it": 1000}

    def masks(b):      # nn.Dropout is active in the reference (train mode): draw fresh masks like it does
        return [(torch.rand(b, c, H >> i, W >> i, generator=g) >= p).float() for i, (c, p) in enumerate(zip(O.UNET_FT, O.UNET_DROPOUT))]

    def step():
        noise = O.clamp_noise(x[lb:], g)
        O.mt2d_step(s_sd, t_sd, bufs, x, y, noise, state["it"], labeled_bs=lb, n_classes=NCLS,
                    student_masks=masks(batch), teacher_masks=masks(batch - lb))
        state["it"] += 1
    return step


def time_cpu(batch, steps, warmup, threads):
    step = oracle_step_runner(batch, threads)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    est_full = 10.0                                     # s per full bs24 step on ~8 cores
    batch = B if total * est_full <= 300 else max(2, int(B * 300 / (total * est_full)) // 2 * 2)
    ts = time_cpu(batch, args.steps, args.warmup, cores)
    sec = sum(ts) / len(ts)
    value = batch / sec
    sample = f"bs{batch} ({batch // 2} labeled/{batch - batch // 2} unlabeled) 256x256 per step, oracle port of the reference step, torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------- our arm
def conv_cost(layer, op):
    """algorithmic bytes / flops of one conv-layer call (each operand read once, result written once)."""
    d = layer.desc
    m_in = d.n * d.id * d.ih * d.iw
    cin = d.c0 + d.c1
    K = layer.T * cin
    by = 4 * (m_in * cin + layer.M * layer.cout + K * layer.cout)
    fl = 2.0 * layer.M * K * layer.cout
    if op in ("b200_bn_stats_fwd",):
        return 4 * layer.M * layer.cout, 0.0
    if op == "b200_bn_act_fwd":
        return 8 * layer.M * layer.cout, 0.0
    if op == "b200_bn_act_bwd":
        return 20 * layer.M * layer.cout, 0.0
    return by, fl


def profile_pass(tr, x, y, reps=3):
    from cv_ssl_mis_b200 import _lib
    graph, tr.use_graph = tr.use_graph, False
    # kernels are timed alone: the side streams (weight gradients, teacher passes) are folded back into the main one
    rts = [m._rt for m in (tr.model, tr.ema_model) if m is not None]
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
    agg = {}
    for name, tag, e0, e1 in rec:
        k = (name, tag)
        agg.setdefault(k, []).append(e0.elapsed_time(e1))
    rows = [(k, sum(v) / len(v) * len(v) / reps, len(v) / reps) for k, v in agg.items()]   # ms per step, calls per step
    total = sum(r[1] for r in rows)
    by_name = {}
    for (name, tag), ms, calls in rows:
        by_name[name] = by_name.get(name, 0.0) + ms
    return rows, by_name, total


def run_ours(args):
    import torch.distributed as dist
    from cv_ssl_mis_b200 import _lib
    from cv_ssl_mis_b200.networks.unet import UNet
    from cv_ssl_mis_b200.trainers import MeanTeacherTrainer

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the single JSON line
    pg = None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
        pg = dist.group.WORLD
    torch.manual_seed(1337)                      # identical student/teacher init on every rank (reference seed)
    student, teacher = UNet(1, NCLS, seed=1337 + rank).cuda(), UNet(1, NCLS, seed=7331 + rank).cuda()
    for p in teacher.parameters():
        p.detach_()
    tr = MeanTeacherTrainer(student, teacher, batch_size=B, labeled_bs=LB, patch_size=(H, W), num_classes=NCLS,
                            start_iter=1000, noise_seed=99 + rank, process_group=pg, use_cuda_graph=not args.no_graph)
    if world > 1:
        dist.broadcast(tr.flat.data, 0)
        dist.broadcast(tr.ema_flat.data, 0)
    host = [synth_batch(1337 + rank * 100 + i, True) for i in range(4)]
    xd, yd = host[0][0].cuda(), host[0][1].cuda()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
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
    loss_dev = tr.lossbuf[:4].tolist()

    for i in range(2):
        tr.step(*host[i % 4], read_loss=True)
    ms_e2e = timed(lambda i: tr.step(*host[i % 4], read_loss=True), args.steps)
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
    rows, by_name, total = profile_pass(tr, xd, yd)
    if rank == 0:
        pk = peaks()
        layers = {l.name: l for l in tr.s_plan.layers + tr.t_plan.layers}
        # dominant kernel = the C-ABI entry point with the largest share of the step; its launches are summed:
        # achieved = (algorithmic bytes or flops over all its launches in one step) / (their summed device time)
        top_name = max(by_name, key=by_name.get)
        t_ms = by_b = by_f = n_calls = 0.0
        per_layer = []
        for (name, tag), ms_step, calls in rows:
            if name != top_name or tag not in layers:
                continue
            by, fl = conv_cost(layers[tag], name)
            t_ms += ms_step; by_b += by * calls; by_f += fl * calls; n_calls += calls
            per_layer.append((tag, round(ms_step / calls * 1e3, 1), round(by / (ms_step / calls * 1e-3) / 1e9, 1),
                              round(fl / (ms_step / calls * 1e-3) / 1e12, 1)))
        roof = None
        if t_ms > 0:
            ridge = pk["tf_sustained"] * 1e12 / (pk["hbm"] * 1e9)
            if by_f and by_f / by_b > ridge:
                ach = by_f / (t_ms * 1e-3) / 1e12
                roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                        "frac": ach / pk["tf_sustained"], "traffic": None}
            else:
                ach = by_b / (t_ms * 1e-3) / 1e9
                roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None}
            roof.update({"kernel": top_name, "launches_per_step": n_calls, "ms_per_launch": t_ms / n_calls,
                         "algorithmic_bytes_per_launch": by_b / n_calls, "algorithmic_flops_per_launch": by_f / n_calls,
                         "achieved_tflops": by_f / (t_ms * 1e-3) / 1e12,
                         "peak_source": pk["src"] + " (MEASURED_PEAKS.json: HBM copy GB/s; dense bf16 sustained TF/s for tensor)",
                         "share_of_step": t_ms / total,
                         "per_layer_us_GBs_TFs": sorted(per_layer, key=lambda r: -r[1])[:6]})
        # dram__bytes_read + dram__bytes_write per launch of that kernel, from the committed ncu --set full capture
        tfile = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")
        if roof is not None and os.path.exists(tfile):
            with open(tfile) as f:
                tr_ = json.load(f)
            if tr_.get("kernel") == top_name:
                roof["traffic"] = tr_["dram_bytes_per_launch"]
                roof["traffic_source"] = tr_["source"]
        # per entry point: algorithmic bytes / flops of all its launches in one step over their summed device time
        # (eager pass, CUDA events on the launching stream; small kernels carry ~2 us of event granularity each)
        Sx, U_ = H * W, B - B // 2
        extra = {"b200_sgd_ema_step": 28.0 * tr.flat.padded,
                 "b200_ssl_loss_fwd": 4.0 * NCLS * Sx * (B + U_) + (B // 2) * Sx,
                 "b200_ssl_loss_bwd": 4.0 * NCLS * Sx * (B + U_) + (B // 2) * Sx + 4.0 * NCLS * Sx * B,
                 "b200_noise_add": 8.0 * U_ * Sx}
        table = {}
        for (name, tag), ms_step, calls in rows:
            if tag in layers and name.startswith(("b200_conv", "b200_bn", "b200_linear")):
                by, fl = conv_cost(layers[tag], name)
            elif name in extra:
                by, fl = extra[name] / calls, 0.0
            else:
                continue
            e = table.setdefault(name, [0.0, 0.0, 0.0, 0.0])
            e[0] += ms_step; e[1] += by * calls; e[2] += fl * calls; e[3] += calls
        ridge = pk["tf_sustained"] * 1e12 / (pk["hbm"] * 1e9)
        kernels = []
        for name, (ms_k, by_k, fl_k, calls_k) in sorted(table.items(), key=lambda kv: -kv[1][0]):
            gbs, tfs = by_k / (ms_k * 1e-3) / 1e9, fl_k / (ms_k * 1e-3) / 1e12
            bound = "tensor" if (fl_k and fl_k / by_k > ridge) else "hbm"
            kernels.append({"entry": name, "launches_per_step": round(calls_k, 1), "ms_per_step": round(ms_k, 4),
                            "GB/s": round(gbs, 1), "TFLOP/s": round(tfs, 1), "bound": bound,
                            "frac": round(tfs / pk["tf_sustained"] if bound == "tensor" else gbs / pk["hbm"], 3)})
            if bound == "tensor":      # the path computes in TF32, whose tensor-pipe rate is half the dense bf16 peak used above
                kernels[-1]["frac_of_tf32_rate"] = round(2 * tfs / pk["tf_sustained"], 3)
        shares = {k: round(v / total, 4) for k, v in sorted(by_name.items(), key=lambda kv: -kv[1])[:8]}
        value = B * world * args.steps / (ms * 1e-3)
        e2e_v = B * world * args.steps / (ms_e2e * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "iter_num": "1000+ (consistency term live)", "cuda_graph": tr.use_graph,
                       "numerics": "fp32 storage, TF32 tensor-core products, fp32 accumulation (the class cuDNN runs for the reference)",
                       "l2": "per-step working set (~6 GB of activations) >> 126 MB L2; no explicit flush needed"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_v, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": B * H * W * 4 + B * H * W + 32, "d2h_bytes_per_step": 16},
            "roofline": roof, "kernels": kernels, "step_time_shares": shares, "profiled_eager_ms_per_step": total, "loss": loss_dev,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            ts = time_cpu(B, 2, 1, cores)
            sec = sum(ts) / len(ts)
            out["cpu_baseline"] = {"value": B / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "2 timed + 1 warm-up full bs24 256x256 Mean-Teacher steps (oracle port, torch CPU fp32)"}
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
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
