"""Shared pieces of the train_*.py command lines.

Each script in this package keeps the flag set of the reference script of the same name (code/train_*.py: --root_path
--exp --model --max_iterations --batch_size --deterministic --base_lr --patch_size --seed --num_classes --labeled_bs
--labeled_num --ema_decay --consistency_type --consistency --consistency_rampup, plus the Swin flags --cfg --opts ...)
so existing launch lines keep working, builds the models through net_factory / net_factory_3d / ViT_seg like the
reference, and replaces the loop body by the fused trainer step (one CUDA graph per iteration).

Data loading is outside the hot path: `main(argv, loader=...)` accepts any iterable of
`{'image': [B,1,*patch] float32, 'label': [B,*patch] uint8|int64}` batches (the reference's `sampled_batch` dicts, e.g.
its own DataLoader over BaseDataSets + TwoStreamBatchSampler); without one, `--synthetic 1` (the default when the
reference's h5 datasets cannot be read here) draws ACDC- / BraTS-shaped random batches."""
from __future__ import annotations

import argparse
import logging
import os
import sys
import time

import torch


def base_parser(exp, model, batch_size, patch_size, labeled_bs, labeled_num, root_path, num_classes=None, semi=True):
    p = argparse.ArgumentParser()
    p.add_argument('--root_path', type=str, default=root_path, help='Name of Experiment')
    p.add_argument('--exp', type=str, default=exp, help='experiment_name')
    p.add_argument('--model', type=str, default=model, help='model_name')
    p.add_argument('--max_iterations', type=int, default=30000, help='maximum epoch number to train')
    p.add_argument('--batch_size', type=int, default=batch_size, help='batch_size per gpu')
    p.add_argument('--deterministic', type=int, default=1, help='whether use deterministic training')
    p.add_argument('--base_lr', type=float, default=0.01, help='segmentation network learning rate')
    p.add_argument('--patch_size', type=int, nargs='+', default=list(patch_size), help='patch size of network input')
    p.add_argument('--seed', type=int, default=1337, help='random seed')
    if num_classes is not None:
        p.add_argument('--num_classes', type=int, default=num_classes, help='output channel of network')
    p.add_argument('--labeled_num', type=int, default=labeled_num, help='labeled data')
    if semi:
        p.add_argument('--labeled_bs', type=int, default=labeled_bs, help='labeled_batch_size per gpu')
        p.add_argument('--ema_decay', type=float, default=0.99, help='ema_decay')
        p.add_argument('--consistency_type', type=str, default="mse", help='consistency_type')
        p.add_argument('--consistency', type=float, default=0.1, help='consistency')
        p.add_argument('--consistency_rampup', type=float, default=200.0, help='consistency_rampup')
    # ours
    p.add_argument('--synthetic', type=int, default=1, help='draw synthetic batches of the dataset\'s shape (no h5 reader here)')
    p.add_argument('--log_every', type=int, default=50, help='read the losses back every k iterations (0: never)')
    p.add_argument('--save_every', type=int, default=3000, help='checkpoint interval (reference: 3000)')
    p.add_argument('--resume_trainer', type=str, default=None, help='trainer_iter_<n>.pth written next to the checkpoints: continue that run')
    p.add_argument('--val_every', type=int, default=200, help='validation interval when main() gets a val_loader (reference: 200)')
    p.add_argument('--tensorboard', type=int, default=1, help='write the reference\'s tensorboard scalars under <snapshot>/log')
    p.add_argument('--no_graph', action='store_true', help='launch eagerly instead of replaying one CUDA graph per step')
    return p


def add_swin_flags(p):
    """Flags of the Swin-UNet scripts (code/train_cross_teaching_between_cnn_transformer_2D.py:66-92); accepted for
    launch-line compatibility; --cfg / --opts reach the model through build_swin_config."""
    p.add_argument('--cfg', type=str, default="../code/configs/swin_tiny_patch4_window7_224_lite.yaml", help='path to config file')
    p.add_argument("--opts", help="Modify config options by adding 'KEY VALUE' pairs. ", default=None, nargs='+')
    p.add_argument('--zip', action='store_true', help='use zipped dataset instead of folder dataset')
    p.add_argument('--cache-mode', type=str, default='part', choices=['no', 'full', 'part'])
    p.add_argument('--resume', help='resume from checkpoint')
    p.add_argument('--accumulation-steps', type=int, help="gradient accumulation steps")
    p.add_argument('--use-checkpoint', action='store_true')
    p.add_argument('--amp-opt-level', type=str, default='O1', choices=['O0', 'O1', 'O2'])
    p.add_argument('--tag', help='tag of experiment')
    p.add_argument('--eval', action='store_true', help='Perform evaluation only')
    p.add_argument('--throughput', action='store_true', help='Test throughput only')
    return p


def seed_everything(args):
    """code/train_mean_teacher_2D.py:316-326"""
    import random
    import numpy as np
    torch.backends.cudnn.benchmark = not args.deterministic
    torch.backends.cudnn.deterministic = bool(args.deterministic)
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(args.seed)


def process_group():
    """torchrun launch -> NCCL process group and device; otherwise single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return None, 0
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.group.WORLD, dist.get_rank()


def synthetic_batches(batch_size, patch, num_classes, seed, pinned=True):
    """Endless {'image', 'label'} batches: 2D = ACDC-shaped ([0,1] slices, uint8 labels), 3D = BraTS-shaped (z-scored
    volumes, int64 labels); blocky labels keep the Dice terms non-degenerate."""
    g = torch.Generator().manual_seed(seed)
    dims = len(patch)
    blk = 8
    while True:
        if dims == 2:
            x = torch.rand(batch_size, 1, *patch, generator=g)
        else:
            x = torch.randn(batch_size, 1, *patch, generator=g)
        low = torch.randint(0, num_classes, (batch_size, *[max(1, s // blk) for s in patch]), generator=g)
        y = low
        for d in range(dims):
            y = y.repeat_interleave(blk, d + 1)
        y = y[(slice(None),) + tuple(slice(0, s) for s in patch)]
        y = y.to(torch.uint8) if dims == 2 else y.to(torch.int64)
        if pinned and torch.cuda.is_available():
            x, y = x.pin_memory(), y.pin_memory()
        yield {"image": x, "label": y.contiguous()}


class _Node(dict):
    """Attribute-style nested config node (the slice of yacs.CfgNode the Swin-UNet reads)."""
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def build_swin_config(args):
    """The reference's `get_config(args)` (code/config.py:28-74,176-220) for the keys the Swin-UNet reads: code defaults,
    then the yaml at --cfg (if the file exists; the lite yaml's values are the built-in fallback), then --opts KEY VALUE
    pairs.  Returns a yacs-like object for `net_factory(config=...)` and `SwinUnet.load_from(config)`."""
    import ast
    cfg = _Node(DATA=_Node(IMG_SIZE=224),
                MODEL=_Node(TYPE="swin", NAME="swin_tiny_patch4_window7_224", DROP_RATE=0.0, DROP_PATH_RATE=0.1,
                            PRETRAIN_CKPT="./pretrained_ckpt/swin_tiny_patch4_window7_224.pth",
                            SWIN=_Node(PATCH_SIZE=4, IN_CHANS=3, EMBED_DIM=96, DEPTHS=[2, 2, 6, 2], DECODER_DEPTHS=[2, 2, 6, 2],
                                       NUM_HEADS=[3, 6, 12, 24], WINDOW_SIZE=7, MLP_RATIO=4.0, QKV_BIAS=True, QK_SCALE=None,
                                       APE=False, PATCH_NORM=True, FINAL_UPSAMPLE="expand_first")))
    lite = {"MODEL": {"DROP_PATH_RATE": 0.2, "PRETRAIN_CKPT": "../code/pretrained_ckpt/swin_tiny_patch4_window7_224.pth",
                      "SWIN": {"DEPTHS": [2, 2, 2, 2], "DECODER_DEPTHS": [2, 2, 2, 1]}}}

    def merge(node, upd):
        for k, v in upd.items():
            if isinstance(v, dict):
                merge(node.setdefault(k, _Node()), v)
            else:
                node[k] = v

    path = getattr(args, "cfg", None)
    if path and os.path.exists(path):
        import yaml
        with open(path) as f:
            merge(cfg, yaml.safe_load(f) or {})
    else:
        if path and os.path.basename(path) != "swin_tiny_patch4_window7_224_lite.yaml":
            raise SystemExit(f"--cfg {path}: file not found")
        merge(cfg, lite)           # the reference's default yaml, restated
    opts = getattr(args, "opts", None) or []
    if len(opts) % 2:
        raise SystemExit("--opts takes KEY VALUE pairs")
    for key, val in zip(opts[0::2], opts[1::2]):
        node = cfg
        *parents, leaf = key.split(".")
        for k in parents:
            if k not in node:
                raise SystemExit(f"--opts {key}: unknown config key")
            node = node[k]
        if leaf not in node:
            raise SystemExit(f"--opts {key}: unknown config key")
        try:
            node[leaf] = ast.literal_eval(val)
        except (ValueError, SyntaxError):
            node[leaf] = val
    if getattr(args, "batch_size", None):
        cfg.DATA.BATCH_SIZE = args.batch_size
    return cfg


def snapshot_dir(args):
    path = "../model/{}_{}_labeled/{}".format(args.exp, args.labeled_num, args.model)      # reference: train_*.py __main__
    os.makedirs(path, exist_ok=True)
    return path


def setup_logging(path):
    logging.basicConfig(filename=os.path.join(path, "log.txt"), level=logging.INFO,
                        format='[%(asctime)s.%(msecs)03d] %(message)s', datefmt='%H:%M:%S', force=True)
    logging.getLogger().addHandler(logging.StreamHandler(sys.stdout))


def _summary_writer(snapshot_path):
    """tensorboard scalars like the reference's tensorboardX writer (code/train_mean_teacher_2D.py:193); None if unavailable."""
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(os.path.join(snapshot_path, "log"))
    except Exception:                                            # tensorboard not installed: the text log carries the numbers
        return None


def _validate(args, trainer, val_loader, snapshot_path, models, writer, best, val_fn):
    """In-loop validation of the reference scripts (code/train_mean_teacher_2D.py:263-294, two-model form:
    code/train_cross_teaching_between_cnn_transformer_2D.py:283-345): per-class Dice / HD95 averaged over the validation
    volumes, best-so-far checkpoints `iter_<n>_dice_<d>.pth` and `<model>_best_model.pth`."""
    import numpy as np
    it = trainer.iter_num
    named = [(k, m) for k, m in models.items() if not k.startswith("ema")]
    for idx, (prefix, model) in enumerate(named):
        tag = "" if len(named) == 1 else f"model{idx + 1}_"
        was_training = model.training
        model.eval()
        metric_list, n = 0.0, 0
        for batch in val_loader:
            metric_list = metric_list + np.array(val_fn(batch["image"], batch["label"], model), dtype=float)
            n += 1
        model.train(was_training)
        if n == 0:
            raise ValueError("the validation loader yielded no volumes")
        metric_list = metric_list / n
        performance, mean_hd95 = float(np.mean(metric_list, axis=0)[0]), float(np.mean(metric_list, axis=0)[1])
        if writer is not None:
            for c in range(metric_list.shape[0]):
                writer.add_scalar(f"info/{tag}val_{c + 1}_dice", metric_list[c, 0], it)
                writer.add_scalar(f"info/{tag}val_{c + 1}_hd95", metric_list[c, 1], it)
            writer.add_scalar(f"info/{tag}val_mean_dice", performance, it)
            writer.add_scalar(f"info/{tag}val_mean_hd95", mean_hd95, it)
        if performance > best.get(prefix, 0.0):
            best[prefix] = performance
            stem = prefix if len(named) > 1 else ""
            torch.save(model.state_dict(), os.path.join(snapshot_path, f"{stem}iter_{it}_dice_{round(performance, 4)}.pth"))
            torch.save(model.state_dict(), os.path.join(snapshot_path, f"{args.model}_best_{prefix.rstrip('_') or 'model'}.pth"))
        logging.info('iteration %d : %smean_dice : %f %smean_hd95 : %f' % (it, tag, performance, tag, mean_hd95))


def run_loop(args, trainer, loader, snapshot_path, models, fmt, rank=0, val_loader=None, val_fn=None, scalars=None):
    """The iteration loop of the reference scripts with the body replaced by `trainer.step`.
    models: {checkpoint prefix: module}; fmt(iter_num, losses) -> log line; scalars(iter_num, losses) -> {tag: value} for
    tensorboard; val_loader / val_fn(image, label, model) -> [(dice, hd95)] per class: validation every --val_every iterations."""
    if getattr(args, "resume_trainer", None):
        trainer.load_state_dict(torch.load(args.resume_trainer, weights_only=False))
        logging.info("resumed from %s at iteration %d" % (args.resume_trainer, trainer.iter_num))
    it0 = trainer.iter_num
    t0 = time.time()
    writer = _summary_writer(snapshot_path) if (rank == 0 and getattr(args, "tensorboard", 1)) else None
    best = {}
    # the reference wraps its loader in `for epoch_num in range(max_iterations // len(trainloader) + 1)`
    # (code/train_mean_teacher_2D.py:199-201,296-301): a finite loader is re-iterated until max_iterations is reached
    while trainer.iter_num < args.max_iterations:
        seen = 0
        for batch in loader:
            seen += 1
            log = args.log_every and (trainer.iter_num + 1) % args.log_every == 0
            out = trainer.step(batch["image"], batch["label"], read_loss=bool(log))
            if log and rank == 0:
                logging.info(fmt(trainer.iter_num, out))
                if writer is not None and scalars is not None:
                    for tag, v in scalars(trainer.iter_num, out).items():
                        writer.add_scalar(tag, v, trainer.iter_num)
            if (rank == 0 and val_loader is not None and getattr(args, "val_every", 0) and trainer.iter_num > 0
                    and trainer.iter_num % args.val_every == 0):
                _validate(args, trainer, val_loader, snapshot_path, models, writer, best, val_fn)
            if rank == 0 and args.save_every and trainer.iter_num % args.save_every == 0:
                for prefix, m in models.items():
                    torch.save(m.state_dict(), os.path.join(snapshot_path, f"{prefix}iter_{trainer.iter_num}.pth"))
                torch.save(trainer.state_dict(), os.path.join(snapshot_path, f"trainer_iter_{trainer.iter_num}.pth"))
            if trainer.iter_num >= args.max_iterations:
                break
        if seen == 0:
            raise ValueError("the data loader yielded no batches")
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        n = trainer.iter_num - it0
        logging.info("%d iterations in %.2f s (%.1f samples/s per GPU)" % (n, dt, n * args.batch_size / max(dt, 1e-9)))
        for prefix, m in models.items():
            torch.save(m.state_dict(), os.path.join(snapshot_path, f"{prefix}iter_{trainer.iter_num}.pth"))
        if writer is not None:
            writer.close()
    return "Training Finished!"
