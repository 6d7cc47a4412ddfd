"""Command line of code/train_fully_supervised_3D_ViT.py (UNETR by default; also code/train_fully_supervised_3D.py
with --model vnet)."""
import sys

from ._common import base_parser, process_group, run_loop, seed_everything, setup_logging, snapshot_dir, synthetic_batches


def main(argv=None, loader=None, defaults=None):
    p = base_parser("BraTs2019_Fully_Supervised", "unetr", 2, (96, 96, 96), 2, 25, "../data/BraTS2019", semi=False)
    if defaults:
        p.set_defaults(**defaults)
    args = p.parse_args(argv)
    args.num_classes = 2
    seed_everything(args)
    from ..networks.net_factory_3d import net_factory_3d
    from ..trainers import MeanTeacherTrainer
    pg, rank = process_group()
    model = net_factory_3d(net_type=args.model, in_chns=1, class_num=args.num_classes)            # :98
    if model is None:
        raise SystemExit(f"--model {args.model}: not built (available: unet_3D, vnet, unetr)")
    if pg is not None:
        import torch.distributed as dist
        dist.broadcast(model.materialize().data, 0)
    trainer = MeanTeacherTrainer(model, None, batch_size=args.batch_size, labeled_bs=args.batch_size,
                                 patch_size=tuple(args.patch_size), num_classes=args.num_classes, base_lr=args.base_lr,
                                 max_iterations=args.max_iterations, process_group=pg, use_cuda_graph=not args.no_graph)
    if loader is None:
        if not args.synthetic:
            raise SystemExit("no h5 dataset reader in this package: pass batches to main() or use --synthetic 1")
        loader = synthetic_batches(args.batch_size, args.patch_size, args.num_classes, args.seed + rank)
    path = snapshot_dir(args)
    setup_logging(path)
    fmt = lambda it, l: 'iteration %d : loss : %f, loss_ce: %f, loss_dice: %f' % (it, l[3], l[0], l[1])      # :133-135
    return run_loop(args, trainer, loader, path, {"": model}, fmt, rank)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
