"""Command line of code/train_interpolation_consistency_training_3D.py (default --model unet_3D, batch 4 = 2 labeled + the
two unlabeled patches that get mixed, :150-176): the ICT loop of the 2-D script over `net_factory_3d`."""
import sys

from ._common import base_parser, process_group, run_loop, seed_everything, setup_logging, snapshot_dir, synthetic_batches


def main(argv=None, loader=None):
    p = base_parser("BraTS2019_Interpolation_Consistency_Training", "unet_3D", 4, (96, 96, 96), 2, 14, "../data/BraTS2019")
    p.add_argument('--ict_alpha', type=int, default=0.2, help='ict_alpha')           # reference declares type=int, default 0.2
    args = p.parse_args(argv)
    args.num_classes = 2
    seed_everything(args)
    from ..networks.net_factory_3d import net_factory_3d
    from ..trainers import ICTTrainer
    pg, rank = process_group()

    def create_model(ema=False):
        model = net_factory_3d(net_type=args.model, in_chns=1, class_num=args.num_classes)
        if model is None:
            raise SystemExit(f"--model {args.model}: not built (available: unet_3D, vnet)")
        if ema:
            for param in model.parameters():
                param.detach_()
        return model

    model, ema_model = create_model(), create_model(ema=True)
    if pg is not None:
        import torch.distributed as dist
        for m in (model, ema_model):
            dist.broadcast(m.materialize().data, 0)
    trainer = ICTTrainer(model, ema_model, batch_size=args.batch_size, labeled_bs=args.labeled_bs, ict_alpha=args.ict_alpha,
                         patch_size=tuple(args.patch_size), num_classes=args.num_classes, base_lr=args.base_lr,
                         max_iterations=args.max_iterations, ema_decay=args.ema_decay, consistency=args.consistency,
                         consistency_rampup=args.consistency_rampup, process_group=pg, use_cuda_graph=not args.no_graph)
    if loader is None:
        if not args.synthetic:
            raise SystemExit("no h5 dataset reader in this package: pass batches to main() or use --synthetic 1")
        loader = synthetic_batches(args.batch_size, args.patch_size, args.num_classes, args.seed + rank)
    path = snapshot_dir(args)
    setup_logging(path)
    fmt = lambda it, l: 'iteration %d : loss : %f, loss_ce: %f, loss_dice: %f' % (it, l[3], l[0], l[1])
    return run_loop(args, trainer, loader, path, {"": model}, fmt, rank)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
