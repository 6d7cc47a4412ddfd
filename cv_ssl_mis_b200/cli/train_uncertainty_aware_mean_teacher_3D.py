"""Command line of code/train_uncertainty_aware_mean_teacher_3D.py (and code/train_mean_teacher_3D.py with
--uncertainty_T 0)."""
import sys

from ._common import base_parser, process_group, run_loop, seed_everything, setup_logging, snapshot_dir, synthetic_batches


def main(argv=None, loader=None, defaults=None):
    # the reference's defaults, including its experiment name (train_uncertainty_aware_mean_teacher_3D.py:33-36)
    p = base_parser("BraTs2019_Mean_Teacher", "unet_3D", 4, (96, 96, 96), 2, 25, "../data/BraTS2019")
    p.add_argument('--uncertainty_T', type=int, default=8, help='stochastic teacher passes (reference: T = 8)')
    if defaults:                                                  # same loop under another reference script name
        p.set_defaults(**defaults)
    args = p.parse_args(argv)
    args.num_classes = 2                                          # :100
    seed_everything(args)
    from ..networks.net_factory_3d import net_factory_3d
    from ..trainers import MeanTeacherTrainer
    pg, rank = process_group()

    def create_model(ema=False):                                  # :103-110
        net = net_factory_3d(net_type=args.model, in_chns=1, class_num=args.num_classes)
        if net is None:
            raise SystemExit(f"--model {args.model}: not built (available: unet_3D, vnet, unetr)")
        if ema:
            for param in net.parameters():
                param.detach_()
        return net

    model, ema_model = create_model(), create_model(ema=True)
    if pg is not None:
        import torch.distributed as dist
        for m in (model, ema_model):
            dist.broadcast(m.materialize().data, 0)
    trainer = MeanTeacherTrainer(model, ema_model, batch_size=args.batch_size, labeled_bs=args.labeled_bs,
                                 patch_size=tuple(args.patch_size), num_classes=args.num_classes, base_lr=args.base_lr,
                                 max_iterations=args.max_iterations, ema_decay=args.ema_decay, consistency=args.consistency,
                                 consistency_rampup=args.consistency_rampup, uncertainty_T=args.uncertainty_T,
                                 consistency_gate_iters=0, process_group=pg, use_cuda_graph=not args.no_graph)
    if loader is None:
        if not args.synthetic:
            raise SystemExit("no h5 dataset reader in this package: pass batches to main() or use --synthetic 1")
        loader = synthetic_batches(args.batch_size, args.patch_size, args.num_classes, args.seed + rank)
    path = snapshot_dir(args)
    setup_logging(path)
    fmt = lambda it, l: 'iteration %d : loss : %f, loss_ce: %f, loss_dice: %f' % (it, l[3], l[0], l[1])      # :202-204
    return run_loop(args, trainer, loader, path, {"": model}, fmt, rank)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
