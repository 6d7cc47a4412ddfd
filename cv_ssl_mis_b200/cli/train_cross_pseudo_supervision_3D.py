"""Command line of code/train_cross_pseudo_supervision_3D.py: two 3-D networks of the same architecture (default unet_3D), each
trained with 0.5 (CE + Dice) on the labeled patches plus w * CE against the argmax pseudo labels of the OTHER network on the
unlabeled ones (:152-176) -- the Cross-Pseudo-Supervision loop of the 2-D script over `net_factory_3d`."""
import sys

from ._common import base_parser, process_group, run_loop, seed_everything, setup_logging, snapshot_dir, synthetic_batches


def main(argv=None, loader=None):
    p = base_parser("BraTs2019_Cross_Pseudo_Supervision", "unet_3D", 4, (96, 96, 96), 2, 25, "../data/BraTS2019")
    args = p.parse_args(argv)
    args.num_classes = 2
    seed_everything(args)
    import torch
    from ..networks.net_factory_3d import net_factory_3d
    from ..trainers import CrossTeachingTrainer
    pg, rank = process_group()
    models = [net_factory_3d(net_type=args.model, in_chns=1, class_num=args.num_classes) for _ in range(2)]     # :107-112
    if models[0] is None:
        raise SystemExit(f"--model {args.model}: not built (available: unet_3D, vnet)")
    if pg is not None:
        import torch.distributed as dist
        for m in models:
            dist.broadcast(m.materialize().data, 0)
    trainer = CrossTeachingTrainer(models[0], models[1], batch_size=args.batch_size, labeled_bs=args.labeled_bs,
                                   patch_size=tuple(args.patch_size), num_classes=args.num_classes, base_lr=args.base_lr,
                                   max_iterations=args.max_iterations, consistency=args.consistency,
                                   consistency_rampup=args.consistency_rampup, label_dtype=torch.int64, process_group=pg,
                                   use_cuda_graph=not args.no_graph, pseudo_loss="ce")
    if loader is None:
        if not args.synthetic:
            raise SystemExit("no h5 dataset reader in this package: pass batches to main() or use --synthetic 1")
        loader = synthetic_batches(args.batch_size, args.patch_size, args.num_classes, args.seed + rank)
    path = snapshot_dir(args)
    setup_logging(path)
    fmt = lambda it, l: 'iteration %d : model1 loss : %f model2 loss : %f' % (it, l[3], l[7])                   # :191-192
    return run_loop(args, trainer, loader, path, {"model1_": models[0], "model2_": models[1]}, fmt, rank)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
