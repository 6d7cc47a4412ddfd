"""Command line of code/train_mean_teacher_2D.py (also code/train_uncertainty_aware_mean_teacher_2D.py with
--uncertainty_T 8, and code/train_fully_supervised_2D.py with --labeled_bs equal to --batch_size)."""
import sys

from ._common import add_swin_flags, base_parser, build_swin_config, process_group, run_loop, seed_everything, setup_logging, snapshot_dir, synthetic_batches


def main(argv=None, loader=None, defaults=None, val_loader=None):
    p = base_parser("ACDC/Mean_Teacher", "unet", 24, (224, 224), 12, 7, "../data/ACDC", num_classes=4)
    p.add_argument('--uncertainty_T', type=int, default=0, help='8: the uncertainty-aware variant (MC-dropout mask)')
    p.add_argument('--supervised', type=int, default=0, help='1: no unlabeled half (train_fully_supervised_2D.py)')
    p.add_argument('--vit', type=int, default=0, help='1: build the Swin-UNet ViT_seg whatever --model says (the *_ViT.py scripts)')
    add_swin_flags(p)                                             # only used by the Swin-UNet
    if defaults:                                                  # same loop under another reference script name
        p.set_defaults(**defaults)
    args = p.parse_args(argv)
    if args.supervised:
        args.labeled_bs = args.batch_size
    seed_everything(args)
    from ..networks.net_factory import net_factory
    from ..trainers import MeanTeacherTrainer
    pg, rank = process_group()

    def create_model(ema=False):                                  # code/train_mean_teacher_2D.py:136-144
        name = "ViT_Seg" if (args.vit or args.model == "ViT_Seg") else args.model
        kw = dict(config=build_swin_config(args), img_size=args.patch_size) if name == "ViT_Seg" else {}
        model = net_factory(net_type=name, in_chns=1, class_num=args.num_classes, **kw)
        if model is None:
            raise SystemExit(f"--model {args.model}: not built (available: unet, ViT_Seg)")
        if ema:
            for param in model.parameters():
                param.detach_()
        return model

    model = create_model()
    supervised = args.labeled_bs >= args.batch_size
    ema_model = None if supervised else create_model(ema=True)
    if pg is not None:                                            # identical replicas
        import torch.distributed as dist
        for m in (model, ema_model):
            if m is not None:
                dist.broadcast(m.materialize().data, 0)
    trainer = MeanTeacherTrainer(model, ema_model, batch_size=args.batch_size, labeled_bs=args.labeled_bs,
                                 patch_size=tuple(args.patch_size), num_classes=args.num_classes, base_lr=args.base_lr,
                                 max_iterations=args.max_iterations, ema_decay=args.ema_decay, consistency=args.consistency,
                                 consistency_rampup=args.consistency_rampup, uncertainty_T=args.uncertainty_T,
                                 consistency_gate_iters=0 if args.uncertainty_T else 1000,      # only MT2D gates (:224-225)
                                 process_group=pg, use_cuda_graph=not args.no_graph)
    if loader is None:
        if not args.synthetic:
            raise SystemExit("no h5 dataset reader in this package: pass an iterable of {'image','label'} batches to main(), "
                             "or use --synthetic 1")
        loader = synthetic_batches(args.batch_size, args.patch_size, args.num_classes, args.seed + rank)
    path = snapshot_dir(args)
    setup_logging(path)
    fmt = lambda it, l: 'iteration %d : loss : %f, loss_ce: %f, loss_dice: %f' % (it, l[3], l[0], l[1])      # :252-254
    models = {"": model}
    if ema_model is not None:
        models["ema_"] = ema_model
    from ..val_2D import test_single_volume
    val_fn = lambda image, label, net: test_single_volume(image, label, net, classes=args.num_classes, patch_size=args.patch_size)
    scalars = lambda it, l: {'info/lr': trainer.lr, 'info/total_loss': l[3], 'info/loss_ce': l[0], 'info/loss_dice': l[1],      # :238-246
                             'info/consistency_loss': l[2], 'info/consistency_weight': trainer.consistency_weight(it)}
    return run_loop(args, trainer, loader, path, models, fmt, rank, val_loader=val_loader, val_fn=val_fn, scalars=scalars)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
