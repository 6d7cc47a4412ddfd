"""Command line of code/train_interpolation_consistency_training_2D.py (flags :38-76, incl. --ict_alpha)."""
import sys

from ._common import add_swin_flags, base_parser, build_swin_config, process_group, run_loop, seed_everything, setup_logging, snapshot_dir, synthetic_batches


def main(argv=None, loader=None, defaults=None, val_loader=None):
    p = base_parser("ACDC/Interpolation_Consistency_Training", "unet", 24, (256, 256), 12, 300, "../data/ACDC", num_classes=4)
    p.add_argument('--ict_alpha', type=int, default=0.2, help='ict_alpha')           # reference declares type=int, default 0.2
    p.add_argument('--vit', type=int, default=0, help='1: build the Swin-UNet ViT_seg (train_interpolation_consistency_training_2D_ViT.py)')
    add_swin_flags(p)
    if defaults:
        p.set_defaults(**defaults)
    args = p.parse_args(argv)
    seed_everything(args)
    from ..networks.net_factory import net_factory
    from ..trainers import ICTTrainer
    pg, rank = process_group()

    def create_model(ema=False):
        name = "ViT_Seg" if (args.vit or args.model == "ViT_Seg") else args.model
        kw = dict(config=build_swin_config(args), img_size=args.patch_size) if name == "ViT_Seg" else {}
        model = net_factory(net_type=name, in_chns=1, class_num=args.num_classes, **kw)
        if model is None:
            raise SystemExit(f"--model {args.model}: not built (available: unet, ViT_Seg)")
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
    from ..val_2D import test_single_volume
    val_fn = lambda image, label, net: test_single_volume(image, label, net, classes=args.num_classes, patch_size=args.patch_size)
    scalars = lambda it, l: {'info/lr': trainer.lr, 'info/total_loss': l[3], 'info/loss_ce': l[0], 'info/loss_dice': l[1],
                             'info/consistency_loss': l[2], 'info/consistency_weight': trainer.consistency_weight(it)}
    return run_loop(args, trainer, loader, path, {"": model}, fmt, rank, val_loader=val_loader, val_fn=val_fn, scalars=scalars)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
