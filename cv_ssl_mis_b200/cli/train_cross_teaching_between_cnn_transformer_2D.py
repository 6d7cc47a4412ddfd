"""Command line of code/train_cross_teaching_between_cnn_transformer_2D.py; --pseudo_loss ce with --model2 unet gives
code/train_cross_pseudo_supervision_2D.py."""
import sys

from ._common import (add_swin_flags, base_parser, build_swin_config, process_group, run_loop, seed_everything, setup_logging,
                      snapshot_dir, synthetic_batches)


def main(argv=None, loader=None, defaults=None, val_loader=None):
    p = base_parser("ACDC/Cross_Teaching_Between_CNN_Transformer", "unet", 16, (224, 224), 8, 7, "../data/ACDC", num_classes=4)
    add_swin_flags(p)
    p.add_argument('--model2', type=str, default="ViT_Seg", help='second network (reference: the Swin-UNet ViT_seg)')
    p.add_argument('--vit1', type=int, default=0, help='1: model 1 is a Swin-UNet as well (train_cross_pseudo_supervision_2D_ViT.py)')
    p.add_argument('--pseudo_loss', type=str, default="dice", choices=["dice", "ce"],
                   help='dice: cross teaching (:242-245); ce: cross pseudo supervision')
    if defaults:                                                  # same loop under another reference script name
        p.set_defaults(**defaults)
    args = p.parse_args(argv)
    seed_everything(args)
    from ..networks.net_factory import net_factory
    from ..trainers import CrossTeachingTrainer
    pg, rank = process_group()
    config = build_swin_config(args) if (args.model2 == "ViT_Seg" or args.vit1) else None         # config.py:get_config(args)
    if args.vit1:
        model1 = net_factory(net_type="ViT_Seg", in_chns=1, class_num=args.num_classes, config=config, img_size=args.patch_size)
    else:
        model1 = net_factory(net_type=args.model, in_chns=1, class_num=args.num_classes)          # :134-141
    model2 = net_factory(net_type=args.model2, in_chns=1, class_num=args.num_classes, config=config,
                         img_size=args.patch_size)                                                # :142-144 (ViT_seg)
    if model1 is None or model2 is None:
        raise SystemExit("--model / --model2: not built (available: unet, ViT_Seg)")
    if config is not None:
        # :145 `model2.load_from(config)`: ImageNet-pretrained Swin encoder, copied into the decoder as well.  The
        # reference dies without the checkpoint; here a missing file is reported and the Swin-UNet starts from its
        # random initialisation (the synthetic benchmarks have no checkpoint).
        import os
        if os.path.exists(config.MODEL.PRETRAIN_CKPT):
            model2.load_from(config)
            model2._flat, model2._plans = None, {}         # parameters were replaced: re-home them on first use
        else:
            print(f"pretrained checkpoint {config.MODEL.PRETRAIN_CKPT} not found: Swin-UNet trains from scratch")
    if pg is not None:
        import torch.distributed as dist
        for m in (model1, model2):
            dist.broadcast(m.materialize().data, 0)
    trainer = CrossTeachingTrainer(model1, model2, batch_size=args.batch_size, labeled_bs=args.labeled_bs,
                                   patch_size=tuple(args.patch_size), num_classes=args.num_classes, base_lr=args.base_lr,
                                   max_iterations=args.max_iterations, consistency=args.consistency,
                                   consistency_rampup=args.consistency_rampup, process_group=pg,
                                   use_cuda_graph=not args.no_graph, pseudo_loss=args.pseudo_loss)
    if loader is None:
        if not args.synthetic:
            raise SystemExit("no h5 dataset reader in this package: pass batches to main() or use --synthetic 1")
        loader = synthetic_batches(args.batch_size, args.patch_size, args.num_classes, args.seed + rank)
    path = snapshot_dir(args)
    setup_logging(path)
    fmt = lambda it, l: 'iteration %d : model1 loss : %f model2 loss : %f' % (it, l[3], l[7])      # :271-272
    from ..val_2D import test_single_volume
    val_fn = lambda image, label, net: test_single_volume(image, label, net, classes=args.num_classes, patch_size=args.patch_size)
    scalars = lambda it, l: {'lr': trainer.lr, 'consistency_weight/consistency_weight': trainer.consistency_weight(it),       # :263-268
                             'loss/model1_loss': l[3], 'loss/model2_loss': l[7]}
    return run_loop(args, trainer, loader, path, {"model1_": model1, "model2_": model2}, fmt, rank, val_loader=val_loader, val_fn=val_fn, scalars=scalars)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
