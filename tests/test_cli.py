"""The train_*.py command lines keep the reference's flag sets (code/train_*.py argparse blocks) and drive the fused
trainers; exercised on CPU with the stand-in ops at toy sizes (two iterations each, synthetic batches)."""
import os

import pytest
import torch

from tests import fake_ops


@pytest.fixture()
def cpu_env(monkeypatch, tmp_path):
    fake_ops.install(monkeypatch)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)     # net_factory(...).cuda() like the reference
    code = tmp_path / "code"
    code.mkdir()
    monkeypatch.chdir(code)                                                       # snapshots go to ../model/<exp>_<n>_labeled/<model>
    return tmp_path


COMMON = ["--max_iterations", "2", "--log_every", "1", "--save_every", "0", "--no_graph", "--seed", "7"]


def test_mean_teacher_2d_cli(cpu_env):
    from cv_ssl_mis_b200.cli import train_mean_teacher_2D as cli
    out = cli.main(["--root_path", "../data/ACDC", "--exp", "ACDC/Mean_Teacher", "--model", "unet", "--batch_size", "4",
                    "--labeled_bs", "2", "--patch_size", "32", "32", "--num_classes", "4", "--labeled_num", "7",
                    "--ema_decay", "0.99", "--consistency_type", "mse", "--consistency", "0.1", "--consistency_rampup", "200.0",
                    "--base_lr", "0.01", "--deterministic", "1"] + COMMON)
    assert out == "Training Finished!"
    snap = cpu_env / "model" / "ACDC" / "Mean_Teacher_7_labeled" / "unet"
    sd = torch.load(snap / "iter_2.pth")
    assert "encoder.in_conv.conv_conv.0.weight" in sd and "decoder.out_conv.bias" in sd       # reference checkpoint schema
    assert (snap / "ema_iter_2.pth").exists()
    log = (snap / "log.txt").read_text()
    assert "iteration 2 : loss :" in log and "loss_ce:" in log and "loss_dice:" in log


def test_mean_teacher_2d_cli_uamt_and_supervised(cpu_env):
    from cv_ssl_mis_b200.cli import train_mean_teacher_2D as cli
    assert cli.main(["--batch_size", "4", "--labeled_bs", "2", "--patch_size", "32", "32", "--uncertainty_T", "8"] + COMMON) \
        == "Training Finished!"
    assert cli.main(["--batch_size", "2", "--labeled_bs", "2", "--patch_size", "32", "32", "--exp", "ACDC/Fully_Supervised"]
                    + COMMON) == "Training Finished!"
    with pytest.raises(SystemExit):
        cli.main(["--model", "enet"] + COMMON)


def test_cross_teaching_cli_flags_and_cps(cpu_env):
    from cv_ssl_mis_b200.cli import train_cross_teaching_between_cnn_transformer_2D as cli
    out = cli.main(["--model", "unet", "--model2", "unet", "--pseudo_loss", "ce", "--batch_size", "4", "--labeled_bs", "2",
                    "--patch_size", "32", "32", "--cfg", "../code/configs/swin_tiny_patch4_window7_224_lite.yaml",
                    "--opts", "MODEL.DROP_PATH_RATE", "0.2", "--cache-mode", "part", "--amp-opt-level", "O1"] + COMMON)
    assert out == "Training Finished!"
    snap = cpu_env / "model" / "ACDC" / "Cross_Teaching_Between_CNN_Transformer_7_labeled" / "unet"
    assert (snap / "model1_iter_2.pth").exists() and (snap / "model2_iter_2.pth").exists()
    assert "model1 loss" in (snap / "log.txt").read_text()


def test_3d_clis(cpu_env):
    from cv_ssl_mis_b200.cli import train_uncertainty_aware_mean_teacher_3D as uamt, train_fully_supervised_3D_ViT as fs
    assert uamt.main(["--model", "vnet", "--batch_size", "2", "--labeled_bs", "1", "--patch_size", "16", "16", "16",
                      "--uncertainty_T", "2"] + COMMON) == "Training Finished!"
    assert fs.main(["--model", "vnet", "--batch_size", "1", "--patch_size", "16", "16", "16"] + COMMON) == "Training Finished!"
    with pytest.raises(SystemExit):
        fs.main(["--model", "voxresnet"] + COMMON)


def test_ict_cli(cpu_env):
    from cv_ssl_mis_b200.cli import train_interpolation_consistency_training_2D as cli
    assert cli.main(["--batch_size", "8", "--labeled_bs", "4", "--patch_size", "32", "32", "--ict_alpha", "1"] + COMMON) \
        == "Training Finished!"


def test_finite_loader_is_reiterated_until_max_iterations(cpu_env):
    """The reference loops `for epoch_num in range(max_iterations // len(trainloader) + 1)` over its DataLoader
    (code/train_mean_teacher_2D.py:199-201): a 2-batch loader must still run max_iterations = 5 steps."""
    from cv_ssl_mis_b200.cli import train_mean_teacher_2D as cli
    from cv_ssl_mis_b200.cli._common import synthetic_batches
    gen = synthetic_batches(4, (32, 32), 4, 3, pinned=False)
    batches = [next(gen) for _ in range(2)]
    out = cli.main(["--batch_size", "4", "--labeled_bs", "2", "--patch_size", "32", "32", "--max_iterations", "5", "--log_every", "1",
                    "--save_every", "0", "--no_graph", "--seed", "7"], loader=batches)
    assert out == "Training Finished!"
    snap = cpu_env / "model" / "ACDC" / "Mean_Teacher_7_labeled" / "unet"
    assert (snap / "iter_5.pth").exists() and "iteration 5 :" in (snap / "log.txt").read_text()


def test_swin_config_from_cfg_and_opts(tmp_path):
    """--cfg / --opts reach the Swin-UNet (code/config.py:get_config): yaml over the code defaults, then KEY VALUE pairs."""
    import argparse
    from cv_ssl_mis_b200.cli._common import build_swin_config
    from cv_ssl_mis_b200.networks.swin_unet import _read_config
    ns = argparse.Namespace(cfg="../code/configs/swin_tiny_patch4_window7_224_lite.yaml", opts=None, batch_size=16)
    cfg = build_swin_config(ns)                       # file absent: the lite yaml restated
    assert cfg.MODEL.DROP_PATH_RATE == 0.2 and cfg.MODEL.SWIN.DEPTHS == [2, 2, 2, 2] and cfg.DATA.IMG_SIZE == 224
    assert _read_config(cfg)["drop_path_rate"] == 0.2 and _read_config(cfg)["depths"] == (2, 2, 2, 2)
    y = tmp_path / "my.yaml"
    y.write_text("MODEL:\n  DROP_PATH_RATE: 0.3\n  SWIN:\n    DEPTHS: [2, 2, 6, 2]\n")
    ns = argparse.Namespace(cfg=str(y), opts=["MODEL.DROP_PATH_RATE", "0.05", "MODEL.PRETRAIN_CKPT", "none.pth"], batch_size=8)
    cfg = build_swin_config(ns)
    assert cfg.MODEL.DROP_PATH_RATE == 0.05 and cfg.MODEL.SWIN.DEPTHS == [2, 2, 6, 2] and cfg.MODEL.PRETRAIN_CKPT == "none.pth"
    with pytest.raises(SystemExit):
        build_swin_config(argparse.Namespace(cfg=str(y), opts=["MODEL.NOPE", "1"], batch_size=8))
    with pytest.raises(SystemExit):
        build_swin_config(argparse.Namespace(cfg=str(tmp_path / "missing.yaml"), opts=None, batch_size=8))


def test_swin_config_reaches_the_model():
    """build_swin_config's yacs-like node must be read as a config, not as a keyword dict (found by tools/gpu_smoke_cli_vit.py)."""
    import argparse
    from cv_ssl_mis_b200.cli._common import build_swin_config
    from cv_ssl_mis_b200.networks.swin_unet import _read_config
    cfg = build_swin_config(argparse.Namespace(cfg=None, opts=["MODEL.DROP_PATH_RATE", "0.3"]))
    kw = _read_config(cfg)
    assert kw["depths"] == (2, 2, 2, 2) and abs(kw["drop_path_rate"] - 0.3) < 1e-12 and kw["img_size"] == 224


def test_reference_script_names(cpu_env):
    """The reference's other script names on the path run the same loops with their own defaults
    (code/train_uncertainty_aware_mean_teacher_2D.py, train_fully_supervised_2D.py, train_mean_teacher_3D.py,
    train_cross_pseudo_supervision_2D.py; train_mean_teacher_ViT.py is only parsed here -- a 224^2 Swin-UNet is a GPU job)."""
    from cv_ssl_mis_b200.cli import (train_uncertainty_aware_mean_teacher_2D as uamt2d, train_fully_supervised_2D as fs2d,
                                     train_mean_teacher_3D as mt3d, train_cross_pseudo_supervision_2D as cps,
                                     train_mean_teacher_ViT as mtvit)
    small = ["--batch_size", "4", "--labeled_bs", "2", "--patch_size", "32", "32"]
    assert uamt2d.main(small + COMMON) == "Training Finished!"
    assert (cpu_env / "model" / "ACDC" / "Uncertainty_Aware_Mean_Teacher_136_labeled" / "unet" / "ema_iter_2.pth").exists()
    assert fs2d.main(["--batch_size", "2", "--patch_size", "32", "32"] + COMMON) == "Training Finished!"
    snap = cpu_env / "model" / "ACDC" / "Fully_Supervised_50_labeled" / "unet"
    assert (snap / "iter_2.pth").exists() and not (snap / "ema_iter_2.pth").exists()
    assert cps.main(small + COMMON) == "Training Finished!"
    assert (cpu_env / "model" / "ACDC" / "Cross_Pseudo_Supervision_1_labeled" / "unet" / "model2_iter_2.pth").exists()
    assert mt3d.main(["--batch_size", "2", "--labeled_bs", "1", "--patch_size", "16", "16", "16"] + COMMON) == "Training Finished!"
    assert (cpu_env / "model" / "BraTs2019_Mean_Teacher_25_labeled" / "unet_3D" / "iter_2.pth").exists()       # the reference's default model
    assert mtvit.DEFAULTS["vit"] == 1 and mtvit.DEFAULTS["exp"] == "ACDC/Mean_Teacher_ViT"
    from cv_ssl_mis_b200.cli import (train_fully_supervised_3D as fs3d, train_fully_supervised_2D_ViT as fsvit,
                                     train_uncertainty_aware_mean_teacher_ViT_2D as uamtvit,
                                     train_interpolation_consistency_training_2D_ViT as ictvit,
                                     train_cross_pseudo_supervision_2D_ViT as cpsvit)
    assert fs3d.main(["--batch_size", "1", "--patch_size", "16", "16", "16"] + COMMON) == "Training Finished!"
    assert (cpu_env / "model" / "BraTS2019" / "Fully_Supervised_25_labeled" / "unet_3D" / "iter_2.pth").exists()
    from cv_ssl_mis_b200.cli import train_cross_pseudo_supervision_3D as cps3d
    assert cps3d.main(["--model", "vnet", "--batch_size", "2", "--labeled_bs", "1", "--patch_size", "16", "16", "16"] + COMMON) \
        == "Training Finished!"
    assert (cpu_env / "model" / "BraTs2019_Cross_Pseudo_Supervision_25_labeled" / "vnet" / "model2_iter_2.pth").exists()
    from cv_ssl_mis_b200.cli import train_interpolation_consistency_training_3D as ict3d
    assert ict3d.main(["--model", "vnet", "--batch_size", "4", "--labeled_bs", "2", "--patch_size", "16", "16", "16"] + COMMON) \
        == "Training Finished!"
    # the Swin-UNet variants are only parsed here (a 224^2 Swin-UNet step is a GPU job)
    assert fsvit.DEFAULTS["supervised"] == 1 and uamtvit.DEFAULTS["uncertainty_T"] == 8 and ictvit.DEFAULTS["vit"] == 1
    assert cpsvit.DEFAULTS["vit1"] == 1 and cpsvit.DEFAULTS["pseudo_loss"] == "ce"


def test_in_loop_validation_and_best_checkpoint(cpu_env):
    """code/train_mean_teacher_2D.py:263-294: every --val_every iterations the validation volumes go through
    test_single_volume in eval mode; mean Dice / HD95 are logged and the best model is saved under the reference's names."""
    from cv_ssl_mis_b200.cli import train_mean_teacher_2D as cli
    g = torch.Generator().manual_seed(9)
    val = [{"image": torch.rand(1, 3, 40, 36, generator=g), "label": torch.randint(0, 4, (1, 3, 40, 36), generator=g)} for _ in range(2)]
    out = cli.main(["--batch_size", "4", "--labeled_bs", "2", "--patch_size", "32", "32", "--val_every", "2"] + COMMON, val_loader=val)
    assert out == "Training Finished!"
    snap = cpu_env / "model" / "ACDC" / "Mean_Teacher_7_labeled" / "unet"
    assert (snap / "unet_best_model.pth").exists() and list(snap.glob("iter_2_dice_*.pth"))
    assert "mean_dice" in (snap / "log.txt").read_text()
    assert list((snap / "log").glob("events.out.tfevents.*"))            # tensorboard scalars (info/lr, info/total_loss, ...)


def test_in_loop_validation_two_models(cpu_env):
    """code/train_cross_teaching_between_cnn_transformer_2D.py:283-345: both networks are validated, each with its own
    best-model files (`model1_iter_<n>_dice_<d>.pth`, `<model>_best_model1.pth`, ...)."""
    from cv_ssl_mis_b200.cli import train_cross_pseudo_supervision_2D as cli
    g = torch.Generator().manual_seed(10)
    val = [{"image": torch.rand(1, 2, 40, 36, generator=g), "label": torch.randint(0, 4, (1, 2, 40, 36), generator=g)}]
    assert cli.main(["--batch_size", "4", "--labeled_bs", "2", "--patch_size", "32", "32", "--val_every", "2"] + COMMON,
                    val_loader=val) == "Training Finished!"
    snap = cpu_env / "model" / "ACDC" / "Cross_Pseudo_Supervision_1_labeled" / "unet"
    for k in (1, 2):
        assert (snap / f"unet_best_model{k}.pth").exists() and list(snap.glob(f"model{k}_iter_2_dice_*.pth"))
    log = (snap / "log.txt").read_text()
    assert "model1_mean_dice" in log and "model2_mean_dice" in log


def test_resume_continues_the_same_run(cpu_env):
    """A run interrupted after 2 of 4 iterations and resumed from trainer_iter_2.pth ends with the weights of the
    uninterrupted run (parameters, momentum, teacher, BatchNorm statistics, RNG epochs, iteration counter, learning rate)."""
    from cv_ssl_mis_b200.cli import train_mean_teacher_2D as cli
    from cv_ssl_mis_b200.networks import unet as unet_mod
    base = ["--batch_size", "4", "--labeled_bs", "2", "--patch_size", "32", "32", "--log_every", "1", "--no_graph", "--seed", "7"]

    def fresh_process():                 # the default dropout seeds count the networks built in this process
        unet_mod.UNet._instances = 0

    fresh_process()
    assert cli.main(base + ["--max_iterations", "4", "--save_every", "0", "--exp", "R/straight"]) == "Training Finished!"
    fresh_process()
    from cv_ssl_mis_b200.cli._common import synthetic_batches
    two = synthetic_batches(4, [32, 32], 4, 7, pinned=False)
    with pytest.raises(ValueError):      # the "interruption": the data stream ends after two batches of the same 4-iteration run
        cli.main(base + ["--max_iterations", "4", "--save_every", "2", "--exp", "R/first"], loader=(next(two) for _ in range(2)))
    first = cpu_env / "model" / "R" / "first_7_labeled" / "unet"
    # same data stream as the straight run from iteration 2 on: skip the two batches the first leg consumed
    it = synthetic_batches(4, [32, 32], 4, 7, pinned=False)
    next(it), next(it)
    fresh_process()
    assert cli.main(base + ["--max_iterations", "4", "--save_every", "0", "--exp", "R/second",
                            "--resume_trainer", str(first / "trainer_iter_2.pth")], loader=it) == "Training Finished!"
    a = torch.load(cpu_env / "model" / "R" / "straight_7_labeled" / "unet" / "iter_4.pth")
    b = torch.load(cpu_env / "model" / "R" / "second_7_labeled" / "unet" / "iter_4.pth")
    for k in a:
        torch.testing.assert_close(a[k], b[k], rtol=0, atol=0, msg=lambda m, k=k: f"{k}: {m}")
