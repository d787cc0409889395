"""The Lightning / torchmetrics / FusedAdam stand-ins driving THIS repository's models on the GPU: a LightningModule
shaped like ngp_pl/train.py's NeRFSystem (own code; the reference script itself cannot travel to the GPU box) trains on
the synthetic NSVF dataset through render() + autograd under fp16 autocast, validates one image and writes a
Lightning-style checkpoint that utils.load_ckpt / slim_ckpt read back."""
import argparse
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "google-nerf_b200", "shims")


def test_lightning_stand_in_trains_ngp(built_lib, tmp_path):
    for p in (ROOT, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from apex.optimizers import FusedAdam
    from pytorch_lightning import LightningModule, Trainer
    from pytorch_lightning.callbacks import ModelCheckpoint, TQDMProgressBar
    from pytorch_lightning.loggers import TensorBoardLogger
    from torchmetrics import PeakSignalNoiseRatio
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.losses import NeRFLoss
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.models.rendering import MAX_SAMPLES, render
    from google_nerf_b200.utils import load_ckpt, slim_ckpt

    res, n_train = 48, 6
    root = os.path.join(tmp_path, "Synthetic_NSVF", "Analytic")
    train_poses, test_poses = syn.write_nsvf_dataset(root, n_train=n_train, n_test=1, res=res)
    K = syn.intrinsics(res, res); dirs = syn.directions(res, res, K)
    gt = torch.stack([syn.shade(*syn.get_rays(dirs, p), 0.5) for p in train_poses])              # (n_train, res*res, 3)
    gt_test = syn.shade(*syn.get_rays(dirs, test_poses[0]), 0.5)

    class System(LightningModule):
        def __init__(self, hp):
            super().__init__()
            self.save_hyperparameters(hp)
            self.loss, self.train_psnr, self.val_psnr = NeRFLoss(), PeakSignalNoiseRatio(data_range=1), PeakSignalNoiseRatio(data_range=1)
            self.model = NGP(scale=0.5, log2_T=15).init_grid_buffers()
            self.losses = []

        def configure_optimizers(self):
            self.register_buffer("directions", dirs.to(self.device)); self.register_buffer("poses", train_poses.to(self.device))
            self.net_opt = FusedAdam(list(self.parameters()), self.hparams.lr, eps=1e-15)
            return [self.net_opt], [torch.optim.lr_scheduler.CosineAnnealingLR(self.net_opt, self.hparams.num_epochs, self.hparams.lr / 30)]

        def train_dataloader(self):
            g = torch.Generator().manual_seed(0)
            out = []
            for _ in range(self.hparams.steps):
                ii = torch.randint(n_train, (self.hparams.batch_size,), generator=g)
                pi = torch.randint(res * res, (self.hparams.batch_size,), generator=g)
                out.append({"img_idxs": ii, "pix_idxs": pi, "rgb": gt[ii, pi]})
            return out

        def val_dataloader(self):
            return [{"pose": test_poses[0], "img_idxs": 0, "rgb": gt_test}]

        def on_train_start(self):
            self.model.mark_invisible_cells(K.to(self.device), self.poses, (res, res))

        def training_step(self, batch, batch_nb, *args):
            if self.global_step % 16 == 0:
                self.model.update_density_grid(0.01 * MAX_SAMPLES / 3 ** 0.5, warmup=self.global_step < 256)
            rays_o, rays_d = syn.get_rays(self.directions[batch["pix_idxs"]], self.poses[batch["img_idxs"]])
            results = render(self.model, rays_o, rays_d)
            loss = sum(v.mean() for v in self.loss(results, batch).values())
            with torch.no_grad():
                self.train_psnr(results["rgb"], batch["rgb"])
            self.log("train/loss", loss); self.log("train/psnr", self.train_psnr, prog_bar=True)
            self.losses.append(float(loss.detach()))
            return loss

        def validation_step(self, batch, batch_nb):
            rays_o, rays_d = syn.get_rays(self.directions, batch["pose"])
            results = render(self.model, rays_o, rays_d, test_time=True)
            self.val_psnr(results["rgb"], batch["rgb"])
            v = self.val_psnr.compute(); self.val_psnr.reset()
            return {"psnr": v}

        def validation_epoch_end(self, outputs):
            self.log("test/psnr", torch.stack([o["psnr"] for o in outputs]).mean(), prog_bar=True)
            self.test_psnr = float(torch.stack([o["psnr"] for o in outputs]).mean())

    hp = argparse.Namespace(lr=1e-2, num_epochs=1, batch_size=1024, steps=96)
    system = System(hp)
    ck = ModelCheckpoint(dirpath=os.path.join(tmp_path, "ckpts"), filename="{epoch:d}", save_weights_only=True,
                         every_n_epochs=1, save_on_train_epoch_end=True, save_top_k=-1)
    trainer = Trainer(max_epochs=1, check_val_every_n_epoch=1, callbacks=[ck, TQDMProgressBar(refresh_rate=1)],
                      logger=TensorBoardLogger(save_dir=os.path.join(tmp_path, "logs"), name="t", default_hp_metric=False),
                      enable_model_summary=False, accelerator="gpu", devices=1, strategy=None, num_sanity_val_steps=0,
                      precision=16)
    trainer.fit(system, ckpt_path=None)
    assert trainer.global_step == 96
    first, last = sum(system.losses[:8]) / 8, sum(system.losses[-8:]) / 8
    assert last < 0.7 * first, (first, last)                                    # it learns through the autograd path
    assert system.test_psnr > 12.0, system.test_psnr
    path = os.path.join(tmp_path, "ckpts", "epoch=0.ckpt")
    assert os.path.exists(path)
    slim = slim_ckpt(path)
    assert "model.xyz_encoder.params" in slim and "model.density_grid" not in slim and "directions" not in slim
    fresh = NGP(scale=0.5, log2_T=15).init_grid_buffers().to("cuda")
    load_ckpt(fresh, path)
    torch.testing.assert_close(fresh.xyz_encoder.params, system.model.xyz_encoder.params)
    assert torch.equal(fresh.density_bitfield, system.model.density_bitfield)
