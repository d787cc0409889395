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


def test_train_mika_step_sequence_matches_oracle(built_lib):
    """The one-step loop of ngp_pl/train_mika.py:68-173 (the SCADE-style script: no Lightning) spelled with the module
    names that script imports -- kornia.create_meshgrid3d, apex FusedAdam, torchmetrics PSNR, NGP / render / NeRFLoss --
    on ScanNet-shaped cameras; its printed loss and PSNR equal the oracle's for the same occupancy bitfield and jitter.
    (The script itself cannot travel to the GPU box: the reference tree is not there.)"""
    for p in (ROOT, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from apex.optimizers import FusedAdam
    from kornia.utils.grid import create_meshgrid3d
    from torch.optim.lr_scheduler import CosineAnnealingLR
    from torchmetrics import PeakSignalNoiseRatio
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.losses import NeRFLoss
    from google_nerf_b200.models.custom_functions import RayMarcher
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.models.rendering import MAX_SAMPLES, render
    from oracle import ngp_ref as O
    device, S, batch_size, lr, num_epochs = "cuda", 16, 2048, 1e-2, 30
    W, H = 624, 468
    K = syn.intrinsics(W, H, fx=577.87 * 624 / 640); directions = syn.directions(W, H, K).to(device)
    train_poses = syn.room_poses(18, seed=3).to(device)
    # ---- train_mika.py:68-76
    loss_func = NeRFLoss()
    train_psnr = PeakSignalNoiseRatio(data_range=1)
    model = NGP(scale=0.5, log2_T=15)
    G = model.grid_size
    model.register_buffer("density_grid", torch.zeros(model.cascades, G ** 3))
    model.register_buffer("grid_coords", create_meshgrid3d(G, G, G, False, dtype=torch.int32).reshape(-1, 3))
    model.to(device, dtype=torch.float)
    # ---- :103-121
    net_opt = FusedAdam(model.parameters(), lr, eps=1e-15)
    net_sch = CosineAnnealingLR(net_opt, num_epochs, lr / 30)
    model.mark_invisible_cells(K.to(device), train_poses, (W, H))
    model.train()
    g = torch.Generator().manual_seed(0)
    ii = torch.randint(18, (batch_size,), generator=g); pi = torch.randint(W * H, (batch_size,), generator=g)
    ro_cpu, rd_cpu = syn.get_rays(directions.cpu()[pi], train_poses.cpu()[ii])
    batch = {"rgb": syn.scene_shade(ro_cpu, rd_cpu, syn.ROOM)[0].to(device), "img_idxs": ii.to(device), "pix_idxs": pi.to(device)}
    global_step = 0
    # ---- :130-167
    if global_step % S == 0:
        model.update_density_grid(0.01 * MAX_SAMPLES / 3 ** 0.5, warmup=global_step < 256, erode=False)
    poses = train_poses[batch["img_idxs"]]
    rays_o, rays_d = syn.get_rays(directions[batch["pix_idxs"]], poses)
    noise = torch.rand(batch_size, generator=g)
    RayMarcher.noise = noise.to(device)
    try:
        results = render(model, rays_o, rays_d, test_time=False)
    finally:
        RayMarcher.noise = None
    loss_d = loss_func(results, batch)
    loss = sum(lo.mean() for lo in loss_d.values())
    p_before, r_before = model.xyz_encoder.params.detach().clone(), model.rgb_net.params.detach().clone()
    net_opt.zero_grad()
    loss.backward()
    net_opt.step()
    net_sch.step()
    with torch.no_grad():
        psnr = float(train_psnr(results["rgb"], batch["rgb"]))
    s_per_ray = float(results["total_samples"]) / len(rays_o)
    # ---- the oracle on the same weights (before the step), bitfield and jitter
    ref = O.NGPRef(0.5, log2_T=15)
    with torch.no_grad():
        ref.xyz_params.copy_(p_before.cpu()); ref.rgb_params.copy_(r_before.cpu())
    ref.density_bitfield = model.density_bitfield.cpu()
    res_ref = O.render(ref, ro_cpu, rd_cpu.clone(), noise=noise)
    loss_ref = O.nerf_loss(res_ref, batch["rgb"].cpu())
    mse = ((res_ref["rgb"].detach() - batch["rgb"].cpu()) ** 2).mean()
    psnr_ref = float(-10 * torch.log10(mse))
    assert int(results["total_samples"]) == res_ref["total_samples"] and s_per_ray > 1
    assert abs(loss.item() - loss_ref.item()) < 2e-3 * abs(loss_ref.item())
    assert abs(psnr - psnr_ref) < 0.02
    assert not torch.equal(p_before, model.xyz_encoder.params.detach()) and net_sch.get_last_lr()[0] < lr
