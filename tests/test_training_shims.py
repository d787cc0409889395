"""Stand-ins for pytorch_lightning / torchmetrics / imageio and the synthetic NSVF dataset writer, on the
CPU.  Where /root/reference exists (this container, not the GPU box) the reference's own dataset class and train.py
are imported against them, unchanged."""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "google-nerf_b200", "shims")
REF = "/root/reference/ngp_pl"


@pytest.fixture()
def shims():
    added = [p for p in (ROOT, SHIMS) if p not in sys.path]
    for p in added:
        sys.path.insert(0, p)
    yield
    for p in added:
        sys.path.remove(p)


def test_torchmetrics_psnr_ssim(shims):
    from torchmetrics import PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure
    g = torch.Generator().manual_seed(0)
    a = torch.rand(2, 3, 32, 32, generator=g); b = (a + 0.05 * torch.randn(2, 3, 32, 32, generator=g)).clamp(0, 1)
    psnr = PeakSignalNoiseRatio(data_range=1)
    v1 = psnr(a[:1], b[:1]); v2 = psnr(a[1:], b[1:])
    want = lambda x, y: -10 * torch.log10(((x - y) ** 2).mean())
    torch.testing.assert_close(v1, want(a[:1], b[:1])); torch.testing.assert_close(v2, want(a[1:], b[1:]))
    torch.testing.assert_close(psnr.compute(), want(a, b))           # accumulated over both calls
    psnr.reset(); psnr(a, a * 0 + b)
    torch.testing.assert_close(psnr.compute(), want(a, b))
    ssim = StructuralSimilarityIndexMeasure(data_range=1)
    assert abs(float(ssim(a, a)) - 1.0) < 1e-5                        # identical images
    ssim.reset()
    s_noise = float(ssim(a, b)); ssim.reset()
    s_far = float(ssim(a, 1 - a))
    assert 0.2 < s_noise < 1.0 and s_far < s_noise
    # constant images: SSIM reduces to the luminance term (2 mu1 mu2 + c1) / (mu1^2 + mu2^2 + c1)
    c = StructuralSimilarityIndexMeasure(data_range=1)(torch.full((1, 1, 16, 16), 0.2), torch.full((1, 1, 16, 16), 0.6))
    assert abs(float(c) - (2 * 0.2 * 0.6 + 1e-4) / (0.04 + 0.36 + 1e-4)) < 2e-4           # (fp32 E[x^2] - mu^2 cancellation)
    with pytest.raises(RuntimeError):
        from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity
        LearnedPerceptualImagePatchSimilarity("vgg")


def test_imageio_roundtrip(shims, tmp_path):
    import imageio
    img = (np.random.RandomState(0).rand(8, 9, 3) * 255).astype(np.uint8)
    p = os.path.join(tmp_path, "a.png")
    imageio.imsave(p, img)
    assert np.array_equal(imageio.imread(p), img)
    with pytest.warns(UserWarning):
        imageio.mimsave(os.path.join(tmp_path, "v.mp4"), [img, img[::-1]], fps=30, macro_block_size=1)
    assert os.path.exists(os.path.join(tmp_path, "v.gif"))


def test_lightning_fit_loop(shims, tmp_path):
    """Automatic optimisation with two optimisers, an epoch-interval scheduler, validation every 2 epochs, a checkpoint
    named like Lightning's and a scalar log -- the behaviours ngp_pl/train.py relies on."""
    from pytorch_lightning import LightningModule, Trainer
    from pytorch_lightning.callbacks import ModelCheckpoint, TQDMProgressBar
    from pytorch_lightning.loggers import TensorBoardLogger
    from pytorch_lightning.plugins import DDPPlugin
    from pytorch_lightning.utilities.distributed import all_gather_ddp_if_available

    class Toy(LightningModule):
        def __init__(self, hp):
            super().__init__()
            self.save_hyperparameters(hp)
            self.w = torch.nn.Parameter(torch.zeros(3)); self.b = torch.nn.Parameter(torch.zeros(1))
            self.calls, self.val_runs = [], 0

        def configure_optimizers(self):
            self.register_buffer("target", torch.tensor([1.0, -2.0, 0.5]))
            o1 = torch.optim.SGD([self.w], lr=self.hparams.lr); o2 = torch.optim.SGD([self.b], lr=self.hparams.lr)
            self.o1 = o1
            return [o1, o2], [torch.optim.lr_scheduler.StepLR(o1, 1, gamma=0.5)]

        def train_dataloader(self):
            return [{"x": torch.eye(3)} for _ in range(5)]

        def val_dataloader(self):
            return [{"i": i} for i in range(3)]

        def training_step(self, batch, batch_nb, *args):
            self.calls.append((self.global_step, batch_nb) + args)
            loss = ((batch["x"] @ self.w + self.b - self.target) ** 2).mean()
            self.log("train/loss", loss, prog_bar=True); self.log("lr", self.o1.param_groups[0]["lr"])
            return loss

        def validation_step(self, batch, batch_nb):
            assert not torch.is_grad_enabled() and not self.training
            return {"v": torch.tensor(float(batch["i"]))}

        def validation_epoch_end(self, outputs):
            self.val_runs += 1
            self.log("test/v", all_gather_ddp_if_available(torch.stack([o["v"] for o in outputs])).mean())

    hp = types.SimpleNamespace(lr=0.5)
    import argparse
    m = Toy(argparse.Namespace(lr=0.5))
    ck = ModelCheckpoint(dirpath=os.path.join(tmp_path, "ck"), filename="{epoch:d}", save_weights_only=True, every_n_epochs=4,
                         save_on_train_epoch_end=True, save_top_k=-1)
    tr = Trainer(max_epochs=4, check_val_every_n_epoch=2, callbacks=[ck, TQDMProgressBar(refresh_rate=1)],
                 logger=TensorBoardLogger(save_dir=os.path.join(tmp_path, "logs"), name="exp", default_hp_metric=False),
                 enable_model_summary=False, accelerator=None, devices=1, strategy=None, num_sanity_val_steps=0,
                 precision=16, log_every_n_steps=5)
    tr.fit(m, ckpt_path=None)
    assert tr.global_step == 20 and m.global_step == 20 and m.current_epoch == 3
    assert len(m.calls) == 40 and m.calls[0] == (0, 0, 0) and m.calls[1] == (0, 0, 1)       # one call per optimiser
    assert m.val_runs == 2
    assert abs(m.o1.param_groups[0]["lr"] - 0.5 * 0.5 ** 4) < 1e-12                         # scheduler stepped per epoch
    assert float(((m.w - m.target) ** 2).sum()) < 0.5                                        # it trained
    path = os.path.join(tmp_path, "ck", "epoch=3.ckpt")
    assert ck.last_path == path and os.path.exists(path)
    state = torch.load(path)
    assert set(state["state_dict"]) == {"w", "b", "target"} and state["epoch"] == 3
    sys.path.insert(0, ROOT)
    from google_nerf_b200.utils import slim_ckpt
    assert "w" in slim_ckpt(path)
    log = open(os.path.join(tmp_path, "logs", "exp", "metrics.csv")).read()
    assert "train/loss" in log and "test/v,1.0" in log
    assert m.get_progress_bar_dict().keys() >= {"v_num", "train/loss"}
    # devices > 1 is DDP with one process per device: outside such a launch fit() says how to start it
    with pytest.raises(RuntimeError, match="one process per device"):
        Trainer(devices=2, strategy=DDPPlugin(find_unused_parameters=False)).fit(Toy(argparse.Namespace(lr=0.5)))


def test_synthetic_nsvf_dataset_layout(shims, tmp_path):
    from google_nerf_b200 import synthetic as syn
    root = os.path.join(tmp_path, "Synthetic_NSVF", "Analytic")
    train_poses, test_poses = syn.write_nsvf_dataset(root, n_train=3, n_test=2, res=40)
    assert sorted(os.listdir(os.path.join(root, "rgb"))) == ["0_0000.png", "0_0001.png", "0_0002.png", "2_0003.png", "2_0004.png"]
    bbox = np.loadtxt(os.path.join(root, "bbox.txt"))[:6].reshape(2, 3)
    scale_ds = (bbox[1] - bbox[0]).max() / 2 * 1.05
    raw = np.loadtxt(os.path.join(root, "pose", "0_0001.txt"))
    np.testing.assert_allclose(raw[:3, 3] / (2 * scale_ds), train_poses[1][:, 3].numpy(), rtol=1e-6)   # nsvf.py:86-87
    with pytest.raises(ValueError):
        syn.write_nsvf_dataset(os.path.join(tmp_path, "elsewhere"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_reference_dataset_and_train_script_import_unchanged(shims, tmp_path, monkeypatch):
    """ngp_pl's NSVFDataset reads the synthetic dataset, and ngp_pl/train.py imports and builds its LightningModule
    against the stand-ins (vren / tinycudann / apex / kornia / torch_scatter / pytorch_lightning / torchmetrics /
    imageio) without a single edit."""
    from google_nerf_b200 import synthetic as syn
    root = os.path.join(tmp_path, "Synthetic_NSVF", "Analytic")
    train_poses, _ = syn.write_nsvf_dataset(root, n_train=3, n_test=1, res=40)
    monkeypatch.syspath_prepend(REF)
    for name in [k for k in sys.modules if k.split(".")[0] in ("datasets", "models", "train", "opt", "losses", "utils", "metrics")]:
        monkeypatch.delitem(sys.modules, name)
    from datasets.nsvf import NSVFDataset
    ds = NSVFDataset(root, split="train", downsample=40 / 800)
    assert ds.img_wh == (40, 40) and ds.rays.shape == (3, 1600, 3) and ds.poses.shape == (3, 3, 4)
    torch.testing.assert_close(ds.poses, train_poses, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ds.directions, syn.directions(40, 40, syn.intrinsics(40, 40)), rtol=1e-6, atol=1e-6)
    ro, rd = syn.get_rays(ds.directions, ds.poses[2])
    assert float((ds.rays[2] - syn.shade(ro, rd, 0.5)).abs().max()) <= 1.0 / 255 + 1e-6    # 8-bit PNG
    ds.batch_size = 64
    b = ds[0]
    assert set(b) == {"rgb", "img_idxs", "pix_idxs"} and b["rgb"].shape == (64, 3)

    monkeypatch.setattr(sys, "argv", ["train.py", "--root_dir", root, "--exp_name", "t", "--downsample", str(40 / 800),
                                      "--num_epochs", "1", "--batch_size", "64", "--no_save_test"])
    import train as ref_train
    hparams = ref_train.get_opts()
    monkeypatch.setattr(ref_train, "hparams", hparams, raising=False)
    system = ref_train.NeRFSystem(hparams)
    assert system.model.density_grid.shape == (1, 128 ** 3) and system.model.grid_coords.shape == (128 ** 3, 3)
    assert system.model.xyz_encoder.params.numel() > 0
    system.setup("fit")
    assert len(system.train_dataset.poses) == 3 and system.train_dataset.batch_size == 64
    for m in [k for k in list(sys.modules) if k.split(".")[0] in ("datasets", "models", "train", "opt", "losses", "utils", "metrics")]:
        sys.modules.pop(m, None)
