"""Minimal stand-in for the pytorch-lightning 1.6 surface that ngp_pl/train.py uses (train.py:35-39,54-58,253-282):
LightningModule (save_hyperparameters, log, device, global_step, current_epoch), Trainer.fit with automatic
optimisation (one training_step + backward + step per optimiser and batch, epoch-interval LR schedulers, fp16
autocast + GradScaler for precision=16), a validation loop every `check_val_every_n_epoch` epochs, callbacks
(ModelCheckpoint, TQDMProgressBar) and a scalar logger.  `devices > 1` (train.py:260-263: DDPPlugin) is DDP over
torch.distributed with ONE PROCESS PER GPU started by torchrun (Lightning 1.6 accepts that launch too): parameters and
buffers are broadcast from rank 0, every rank draws its own batches, gradients are averaged over the ranks before the
optimiser step, callbacks / logging run on rank 0.  (The fast data-parallel path of this repository is `NGPTrainer`
under torchrun)

Not a general Lightning replacement -- just enough for the reference's training scripts to run unchanged.
"""
import argparse
import os
import time

import torch
from torch import nn

__version__ = "1.6.5+b2n-shim"


def seed_everything(seed):
    import random
    import numpy as np
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    return seed


class LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.trainer = None
        self._logged, self._prog = {}, {}
        self.hparams = argparse.Namespace()

    # ---- what user code calls
    def save_hyperparameters(self, hparams=None, **kwargs):
        if hparams is None:
            hparams = kwargs
        self.hparams = hparams if isinstance(hparams, argparse.Namespace) else argparse.Namespace(**dict(hparams))

    def log(self, name, value, prog_bar=False, **kwargs):
        self._logged[name] = value
        if prog_bar:
            self._prog[name] = value

    def log_dict(self, d, **kwargs):
        for k, v in d.items():
            self.log(k, v, **kwargs)

    def print(self, *a, **k):
        print(*a, **k)

    @property
    def device(self):
        for t in list(self.parameters()) + list(self.buffers()):
            return t.device
        return torch.device("cpu")

    @property
    def global_step(self):
        return self.trainer.global_step if self.trainer is not None else 0

    @property
    def current_epoch(self):
        return self.trainer.current_epoch if self.trainer is not None else 0

    def get_progress_bar_dict(self):
        return {"v_num": 0, **{k: _scalar(v) for k, v in self._prog.items()}}

    # ---- hooks with Lightning's defaults
    def setup(self, stage=None): pass
    def configure_optimizers(self): raise NotImplementedError
    def train_dataloader(self): raise NotImplementedError
    def val_dataloader(self): return None
    def on_train_start(self): pass
    def on_train_end(self): pass
    def on_validation_start(self): pass
    def validation_epoch_end(self, outputs): pass
    def training_epoch_end(self, outputs): pass


def _scalar(v):
    if hasattr(v, "compute"):                                # a torchmetrics Metric
        v = v.compute()
    if torch.is_tensor(v):
        v = v.detach().float().mean().item()
    return float(v)


def _to_device(batch, device):
    if torch.is_tensor(batch):
        return batch.to(device, non_blocking=True)
    if isinstance(batch, dict):
        return {k: _to_device(v, device) for k, v in batch.items()}
    if isinstance(batch, (list, tuple)):
        return type(batch)(_to_device(v, device) for v in batch)
    try:
        import numpy as np
        if isinstance(batch, np.ndarray):
            return torch.as_tensor(batch).to(device, non_blocking=True)
    except ImportError:
        pass
    return batch


class Trainer:
    def __init__(self, max_epochs=1, check_val_every_n_epoch=1, callbacks=None, logger=None, accelerator=None,
                 devices=1, strategy=None, num_sanity_val_steps=0, precision=32, enable_model_summary=True,
                 log_every_n_steps=50, limit_train_batches=None, limit_val_batches=None, **kwargs):
        self.world = int(devices) if isinstance(devices, int) else 1
        self.rank = 0
        self.max_epochs, self.check_val = max_epochs, check_val_every_n_epoch
        self.callbacks, self.logger = list(callbacks or []), logger
        self.accelerator, self.precision = accelerator, int(precision) if str(precision).isdigit() else 32
        self.sanity_val = num_sanity_val_steps
        self.log_every, self.limit_train, self.limit_val = log_every_n_steps, limit_train_batches, limit_val_batches
        self.global_step = self.current_epoch = 0
        self.model = None

    # ---- DDP over torch.distributed (devices > 1, one process per device under torchrun)
    @property
    def is_global_zero(self):
        return self.rank == 0

    def _init_ddp(self):
        import torch.distributed as dist
        if self.world <= 1:
            return None
        if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) != self.world:
            raise RuntimeError(f"devices={self.world}: start one process per device, e.g. `python -m torch.distributed.run "
                               f"--nproc-per-node {self.world} --master-addr 127.0.0.1 train.py ...` (this stand-in does "
                               "not spawn the workers itself)")
        if not dist.is_initialized():
            dist.init_process_group("nccl" if (self.accelerator == "gpu" and torch.cuda.is_available()) else "gloo")
        self.rank = dist.get_rank()
        return dist

    def _sync_module(self, dist, model):
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)

    def _average_grads(self, dist, optimizer):
        for group in optimizer.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    dist.all_reduce(p.grad)
                    p.grad.div_(self.world)

    # ---- loops
    def fit(self, model, ckpt_path=None):
        self.model, model.trainer = model, self
        dist = self._init_ddp()
        local = int(os.environ.get("LOCAL_RANK", 0)) if dist is not None else 0
        device = torch.device("cuda", local) if (self.accelerator == "gpu" and torch.cuda.is_available()) else torch.device("cpu")
        if self.accelerator == "gpu" and device.type != "cuda":
            raise RuntimeError("accelerator='gpu' requested but CUDA is not available")
        if device.type == "cuda":
            torch.cuda.set_device(device)
        if dist is not None and not self.is_global_zero:      # rank 0 owns checkpoints, progress and logs
            self.callbacks, self.logger = [], None
        model.to(device)
        model.setup("fit")
        opt = model.configure_optimizers()
        model.to(device)                                     # buffers / parameters registered in configure_optimizers
        optimizers, schedulers = _split_optimizers(opt)
        if ckpt_path:
            state = torch.load(ckpt_path, map_location="cpu")
            model.load_state_dict(state.get("state_dict", state), strict=False)
            self.current_epoch = int(state.get("epoch", -1)) + 1 if "epoch" in state else 0
            self.global_step = int(state.get("global_step", 0))
        if dist is not None:
            self._sync_module(dist, model)                   # every replica starts from rank 0's weights and buffers
        amp = self.precision == 16 and device.type == "cuda"
        scaler = torch.amp.GradScaler("cuda", enabled=amp) if hasattr(torch.amp, "GradScaler") else \
            torch.cuda.amp.GradScaler(enabled=amp)
        train_loader, val_loader = model.train_dataloader(), model.val_dataloader()
        for cb in self.callbacks:
            cb.on_fit_start(self, model)
        if self.sanity_val == -1 and val_loader is not None:  # Lightning: run the whole validation set first
            self._validate(model, val_loader, device)
        model.on_train_start()
        start_epoch = self.current_epoch
        for epoch in range(start_epoch, self.max_epochs):
            self.current_epoch = epoch
            model.train()
            for cb in self.callbacks:
                cb.on_train_epoch_start(self, model)
            for batch_idx, batch in enumerate(train_loader):
                if self.limit_train is not None and batch_idx >= self.limit_train:
                    break
                batch = _to_device(batch, device)
                for oi, optimizer in enumerate(optimizers):
                    args = (batch, batch_idx) + ((oi,) if len(optimizers) > 1 else ())
                    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                        loss = model.training_step(*args)
                    if isinstance(loss, dict):
                        loss = loss["loss"]
                    optimizer.zero_grad(set_to_none=True)
                    scaler.scale(loss).backward()
                    if dist is not None:
                        self._average_grads(dist, optimizer)  # what DistributedDataParallel's reducer does
                    scaler.step(optimizer)
                    scaler.update()
                self.global_step += 1
                if self.global_step % self.log_every == 0 and self.logger is not None:
                    self.logger.log_metrics({k: _scalar(v) for k, v in model._logged.items()}, self.global_step)
                for cb in self.callbacks:
                    cb.on_train_batch_end(self, model, batch_idx)
            for sch in schedulers:
                sch.step()
            for cb in self.callbacks:
                cb.on_train_epoch_end(self, model)
            if val_loader is not None and self.check_val and (epoch + 1) % self.check_val == 0:
                self._validate(model, val_loader, device)
        model.on_train_end()
        for cb in self.callbacks:
            cb.on_fit_end(self, model)
        if self.logger is not None:
            self.logger.finalize()

    @torch.no_grad()
    def _validate(self, model, loader, device):
        model.eval()
        model.on_validation_start()
        outputs = []
        for batch_idx, batch in enumerate(loader):
            if self.limit_val is not None and batch_idx >= self.limit_val:
                break
            outputs.append(model.validation_step(_to_device(batch, device), batch_idx))
        model.validation_epoch_end(outputs)
        if self.logger is not None:
            self.logger.log_metrics({k: _scalar(v) for k, v in model._logged.items()}, self.global_step)
        for cb in self.callbacks:
            cb.on_validation_end(self, model)
        model.train()

    def validate(self, model, ckpt_path=None):
        self.model, model.trainer = model, self
        device = torch.device("cuda") if (self.accelerator == "gpu" and torch.cuda.is_available()) else torch.device("cpu")
        model.to(device); model.setup("validate")
        if ckpt_path:
            state = torch.load(ckpt_path, map_location="cpu")
            model.load_state_dict(state.get("state_dict", state), strict=False)
        self._validate(model, model.val_dataloader(), device)

    def save_checkpoint(self, path, weights_only=False):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save({"state_dict": self.model.state_dict(), "epoch": self.current_epoch, "global_step": self.global_step,
                    "pytorch-lightning_version": __version__}, path)


def _split_optimizers(opt):
    """configure_optimizers may return one optimiser, a list, or (optimisers, schedulers)."""
    if isinstance(opt, tuple) and len(opt) == 2 and isinstance(opt[0], (list, tuple)):
        optimizers, schedulers = list(opt[0]), list(opt[1])
    elif isinstance(opt, (list, tuple)):
        optimizers, schedulers = list(opt), []
    else:
        optimizers, schedulers = [opt], []
    schedulers = [s["scheduler"] if isinstance(s, dict) else s for s in schedulers]
    return optimizers, schedulers
