"""Callbacks used at ngp_pl/train.py:253-259."""
import os
import sys
import time


class Callback:
    def on_fit_start(self, trainer, module): pass
    def on_fit_end(self, trainer, module): pass
    def on_train_epoch_start(self, trainer, module): pass
    def on_train_epoch_end(self, trainer, module): pass
    def on_train_batch_end(self, trainer, module, batch_idx): pass
    def on_validation_end(self, trainer, module): pass


class ModelCheckpoint(Callback):
    """Saves `{dirpath}/{filename}.ckpt` every `every_n_epochs` epochs at the end of the training epoch.  `filename`
    understands Lightning's '{epoch:d}' style: the key is kept, i.e. 'epoch=29.ckpt' (train.py:253-258,289)."""

    def __init__(self, dirpath=None, filename="{epoch:d}", save_weights_only=False, every_n_epochs=1,
                 save_on_train_epoch_end=True, save_top_k=-1, **kwargs):
        self.dirpath, self.filename, self.every = dirpath or "checkpoints", filename, max(1, int(every_n_epochs or 1))
        self.weights_only = save_weights_only
        self.last_path = None

    def format_name(self, epoch, step):
        name = self.filename
        for key, val in (("epoch", epoch), ("step", step)):
            for spec in ("{%s:d}" % key, "{%s}" % key):
                name = name.replace(spec, f"{key}={val}")
        return name + ".ckpt"

    def on_train_epoch_end(self, trainer, module):
        if (trainer.current_epoch + 1) % self.every == 0:
            self.last_path = os.path.join(self.dirpath, self.format_name(trainer.current_epoch, trainer.global_step))
            trainer.save_checkpoint(self.last_path, self.weights_only)


class TQDMProgressBar(Callback):
    """A line of text every `refresh_rate * 100` steps (a real tqdm bar would force a metric sync every step)."""

    def __init__(self, refresh_rate=1, **kwargs):
        self.every = max(1, int(refresh_rate)) * 100
        self.t0 = None

    def on_train_epoch_start(self, trainer, module):
        self.t0 = self.t0 or time.time()

    def on_train_batch_end(self, trainer, module, batch_idx):
        if trainer.global_step % self.every == 0:
            items = module.get_progress_bar_dict()
            text = " ".join(f"{k}={v:.4g}" for k, v in items.items() if k != "v_num")
            print(f"epoch {trainer.current_epoch} step {trainer.global_step} [{time.time() - self.t0:.1f}s] {text}",
                  file=sys.stderr, flush=True)
