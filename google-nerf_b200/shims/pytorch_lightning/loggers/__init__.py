"""Scalar logger with TensorBoardLogger's constructor (train.py:261-263): writes `<save_dir>/<name>/metrics.csv`
(step, name, value) -- TensorBoard itself is not needed offline."""
import os


class TensorBoardLogger:
    def __init__(self, save_dir="logs", name="default", version=None, default_hp_metric=True, **kwargs):
        self.log_dir = os.path.join(save_dir, name or "")
        self._fh = None

    def log_metrics(self, metrics, step):
        if self._fh is None:
            os.makedirs(self.log_dir, exist_ok=True)
            self._fh = open(os.path.join(self.log_dir, "metrics.csv"), "a")
        for k, v in metrics.items():
            self._fh.write(f"{step},{k},{v}\n")
        self._fh.flush()

    def log_hyperparams(self, *a, **k): pass

    def finalize(self, *a, **k):
        if self._fh is not None:
            self._fh.close(); self._fh = None
