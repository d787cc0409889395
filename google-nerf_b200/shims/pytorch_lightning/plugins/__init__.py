"""`DDPPlugin` is only constructed when num_gpus > 1 (train.py:271-272); multi-GPU training in this repository is
`NGPTrainer` under torchrun, so the stand-in Trainer refuses devices > 1 and this class is a placeholder."""


class DDPPlugin:
    def __init__(self, find_unused_parameters=False, **kwargs):
        self.find_unused_parameters = find_unused_parameters
