"""`DDPPlugin` is only constructed when num_gpus > 1 (train.py:271-272).  The stand-in Trainer implements the DDP
semantics itself (one process per device under torchrun: broadcast from rank 0, gradient averaging); the plugin object
only carries its options."""


class DDPPlugin:
    def __init__(self, find_unused_parameters=False, **kwargs):
        self.find_unused_parameters = find_unused_parameters
