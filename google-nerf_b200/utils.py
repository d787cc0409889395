"""Checkpoint helpers with the reference's names and behaviour (ngp_pl/utils.py:4-38).

Checkpoint layout: a Lightning checkpoint stores the NGP state under the prefix 'model.'; the
slim form drops the training-only buffers `directions`, `model.density_grid`, `model.grid_coords` (and `poses` unless
asked for) and any `val_lpips*` entries.  `NGP.state_dict()` of this package has the same keys and shapes.
"""
import torch

_TRAIN_ONLY = ("directions", "model.density_grid", "model.grid_coords")


def extract_model_state_dict(ckpt_path, model_name="model", prefixes_to_ignore=()):
    ckpt = torch.load(ckpt_path, map_location="cpu")
    ckpt = ckpt.get("state_dict", ckpt)                      # Lightning checkpoint or a bare state dict
    head = model_name + "."
    out = {}
    for key, value in ckpt.items():
        if not key.startswith(model_name):
            continue
        sub = key[len(head):]
        if not any(sub.startswith(p) for p in prefixes_to_ignore):
            out[sub] = value
    return out


def load_ckpt(model, ckpt_path, model_name="model", prefixes_to_ignore=()):
    if not ckpt_path:
        return
    state = model.state_dict()
    state.update(extract_model_state_dict(ckpt_path, model_name, prefixes_to_ignore))
    model.load_state_dict(state)


def slim_ckpt(ckpt_path, save_poses=False):
    state = torch.load(ckpt_path, map_location="cpu")["state_dict"]
    drop = list(_TRAIN_ONLY) + ([] if save_poses else ["poses"]) + [k for k in state if k.startswith("val_lpips")]
    for key in drop:
        state.pop(key, None)
    return state
