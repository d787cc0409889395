"""Checkpoint helpers with the reference's names and behaviour (ngp_pl/utils.py:4-38).

Checkpoint layout: a Lightning checkpoint stores the NGP state under the prefix 'model.'; the
slim form drops the training-only buffers `directions`, `model.density_grid`, `model.grid_coords` (and `poses` unless
asked for) and any `val_lpips*` entries.  `NGP.state_dict()` of this package has the same keys and shapes.
"""
import torch

# Flat-parameter layouts of the tinycudann modules (checkpoint key `xyz_encoder.params` / `rgb_net.params`).  This
# package's own layout ("b2n", documented in tinycudann.py): network weights first, layer by layer, each a row-major
# (out, in) matrix with the input width padded to a multiple of 16 and the output layer padded to 16 rows, then the hash
# table level by level (2 features per entry).  Upstream tiny-cuda-nn is not in the reference tree (.gitignore:25), so
# its layout cannot be verified here; from its published sources it is the same ordering (FullyFusedMLP row-major
# weight matrices, NetworkWithInputEncoding = network parameters followed by encoding parameters), hence "tcnn" is
# registered as the identity.  A checkpoint from a build that differs registers its own pair of converters here.
PARAM_LAYOUTS = {"b2n": (lambda name, p: p, lambda name, p: p), "tcnn": (lambda name, p: p, lambda name, p: p)}


def register_param_layout(name, to_b2n, from_b2n):
    """to_b2n(key, flat_params) / from_b2n(key, flat_params): convert one flat `*.params` tensor of a checkpoint."""
    PARAM_LAYOUTS[name] = (to_b2n, from_b2n)


def convert_param_layout(state, layout, to_b2n=True):
    fn = PARAM_LAYOUTS[layout][0 if to_b2n else 1]
    return {k: (fn(k, v) if (k == "params" or k.endswith(".params")) else v) for k, v in state.items()}


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


def load_ckpt(model, ckpt_path, model_name="model", prefixes_to_ignore=(), param_layout="b2n"):
    """utils.py:19-25; param_layout names the flat-parameter layout the checkpoint was written with (PARAM_LAYOUTS)."""
    if not ckpt_path:
        return
    state = model.state_dict()
    state.update(convert_param_layout(extract_model_state_dict(ckpt_path, model_name, prefixes_to_ignore), param_layout))
    model.load_state_dict(state)


def slim_ckpt(ckpt_path, save_poses=False):
    state = torch.load(ckpt_path, map_location="cpu")["state_dict"]
    drop = list(_TRAIN_ONLY) + ([] if save_poses else ["poses"]) + [k for k in state if k.startswith("val_lpips")]
    for key in drop:
        state.pop(key, None)
    return state
