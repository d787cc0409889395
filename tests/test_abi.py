"""CPU checks of the drop-in boundary: libb2n.so builds, loads and exports every symbol include/b2n.h declares, the
ctypes signatures agree with the header's parameter counts, and the product fails loudly without CUDA tensors."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "b2n.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"B2N_API\s+[\w\s\*]+?\b(b2n_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return protos


def test_library_exports_every_declared_symbol(built_lib):
    protos = _header_prototypes()
    assert len(protos) >= 30
    lib = built_lib.lib()
    for name in protos:
        assert hasattr(lib, name), f"{name} is declared in include/b2n.h but not exported by libb2n.so"
    assert lib.b2n_version() == 100
    # every binding the Python host uses is declared in the header, with the same number of parameters
    for name, sig in built_lib._SIGS.items():
        assert name in protos, name
        assert len(sig) == protos[name], (name, len(sig), protos[name])


def test_hashgrid_layout_host_function(built_lib):
    from google_nerf_b200 import tinycudann as tc
    import numpy as np
    lay = tc.hashgrid_layout(16, 2, 19, 16, np.exp(np.log(2048 * 0.5 / 16) / 15))
    assert lay.n_params == 11420064 and lay.resolution[0] == 16 and lay.resolution[15] == 1024
    assert lay.x_offset == 0.0 and lay.x_scale == 1.0
    with pytest.raises(RuntimeError, match="bad hash grid config"):
        tc.hashgrid_layout(64, 2, 19, 16, 2.0)


def test_error_convention_and_no_cpu_fallback(built_lib):
    from google_nerf_b200 import vren
    from google_nerf_b200.models.networks import NGP
    o = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vren.ray_aabb_intersect(o, o, torch.zeros(1, 3), torch.ones(1, 3), 1)
    m = NGP(0.5, log2_T=12)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.density(torch.zeros(4, 3))
    # argument validation happens before any launch and surfaces b2n_last_error()
    with pytest.raises(RuntimeError, match="max_hits|bad sizes"):
        built_lib.call_nostream("b2n_ray_aabb_intersect", None, None, None, None, 4, 1, 0, None, None, None, None)


def test_module_contract_on_cpu(built_lib):
    """Module construction, parameter counts and checkpoint keys need no GPU."""
    from google_nerf_b200.models.networks import NGP
    m = NGP(0.5)
    sd = m.state_dict()
    assert sd["xyz_encoder.params"].numel() == 11420064 + 3072 and sd["rgb_net.params"].numel() == 7168
    assert sd["dir_encoder.params"].numel() == 0 and m.cascades == 1 and m.grid_size == 128
    big = NGP(16.0, log2_T=22)
    assert big.cascades == 6
    # test-time render path selection (rendering.py::_WholeRays.supports): the persistent whole-ray kernel for the
    # HashGrid field while its fp16 table fits the L2, the round loop otherwise and for the Frequency encoding
    from google_nerf_b200.models.rendering import _WholeRays
    assert _WholeRays.supports(m, "auto") and _WholeRays.supports(m, True) and not _WholeRays.supports(m, False)
    assert not _WholeRays.supports(big, "auto") and _WholeRays.supports(big, True)
    assert not _WholeRays.supports(NGP(0.5, encoding="Frequency"), True)
    # states cached on the module (captured graphs, ctypes structures) stay out of copies and pickles
    import copy, io
    small = NGP(0.5, log2_T=12)
    small.__dict__["_whole_rays_pool"] = {7: object()}; small.__dict__["_device_loop"] = (lambda: 0)
    twin = copy.deepcopy(small)
    assert twin.__dict__["_whole_rays_pool"] == {} and twin.__dict__["_device_loop"] is None
    assert torch.equal(twin.xyz_encoder.params, small.xyz_encoder.params)
    torch.save(small, io.BytesIO())
    assert NGP(0.5, encoding="Frequency").xyz_encoder.params.numel() == 80 * 64 + 64 * 16
    # utils.load_ckpt-style round trip (ngp_pl/utils.py:4-25): keys under "model." stripped, then load_state_dict
    ckpt = {"state_dict": {"model." + k: v.clone() for k, v in sd.items()}}
    m2 = NGP(0.5)
    m2.load_state_dict({k[len("model."):]: v for k, v in ckpt["state_dict"].items()})
    assert torch.equal(m2.xyz_encoder.params, m.xyz_encoder.params)
    from google_nerf_b200.models import rendering, custom_functions
    assert rendering.MAX_SAMPLES == 1024 and rendering.NEAR_DISTANCE == 0.05
    for name in ("RayAABBIntersector", "RaySphereIntersector", "RayMarcher", "VolumeRenderer", "TruncExp"):
        assert hasattr(custom_functions, name)
