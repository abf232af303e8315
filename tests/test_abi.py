"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/t2p.h
declares, and the parameter tree matches the reference's (no compute calls: there is no GPU here)."""
import ctypes as C
import json
import os
import re

import pytest
import torch

from tests.cfgs import tiny_cfg
from text2protein_b200 import _lib, load_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "t2p.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(t2p_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libt2p.so does not export {n}"
        assert n in _lib.SIGNATURES, f"_lib.SIGNATURES lacks {n}"
    assert lib.t2p_abi_version() == _lib.ABI_VERSION == 3


def test_struct_layouts_match_header_field_counts():
    # a cheap guard against the ctypes mirrors drifting from the C structs
    text = open(os.path.join(ROOT, "include", "t2p.h")).read()
    for cname, cls in [("t2p_unet_cfg", _lib.UnetCfg), ("t2p_step_args", _lib.StepArgs),
                       ("t2p_run_args", _lib.RunArgs), ("t2p_conv_args", _lib.ConvArgs)]:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = [f for stmt in body.split(";") for f in stmt.split(",") if f.strip()]
        assert len(fields) == len(cls._fields_), (cname, len(fields), len(cls._fields_))


def test_struct_sizes_and_field_offsets_match_the_compiled_library():
    """sizeof and offsetof of every argument struct as compiled into libt2p.so against the ctypes mirrors: a
    reordered or retyped field fails here (and at load time, _lib.check_layout)."""
    lib = _lib.lib()
    for which, cls in _lib.STRUCTS.items():
        assert lib.t2p_sizeof(which) == C.sizeof(cls), cls.__name__
        n = len(cls._fields_)
        offs = (C.c_int32 * (n + 4))()
        assert lib.t2p_struct_layout(which, offs, n + 4) == n, cls.__name__
        assert [getattr(cls, f).offset for f, _ in cls._fields_] == list(offs[:n]), cls.__name__
    assert lib.t2p_sizeof(99) == -1

    class Drifted(C.Structure):  # the guard does catch a retyped field
        _fields_ = [(f, (C.c_int64 if f == "snr" else t)) for f, t in _lib.RunArgs._fields_]

    saved = _lib.STRUCTS[2]
    _lib.STRUCTS[2] = Drifted
    try:
        with pytest.raises(_lib.NativeError):
            _lib.check_layout(lib)
    finally:
        _lib.STRUCTS[2] = saved


def _tree(name):
    with open(os.path.join(ROOT, "tests", "golden", f"param_tree_{name}.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("which", ["tiny5", "cond_length", "cond_ss_inpainting", "no_cond", "test_config",
                                   "test_config_large"])
def test_parameter_tree_matches_reference(which):
    """All five BASELINE configurations (705 / 705 / 705 / 1065 / 1415 state_dict keys) + the tiny test network."""
    from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel
    cfg = tiny_cfg(5) if which == "tiny5" else load_config(which, device="cpu")
    tree = _tree(which)
    model = UNetModel(cfg)
    sd = model.state_dict()
    assert [k for k in sd] == [k for k, _, _ in tree["state_dict"]]
    assert [list(v.shape) for v in sd.values()] == [s for _, s, _ in tree["state_dict"]]
    assert [str(v.dtype) for v in sd.values()] == [d for _, _, d in tree["state_dict"]]
    assert [k for k, _ in model.named_parameters()] == tree["parameters"]  # EMA list is positional


def test_get_model_keeps_dataparallel_prefix():
    from text2protein_b200.score_sde_pytorch.utils import get_model
    cfg = tiny_cfg(5)
    model = get_model(cfg)
    assert all(k.startswith("module.") for k in model.state_dict())


def test_missing_n_heads_raises_like_reference():
    from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel
    cfg = load_config("cond_length", device="cpu", inject_missing=False)
    with pytest.raises(AttributeError):
        UNetModel(cfg)


def test_no_cpu_fallback():
    from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel
    cfg = tiny_cfg(5)
    model = UNetModel(cfg)
    x = torch.zeros(1, 5, 32, 32)
    with pytest.raises(_lib.NativeError):
        model(x, torch.zeros(1, dtype=torch.long), torch.zeros(1, 4, 64))


def test_registries_and_errors():
    from text2protein_b200.score_sde_pytorch import sampling, sde_lib
    assert sampling.get_predictor("reverse_diffusion") is sampling.ReverseDiffusionPredictor
    assert sampling.get_corrector("langevin") is sampling.LangevinCorrector
    with pytest.raises(KeyError):
        sampling.get_predictor("euler_maruyama")
    with pytest.raises(ValueError):
        sampling.register_predictor(name="reverse_diffusion")(sampling.ReverseDiffusionPredictor)

    class Other(sde_lib.SDE):
        T = 1
        def sde(self, x, t, context=None): ...
        def marginal_prob(self, x, t): ...
        def prior_sampling(self, shape): ...
        def prior_logp(self, z): ...

    with pytest.raises(NotImplementedError):
        sampling.LangevinCorrector(Other(10), lambda *a: None, 0.1, 1)


@pytest.mark.parametrize("N", [10, 100, 1000, 2000])
def test_host_tables_match_reference(golden_dir, N):
    import numpy as np
    from text2protein_b200.score_sde_pytorch import sampling, sde_lib
    from text2protein_b200.score_sde_pytorch.models import utils as mutils
    from oracle.unet_ref import AttrDict
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=100.0, N=N)
    labels, G = sampling.ve_tables(sde, 1e-5, N)
    assert np.array_equal(labels.numpy(), g[f"labels_{N}"])
    assert np.array_equal(G.numpy(), g[f"G_{N}"])
    assert np.array_equal(sde.discrete_sigmas.numpy(), g[f"discrete_sigmas_{N}"])
    cfg = AttrDict({"model": {"sigma_max": 100.0, "sigma_min": 0.01, "num_scales": N}})
    assert np.array_equal(mutils.get_sigmas(cfg), g[f"model_sigmas_{N}"])


def test_synthetic_weight_recipe_equals_oracle_copy():
    from oracle.unet_ref import rerandomize_ as ref
    from text2protein_b200.synthetic import rerandomize_ as ours
    tree = _tree("tiny5")
    shapes = {k: s for k, s, _ in tree["state_dict"]}
    for prefix in ("", "module."):
        a = [(prefix + k, torch.empty(shapes[k])) for k in tree["parameters"]]
        b = [(prefix + k, torch.empty(shapes[k])) for k in tree["parameters"]]
        ref(a, 42)
        ours(b, 42)
        assert all(torch.equal(x[1], y[1]) for x, y in zip(a, b))
