"""Pins the oracle (oracle/) against the committed outputs of the unmodified reference (tests/golden)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sampler_ref, unet_ref
from oracle.philox_ref import philox4x32_10, philox_normal
from tests.cfgs import synthetic_condition, synthetic_inputs, tiny_cfg


def _tree(golden_dir, name):
    with open(os.path.join(golden_dir, f"param_tree_{name}.json")) as f:
        return json.load(f)


def _tiny_sd(golden_dir, cfg):
    # tiny5 and tiny8 differ only in the first / last conv shapes
    tree = _tree(golden_dir, "tiny5")
    c = cfg.data.num_channels
    fix = {"pre_conv.weight": [cfg.model.nf, c, 3, 3], "out.2.weight": [c, cfg.model.nf, 3, 3], "out.2.bias": [c]}
    tree["state_dict"] = [[k, fix.get(k, s), d] for k, s, d in tree["state_dict"]]
    return unet_ref.state_dict_from_tree(tree, cfg, 42)


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10
    def kat(c, k):
        r = philox4x32_10(np.array([c], dtype=np.uint32), np.array([k], dtype=np.uint32))[0]
        return [int(v) for v in r]

    assert kat([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert kat([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert kat([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_normal_moments_and_offsets():
    n = philox_normal(7, 3, 0, 1 << 18)
    assert abs(float(n.mean())) < 0.01 and abs(float(n.std()) - 1.0) < 0.01
    # global indexing: a shard starting at element 4096 sees the same numbers
    assert np.array_equal(philox_normal(7, 3, 4096, 1024), n[4096:5120])


@pytest.mark.parametrize("c", [5, 8])
def test_unet_matches_reference(golden_dir, c):
    cfg = tiny_cfg(c)
    g = np.load(os.path.join(golden_dir, f"unet_tiny{c}.npz"))
    sd = _tiny_sd(golden_dir, cfg)
    x, labels, ctx = synthetic_inputs(cfg, 2, 8)
    taps = {}
    out = unet_ref.unet_forward(sd, cfg, x, labels, ctx, taps=taps)
    assert out.dtype == torch.float64
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-5, atol=1e-6)
    out_unit = unet_ref.unet_forward(sd, cfg, x, labels, ctx / 0.02)
    np.testing.assert_allclose(out_unit.numpy(), g["out_unit"], rtol=1e-5, atol=1e-6)
    for k in g.files:
        if k.startswith("tap:"):
            np.testing.assert_allclose(taps[k[4:]].numpy(), g[k], rtol=1e-5, atol=2e-5, err_msg=k)


def test_param_tree_cond_length_counts(golden_dir):
    tree = _tree(golden_dir, "cond_length")
    assert len(tree["state_dict"]) == 705 and len(tree["parameters"]) == 704


@pytest.mark.parametrize("N", [10, 100, 1000, 2000])
def test_tables(golden_dir, N):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    sde = sampler_ref.VESDERef(0.01, 100.0, N)
    cfg = unet_ref.AttrDict({"model": {"sigma_max": 100.0, "sigma_min": 0.01, "num_scales": N}})
    assert np.array_equal(unet_ref.get_sigmas(cfg), g[f"model_sigmas_{N}"])
    assert np.array_equal(sde.discrete_sigmas.numpy(), g[f"discrete_sigmas_{N}"])
    ts = torch.linspace(sde.T, 1e-5, sde.N)
    assert np.array_equal(ts.numpy(), g[f"timesteps_{N}"])
    labels = np.array([sde.labels(torch.ones(1) * ts[i]).item() for i in range(N)])
    assert np.array_equal(labels, g[f"labels_{N}"])
    assert np.array_equal(labels, np.arange(N))  # label == loop index
    G = np.array([sde.discretize_G(torch.ones(1) * ts[i]).item() for i in range(N)], dtype=np.float32)
    assert np.array_equal(G, g[f"G_{N}"])
    assert np.array_equal(g[f"tidx_{N}"], N - 1 - np.arange(N))


@pytest.mark.parametrize("name,c,kinds", [("sampler_tiny5_length", 5, ["length"]),
                                          ("sampler_tiny8_all", 8, ["length", "ss", "inpainting"]),
                                          ("sampler_tiny8_nocond", 8, [])])
def test_sampler_matches_reference(golden_dir, name, c, kinds):
    cfg = tiny_cfg(c)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = _tiny_sd(golden_dir, cfg)
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, kinds) if kinds else {}
    sde = sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    shape = (2, c, cfg.data.max_res_num, cfg.data.max_res_num)
    model = lambda x, lab, cx: unet_ref.unet_forward(sd, cfg, x, lab, cx)
    sample, nfe = sampler_ref.pc_sampler_ref(sde, model, shape, cfg.sampling.snr, n_steps=1, eps=1e-5,
                                             condition=cond, context=ctx,
                                             noise_fn=sampler_ref.philox_noise_fn(2024))
    assert nfe == int(g["nfe"]) and sample.dtype == torch.float32
    np.testing.assert_allclose(sample.numpy(), g["sample"], rtol=1e-4, atol=1e-3)
    # mask handling is bit-exact (SURVEY 8c iv)
    if "length" in kinds:
        assert torch.equal(sample[:, -1], cond["length"].float())
    if "ss" in kinds:
        assert torch.equal(sample[:, 4:7], cond["ss"])
    if "inpainting" in kinds:
        keep = ~cond["inpainting"]["mask_inpaint"][:, None].expand_as(sample)
        assert torch.equal(sample[keep], cond["inpainting"]["coords_6d"][keep])


def test_vpsde_sampler_matches_reference(golden_dir):
    """VPSDE through the oracle (float time conditioning, score = -out / sqrt(1 - alpha_bar), DDPM discretisation,
    alpha-scaled Langevin step) against the reference's own pc_sampler run -- SURVEY row a9."""
    g = np.load(os.path.join(golden_dir, "sampler_vpsde_tiny5.npz"))
    N, K = int(g["N"]), int(g["K"])
    cfg = tiny_cfg(5, num_scales=N)
    sd = _tiny_sd(golden_dir, cfg)
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, ["length"])
    sde = sampler_ref.VPSDERef(cfg.model.beta_min, cfg.model.beta_max, N)
    model = lambda x, lab, cx: unet_ref.unet_forward(sd, cfg, x, lab, cx)
    sample, nfe = sampler_ref.pc_sampler_ref(sde, model, (2, 5, 32, 32), cfg.sampling.snr, n_steps=1, eps=1e-3,
                                             condition=cond, context=ctx, noise_fn=sampler_ref.philox_noise_fn(2024),
                                             num_iters=K, generic_streams=True)
    assert nfe == 2 * K and sample.dtype == torch.float32
    np.testing.assert_allclose(sample.numpy(), g["sample"], rtol=1e-4, atol=1e-3)
    assert torch.equal(sample[:, -1], cond["length"].float())


def _full_sd(golden_dir, yaml_name, cfg):
    tree = _tree(golden_dir, yaml_name[:-4])
    return unet_ref.state_dict_from_tree(tree, cfg, 42)


@pytest.mark.parametrize("case", ["cond_length_L77", "cond_ss_inpainting", "test_config_large"])
def test_unet_matches_reference_at_baseline_size(golden_dir, case):
    """The oracle at the real architectures (N = 128 nf = 128; C = 8; N = 256 with d_head = 128 and L = 512, 863 M
    parameters) against the reference's output -- what the full-size GPU parity tests lean on."""
    from tests.cfgs import FULLSIZE_CASES, fullsize_inputs
    from text2protein_b200 import load_config
    fname, B, L = FULLSIZE_CASES[case]
    cfg = load_config(fname, device="cpu")
    g = np.load(os.path.join(golden_dir, f"unet_full_{case}.npz"))
    sd = _full_sd(golden_dir, fname, cfg)
    x, labels, ctx = fullsize_inputs(cfg, B, L)
    # the torch generator drew the same inputs and weights as on the machine that made the golden
    np.testing.assert_allclose([x.double().sum().item(), ctx.double().sum().item(), float(labels.sum())],
                               g["in_sums"], rtol=1e-12)
    out = unet_ref.unet_forward(sd, cfg, x, labels, ctx)
    ref = torch.from_numpy(g["out"]).double()
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 1e-5


def test_sampler_matches_reference_at_baseline_size(golden_dir):
    """K = 2 iterations of the reference pc_sampler on the real cond_length.yml network (B = 2, N = 128)."""
    from tests.cfgs import fullsize_inputs
    from text2protein_b200 import load_config
    cfg = load_config("cond_length", device="cpu")
    g = np.load(os.path.join(golden_dir, "sampler_full_cond_length.npz"))
    K = int(g["K"])
    sd = _full_sd(golden_dir, "cond_length.yml", cfg)
    _, _, ctx = fullsize_inputs(cfg, 2, 77)
    cond = synthetic_condition(cfg, 2, ["length"])
    sde = sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    model = lambda x, lab, cx: unet_ref.unet_forward(sd, cfg, x, lab, cx)
    sample, _ = sampler_ref.pc_sampler_ref(sde, model, (2, 5, 128, 128), cfg.sampling.snr, n_steps=1, eps=1e-5,
                                           condition=cond, context=ctx, noise_fn=sampler_ref.philox_noise_fn(2024),
                                           num_iters=K)
    ref = torch.from_numpy(g["sample"])
    assert ((sample - ref).abs().max() / ref.abs().max()).item() < 1e-5
    assert torch.equal(sample[:, -1], cond["length"].float())


def test_symmetrize_free_specification():
    """The opt-in symmetrisation (product extension): exactly symmetric where both positions are free, untouched
    elsewhere and in channels >= 2; a no-op on an already symmetric map."""
    g = torch.Generator().manual_seed(0)
    u = torch.randn(2, 5, 6, 6, generator=g, dtype=torch.float64)
    cm = torch.ones(2, 5, 6, 6, dtype=torch.bool)
    cm[:, :, 4:, :] = False
    cm[:, :, :, 4:] = False
    cm[0, 0, 1, 2] = False  # asymmetric hole: neither (1,2) nor (2,1) is symmetrised
    s = sampler_ref.symmetrize_free(u, cm)
    assert torch.equal(s[:, 2:], u[:, 2:])
    assert torch.equal(s[:, :2][~cm[:, :2]], u[:, :2][~cm[:, :2]])
    assert torch.equal(s[0, 0, 2, 1], u[0, 0, 2, 1]) and torch.equal(s[0, 0, 1, 2], u[0, 0, 1, 2])
    assert torch.equal(s[1, :2, :4, :4], s[1, :2, :4, :4].transpose(1, 2))
    assert torch.equal(sampler_ref.symmetrize_free(s, cm), s)
