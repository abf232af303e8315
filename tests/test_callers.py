"""Callers either side of the sampling loop: condition builders, embed_tokens gather, Rosetta
pre-processing and the sampling_6d driver.  CPU part: the oracle against vectors produced by the reference's own
source (tests/golden/make_golden_callers.py).  GPU part: the device kernels, through the C ABI, against the oracle --
bit-exact (integer / byte work; the float arithmetic is single-rounding fp32 on both sides)."""
import os
import pickle as pkl

import numpy as np
import pytest
import torch

from oracle import callers_ref
from tests.cfgs import tiny_cfg, with_device


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "callers.npz"))


# ------------------------------------------------------------------------------------------------- oracle vs reference
def test_oracle_length_masks_match_reference(G):
    ref = torch.from_numpy(G["all_lengths_5_24_b3"])
    assert torch.equal(callers_ref.get_mask_all_lengths_ref(5, 24, 3), ref)
    assert torch.equal(callers_ref.length_mask_ref(list(range(5, 25)), 24), ref[:, 0])


def test_oracle_selected_masks_match_reference(G):
    for i in range(3):
        info = str(G[f"selected_{i}_info"])
        assert torch.equal(callers_ref.selected_mask_ref(2, 24, info), torch.from_numpy(G[f"selected_{i}"]))


def test_oracle_restraints_match_reference(G):
    for k in range(3):
        L, npz = callers_ref.restraints_ref(G[f"rosetta_{k}_in"])
        assert L == int(G[f"rosetta_{k}_L"])
        for name, v in npz.items():
            ref = G[f"rosetta_{k}_{name}"]
            assert v.dtype == ref.dtype and np.array_equal(v, ref), name


def test_oracle_restraints_rejects_improper_mask():
    x = np.zeros((5, 8, 8), np.float32)
    x[-1, :3, :3] = 1
    x[-1, 7, 7] = 1  # 10 ones: not a perfect square
    with pytest.raises(ValueError):
        callers_ref.restraints_ref(x)


# ------------------------------------------------------------------------------------------------- device kernels
@pytest.mark.gpu
def test_length_and_inpaint_masks_bit_exact(G):
    from text2protein_b200 import utils as U
    cfg = with_device(tiny_cfg(8), "cuda")
    cfg.data.min_res_num, cfg.data.max_res_num = 5, 24
    cfg.model.condition = ["length", "ss", "inpainting"]
    assert torch.equal(U.get_mask_all_lengths(cfg, batch_size=3).cpu(), torch.from_numpy(G["all_lengths_5_24_b3"]))
    for i in range(3):
        info = str(G[f"selected_{i}_info"])
        batch = U.selected_mask_batch({"coords_6d": torch.zeros(2, 8, 24, 24)}, info, cfg)
        assert torch.equal(batch["mask_inpaint"].cpu(), torch.from_numpy(G[f"selected_{i}"]))
    # ragged per-sample lengths incl. the extremes 0 and N, and per-sample ranges
    lens = [0, 1, 17, 24, 24, 5]
    assert torch.equal(U.length_mask(lens, 24).cpu(), callers_ref.length_mask_ref(lens, 24))
    per = torch.tensor([[[0, 0], [5, 9]], [[23, 23], [2, 1]], [[3, 20], [3, 20]]])
    got = U.inpaint_mask(per, 3, 24).cpu()
    for b in range(3):
        info = ",".join(f"{a}:{e}" for a, e in per[b].tolist() if e >= a)
        assert torch.equal(got[b], callers_ref.selected_mask_ref(1, 24, info)[0])
    assert not U.inpaint_mask([], 2, 24).any()  # no ranges -> nothing is free


@pytest.mark.gpu
@pytest.mark.parametrize("kinds", [["length"], ["length", "ss"], ["length", "ss", "inpainting"], ["inpainting"], []])
def test_conditional_mask_bit_exact(kinds):
    from text2protein_b200 import utils as U
    B, C, N = 5, 8, 32
    lens = [4, 32, 17, 9, 25]
    info = "2:6,11,20:31"
    cond = {}
    if "length" in kinds:
        cond["length"] = callers_ref.length_mask_ref(lens, N)
    if "ss" in kinds:
        cond["ss"] = torch.zeros(B, 3, N, N)
    if "inpainting" in kinds:
        cond["inpainting"] = {"mask_inpaint": callers_ref.selected_mask_ref(B, N, info)}
    ref = callers_ref.conditional_mask_ref((B, C, N, N), cond)
    got = U.conditional_mask(lens if "length" in kinds else None,
                             U.parse_mask_info(info) if "inpainting" in kinds else None, "ss" in kinds, B, C, N)
    assert torch.equal(got.cpu(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_embed_tokens_and_token_context(dtype):
    from tests.gpu_util import make_native
    from text2protein_b200 import utils as U
    from text2protein_b200.score_sde_pytorch.models.ncsnpp import TokenContext
    g = torch.Generator().manual_seed(3)
    cfg, model, sd = make_native(tiny_cfg(5), "fp32")
    V, D = 97, cfg.model.context_dim
    table = (torch.randn(V, D, generator=g) * 0.5).to(dtype).cuda()
    tokens = torch.randint(0, V, (3, 11), generator=g)
    emb = U.embed_tokens(table, tokens.cuda())
    assert torch.equal(emb.cpu(), table.float().cpu()[tokens])  # exactly nn.Embedding's gather
    x = torch.randn(3, 5, 32, 32, generator=g).cuda()
    labels = torch.tensor([0, 2, 3]).cuda()
    a = model(x, labels, emb)
    b = model(x, labels, TokenContext(table, tokens.cuda()))
    assert torch.equal(a, b)  # gather-on-device context == materialised embedding tensor


@pytest.mark.gpu
def test_restraints_bit_exact(G):
    from text2protein_b200.postprocess import restraints_from_samples
    for k in range(3):
        x = torch.from_numpy(G[f"rosetta_{k}_in"])[None].cuda()
        out = restraints_from_samples(x)[0]
        assert out["L"] == int(G[f"rosetta_{k}_L"])
        for name in ("dist", "omega", "theta", "phi", "dist_abs", "omega_abs", "theta_abs", "phi_abs"):
            assert np.array_equal(out[name], G[f"rosetta_{k}_{name}"]), (k, name)
    # a batch of N = 128 maps with ragged lengths against the oracle
    g = torch.Generator().manual_seed(5)
    B, C, N = 6, 8, 128
    x = torch.rand(B, C, N, N, generator=g) * 3 - 1.5
    lens = [40, 128, 77, 1, 64, 101]
    for b, l in enumerate(lens):
        pad = torch.zeros(N, N)
        pad[:l, :l] = 1
        x[b, -1] = pad + (torch.rand(N, N, generator=g) - 0.5) * 0.9
    outs = restraints_from_samples(x.cuda())
    for b in range(B):
        L, npz = callers_ref.restraints_ref(x[b].numpy())
        assert outs[b]["L"] == L == lens[b]
        for name, v in npz.items():
            assert np.array_equal(outs[b][name], v), (b, name)
    bad = x[:1].clone()
    bad[0, -1, 127, 127] = 1.0 if lens[0] < 128 else 0.0
    with pytest.raises(ValueError):
        restraints_from_samples(bad.cuda())


@pytest.mark.gpu
def test_sampling_6d_driver_roundtrip(tmp_path):
    """config + DataParallel/EMA checkpoint + tokens + embedding table -> sampled_<id>.pkl in the reference layout,
    and the file content equals the sampler called directly with the same seed."""
    import yaml
    from text2protein_b200 import sampling_6d
    from text2protein_b200.score_sde_pytorch import sampling, sde_lib
    from text2protein_b200.score_sde_pytorch.models.ema import ExponentialMovingAverage
    from text2protein_b200.score_sde_pytorch.models.ncsnpp import TokenContext
    from text2protein_b200.score_sde_pytorch.utils import get_model, save_checkpoint
    cfg = with_device(tiny_cfg(5), "cuda")
    cfg.data.min_res_num = 8
    cfg.model.condition = ["length"]
    cfg.model.compute_dtype = "fp32"
    cfg_path = tmp_path / "tiny.yml"
    cfg_path.write_text(yaml.safe_dump(eval(repr(dict_from(cfg)))))
    torch.manual_seed(0)
    model = get_model(cfg)
    ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    run = tmp_path / "training" / "tiny" / "run0" / "checkpoints"
    run.mkdir(parents=True)
    opt = sampling_6d._SamplingOptimizer()
    save_checkpoint(str(run / "best.pth"), dict(optimizer=opt, model=model, ema=ema, step=7))
    g = torch.Generator().manual_seed(1)
    table = torch.randn(50, cfg.model.context_dim, generator=g) * 0.3
    tokens = {"1abc": torch.randint(0, 50, (9,), generator=g), "2xyz": torch.randint(0, 50, (9,), generator=g)}
    torch.save(tokens, tmp_path / "tok.pt")
    torch.save(table, tmp_path / "table.pt")
    written = sampling_6d.main([str(cfg_path), str(run / "best.pth"), "--batch_size", "2", "--tag", "t", "--select_length",
                                "True", "--length_index", "5", "--tokens", str(tmp_path / "tok.pt"), "--embed_table",
                                str(tmp_path / "table.pt"), "--out_root", str(tmp_path), "--num_iters", "2", "--seed", "11"])
    assert [p.name for p in written] == ["sampled_1abc.pkl", "sampled_2xyz.pkl"]
    assert written[0].parent == tmp_path / "sampling" / "coords_6d" / "tiny" / "run0" / "t"
    got = torch.cat([pkl.load(open(p, "rb")) for p in written])
    assert got.shape == (2, 5, 32, 32) and got.dtype == torch.float32
    # the same run through the public sampler API
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    fn = sampling.get_pc_sampler(sde, (2, 5, 32, 32), sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector,
                                 snr=cfg.sampling.snr, n_steps=1, eps=1e-5, device="cuda", seed=11, num_iters=2)
    from text2protein_b200 import utils as U
    mask = U.get_mask_all_lengths(cfg, batch_size=2)[4]
    ctx = TokenContext(table.cuda(), torch.stack(list(tokens.values())).cuda())
    ref, _ = fn(model, {"length": mask}, ctx)
    assert torch.equal(got, ref.cpu())
    assert torch.equal(got[:, -1], mask.float().cpu())  # padding channel == length mask (min_res_num + 4 = 12)


@pytest.mark.gpu
def test_sampling_6d_driver_pdb_inpainting(tmp_path):
    """``--pdb X.pdb --chain A --mask_info ...`` (sampling_6d.py:144-149, utils.py:108-137): the chain's own 6D map is
    the inpainting source -- everything outside the rows / columns of the selected residues comes back bit-identical to
    it, the padding channel equals the chain's length mask, the selected region moves."""
    import yaml
    from tests.cfgs import synthetic_backbone, write_pdb
    from text2protein_b200 import sampling_6d
    from text2protein_b200.pdb_conditions import map_from_pdb
    from text2protein_b200.score_sde_pytorch.models.ema import ExponentialMovingAverage
    from text2protein_b200.score_sde_pytorch.utils import get_model, save_checkpoint
    cfg = with_device(tiny_cfg(5), "cuda")
    cfg.data.min_res_num = 8
    cfg.model.condition = ["length", "inpainting"]
    cfg.model.compute_dtype = "fp32"
    cfg_path = tmp_path / "tiny.yml"
    cfg_path.write_text(yaml.safe_dump(eval(repr(dict_from(cfg)))))
    torch.manual_seed(0)
    model = get_model(cfg)
    ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    run = tmp_path / "training" / "tiny" / "run0" / "checkpoints"
    run.mkdir(parents=True)
    save_checkpoint(str(run / "best.pth"), dict(optimizer=sampling_6d._SamplingOptimizer(), model=model, ema=ema, step=1))
    g = torch.Generator().manual_seed(2)
    torch.save(torch.randn(50, cfg.model.context_dim, generator=g) * 0.3, tmp_path / "table.pt")
    torch.save({"a": torch.randint(0, 50, (7,), generator=g), "b": torch.randint(0, 50, (7,), generator=g)}, tmp_path / "tok.pt")
    nres = 20
    write_pdb(str(tmp_path / "chain.pdb"), synthetic_backbone(nres, 4), chain="A")
    written = sampling_6d.main([str(cfg_path), str(run / "best.pth"), "--batch_size", "2", "--tag", "p", "--pdb",
                                str(tmp_path / "chain.pdb"), "--chain", "A", "--mask_info", "3:7,12", "--tokens",
                                str(tmp_path / "tok.pt"), "--embed_table", str(tmp_path / "table.pt"), "--out_root",
                                str(tmp_path), "--num_iters", "2", "--seed", "5"])
    got = torch.cat([pkl.load(open(p_, "rb")) for p_ in written])
    src, n = map_from_pdb(str(tmp_path / "chain.pdb"), "A", cfg)
    assert n == nres and got.shape == (2, 5, 32, 32)
    sel = torch.zeros(32, dtype=torch.bool)
    sel[3:8] = True
    sel[12] = True
    free = sel[:, None] | sel[None, :]
    for b in range(2):
        assert torch.equal(got[b][:, ~free], src[:, ~free])          # outside the inpainted rows / columns: the chain
        assert torch.equal(got[b, -1], src[-1])                       # padding channel == length mask of the chain
        inside = free[:nres, :nres]
        assert not torch.equal(got[b, 0, :nres, :nres][inside], src[0, :nres, :nres][inside])


def dict_from(cfg):
    return {k: (dict_from(v) if isinstance(v, dict) else v) for k, v in cfg.items()}
