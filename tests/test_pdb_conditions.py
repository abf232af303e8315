"""``--pdb`` conditioning (text2protein_b200/pdb_conditions.py): the dense restatement of ``get_coords6d`` against
outputs of the reference's own function (tests/golden/pdb_coords6d.npz, made by make_golden_pdb.py), and the PDB reader
+ masking + padding against maps assembled by hand.  CPU only."""
import os

import numpy as np
import pytest
import torch

from tests.cfgs import synthetic_backbone, write_pdb
from text2protein_b200 import AttrDict
from text2protein_b200.pdb_conditions import coords6d, map_from_pdb, read_backbone


@pytest.mark.parametrize("k,nres,seed", [(0, 37, 1), (1, 64, 2), (2, 9, 3)])
def test_coords6d_matches_reference_function(golden_dir, k, nres, seed):
    g = np.load(os.path.join(golden_dir, "pdb_coords6d.npz"))
    xyz = synthetic_backbone(nres, seed)
    assert np.array_equal(xyz, g[f"xyz_{k}"])  # same synthetic backbone as when the golden was made
    got = np.nan_to_num(coords6d(xyz))
    ref = g[f"coords6d_{k}"]
    assert got.shape == ref.shape == (nres, nres, 4)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    assert (ref[..., 0] == 1.0).any() and (ref[..., 0] < 1.0).any()  # pairs beyond and within the 20 A cut-off


def _cfg(c=5, nmax=48, nmin=5):
    return AttrDict({"data": {"num_channels": c, "max_res_num": nmax, "min_res_num": nmin},
                     "model": {"condition": ["length", "inpainting"]}, "device": "cpu"})


def test_pdb_reader_masks_and_padding(tmp_path):
    nres = 21
    xyz = synthetic_backbone(nres, 7)
    path = tmp_path / "x.pdb"
    write_pdb(str(path), xyz, chain="A", missing=[(9, "C")])  # residue 9 lacks its C: residues 8, 9, 10 are masked
    got_xyz, mask = read_backbone(str(path), "A")
    assert got_xyz.shape == (nres, 3, 3)
    want = np.round(xyz, 3)
    want[9, 2] = 0
    np.testing.assert_allclose(got_xyz, want, atol=1e-9)  # second chain, water, O atoms, altLoc B ignored
    assert mask.tolist() == [1.0] * 8 + [0.0, 0.0, 0.0] + [1.0] * 10
    m, n = map_from_pdb(str(path), "A", _cfg())
    assert n == nres and m.shape == (5, 48, 48) and m.dtype == torch.float32
    assert torch.equal(m[:, nres:], torch.zeros(5, 48 - nres, 48)) and torch.equal(m[:, :, nres:], torch.zeros(5, 48, 48 - nres))
    pair = torch.from_numpy(mask[None, :] * mask[:, None]).float()
    assert torch.equal(m[4, :nres, :nres], pair)  # padding channel = 1 inside the chain, 0 on masked residues
    c6 = torch.from_numpy(np.nan_to_num(coords6d(got_xyz))).float().permute(2, 0, 1)
    assert torch.equal(m[:4, :nres, :nres], c6 * pair)
    assert torch.allclose(m[0, :nres, :nres], m[0, :nres, :nres].T) and torch.allclose(m[1, :nres, :nres], m[1, :nres, :nres].T)


def test_pdb_errors(tmp_path):
    xyz = synthetic_backbone(12, 3)
    path = tmp_path / "y.pdb"
    write_pdb(str(path), xyz)
    with pytest.raises(NotImplementedError):
        map_from_pdb(str(path), "A", _cfg(c=8))       # secondary-structure channels need biotite
    with pytest.raises(ValueError):
        map_from_pdb(str(path), "Z", _cfg())          # no such chain
    with pytest.raises(ValueError):
        map_from_pdb(str(path), "A", _cfg(nmax=8))    # longer than max_res_num
