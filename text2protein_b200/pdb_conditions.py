"""``--pdb`` conditioning of the sampling driver: one chain of a PDB file -> the [C, N, N] 6D map the inpainting /
length conditions are built from.  Host-side restatement of the reference's data path for this flag --
``utils.py:108-137`` (``get_conditions_from_pdb``: isolate the chain, run it through ``ProteinDataset`` and
``PaddingCollate``) and ``dataset.py:170-239,364-450`` (backbone extraction with the rolling mask around residues that
lack N / CA / C, virtual C-beta, ``get_coords6d``) -- without biotite: the PDB ATOM records are parsed here.

Supported: 5-channel configurations (distance, omega, theta, phi, padding).  The 8-channel ones add three
secondary-structure block channels that the reference derives with biotite's P-SEA annotation
(``dataset.py get_coarse_constraints``); asking for them raises.
"""
import math

import numpy as np
import torch

# residue names biotite's filter_amino_acids accepts that occur in practice: the 20 standard ones plus the common
# non-standard / ambiguous codes the reference maps back to standard residues (dataset.py non_standard_to_standard)
AMINO_ACIDS = frozenset(
    "ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL "
    "MSE SEC PYL ASX GLX UNK CSO SEP TPO PTR HYP MLY KCX CME CSD".split())


def read_backbone(path, chain="A"):
    """N, CA, C coordinates [nres, 3, 3] (zeros where an atom is missing) and the per-residue validity mask [nres] of
    the amino-acid residues of ``chain`` in the first model, in file order (dataset.py:196-222).  A residue lacking
    one of the three atoms is masked together with its neighbours (all three atoms feed the C-beta reconstruction)."""
    residues, index = [], {}
    with open(path) as f:
        for line in f:
            rec = line[:6]
            if rec.startswith("ENDMDL"):
                break  # first model only
            if rec not in ("ATOM  ", "HETATM"):
                continue
            if line[21] != chain or line[17:20].strip() not in AMINO_ACIDS:
                continue
            key = (line[22:26], line[26], line[17:20])  # residue id, insertion code, residue name
            if key not in index:
                index[key] = len(residues)
                residues.append({})
            atoms = residues[index[key]]
            name = line[12:16].strip()
            if name in ("N", "CA", "C") and name not in atoms:  # first alternate location wins
                atoms[name] = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
    nres = len(residues)
    xyz = np.zeros((nres, 3, 3))
    mask = np.ones(nres)
    for i, atoms in enumerate(residues):
        for j, a in enumerate(("N", "CA", "C")):
            if a in atoms:
                xyz[i, j] = atoms[a]
            else:
                mask[i] = 0
                if i != 0:
                    mask[i - 1] = 0
                if i != nres - 1:
                    mask[i + 1] = 0
    return xyz, mask


def _dihedrals(a, b, c, d):
    # dataset.py:364-380
    b0 = -1.0 * (b - a)
    b1 = c - b
    b2 = d - c
    with np.errstate(divide="ignore", invalid="ignore"):
        b1 = b1 / np.linalg.norm(b1, axis=-1)[..., None]
        v = b0 - np.sum(b0 * b1, axis=-1)[..., None] * b1
        w = b2 - np.sum(b2 * b1, axis=-1)[..., None] * b1
        x = np.sum(v * w, axis=-1)
        y = np.sum(np.cross(b1, v) * w, axis=-1)
        return np.arctan2(y, x)


def _angles(a, b, c):
    # dataset.py:383-393
    with np.errstate(divide="ignore", invalid="ignore"):
        v = a - b
        v = v / np.linalg.norm(v, axis=-1)[..., None]
        w = c - b
        w = w / np.linalg.norm(w, axis=-1)[..., None]
        return np.arccos(np.sum(v * w, axis=-1))


def coords6d(xyz, dmax=20.0):
    """dataset.py:396-450 ``get_coords6d(xyz, dmax, normalize=True)`` as dense array arithmetic: [nres, nres, 4] =
    C-beta distance (capped at dmax), omega, theta, phi, each scaled to [-1, 1]; pairs farther apart than dmax (and the
    diagonal) keep the fill values the reference leaves there."""
    N, Ca, C = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    b = Ca - N
    c = C - Ca
    a = np.cross(b, c)
    Cb = -0.58273431 * a + 0.56802827 * b - 0.54067466 * c + Ca
    nres = xyz.shape[0]
    i, j = np.meshgrid(np.arange(nres), np.arange(nres), indexing="ij")
    d = np.linalg.norm(Cb[j] - Cb[i], axis=-1)
    near = (d <= dmax) & (i != j)  # cKDTree.query_ball_tree(dmax) pairs, self excluded
    dist = np.where(near, d, dmax)
    omega = np.where(near, _dihedrals(Ca[i], Cb[i], Cb[j], Ca[j]), 0.0)
    theta = np.where(near, _dihedrals(N[i], Ca[i], Cb[i], Cb[j]), 0.0)
    phi = np.where(near, _angles(Ca[i], Cb[i], Cb[j]), 0.0)
    return np.stack([dist / dmax * 2 - 1, omega / math.pi, theta / math.pi, phi / math.pi * 2 - 1], axis=-1)


def map_from_pdb(path, chain, config):
    """The padded [C, N, N] float tensor of the chain (C = 5) and its residue count: what ``ProteinDataset`` +
    ``PaddingCollate(max_res_num)`` hand to ``get_condition_from_batch`` (dataset.py:224-239, 452-500)."""
    if config.data.num_channels != 5:
        raise NotImplementedError("--pdb with secondary-structure channels needs biotite's P-SEA annotation "
                                  "(dataset.py get_coarse_constraints); use a 5-channel config or pass --coords")
    xyz, mask = read_backbone(path, chain)
    nres, nmax = xyz.shape[0], config.data.max_res_num
    if nres == 0:
        raise ValueError(f"{path}: no amino-acid residues in chain {chain!r}")
    if nres > nmax or nres < config.data.min_res_num:
        raise ValueError(f"{path}: chain {chain!r} has {nres} residues, outside [{config.data.min_res_num}, {nmax}]")
    c6 = np.nan_to_num(coords6d(xyz))
    c6 = np.concatenate([c6, np.ones((nres, nres, 1))], axis=-1)
    c6 = c6 * (mask[None, :] * mask[:, None])[..., None]
    out = torch.zeros(5, nmax, nmax, dtype=torch.float32)
    out[:, :nres, :nres] = torch.from_numpy(c6.transpose(2, 0, 1)).float()
    return out, nres
