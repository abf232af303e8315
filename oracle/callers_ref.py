"""TEST INFRASTRUCTURE -- CPU restatement of the reference code either side of the sampling loop.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Each function cites the reference lines it follows; `tests/golden/make_golden_callers.py` pins them against the
reference's own source (the functions / statements are extracted from /root/reference with `ast` and executed,
because the modules themselves import biotite / pyrosetta, which are not installed).
"""
import math

import numpy as np
import torch


def get_mask_all_lengths_ref(min_res_num, max_res_num, batch_size):
    """utils.py:139-148."""
    all_lengths = np.arange(min_res_num, max_res_num + 1)
    mask = torch.zeros(len(all_lengths), batch_size, max_res_num, max_res_num).bool()
    for idx, l in enumerate(all_lengths):
        mask[idx, :, :l, :l] = True
    return mask


def length_mask_ref(lengths, max_res_num):
    """utils.py:89-93 (the "length" branch of get_condition_from_batch)."""
    mask = torch.zeros(len(lengths), max_res_num, max_res_num).bool()
    for idx, l in enumerate(lengths):
        mask[idx, :l, :l] = True
    return mask


def selected_mask_ref(B, N, mask_info):
    """utils.py:66-79 (selected_mask_batch)."""
    mask = torch.zeros(B, N)
    for r in mask_info.split(","):
        if ":" in r:
            start_idx, end_idx = r.split(":")
            mask[:, int(start_idx):int(end_idx) + 1] = 1
        else:
            mask[:, int(r)] = 1
    return torch.logical_or(mask.unsqueeze(-1), mask.unsqueeze(1)).bool()


def conditional_mask_ref(shape, condition):
    """sampling.py:258-281: the bool mask the PC loop applies after every half-step (True = free)."""
    conditional_mask = torch.ones(shape).bool()
    for k, v in condition.items():
        if k == "length":
            conditional_mask = conditional_mask * v.unsqueeze(1)
            conditional_mask[:, -1] = False
        elif k == "ss":
            conditional_mask[:, 4:7] = False
        elif k == "inpainting":
            conditional_mask = conditional_mask * v["mask_inpaint"].unsqueeze(1)
    return conditional_mask


def restraints_ref(coords_6d):
    """sampling_rosetta.py:69-75,92-100.  coords_6d: float32 numpy [C, N, N] (one sample)."""
    msk = np.round(coords_6d[-1])
    L = math.sqrt(len(msk[msk == 1]))
    if not (L).is_integer():
        raise ValueError("Terminated due to improper masking channel...")
    L = int(L)
    npz = {}
    for idx, name in enumerate(["dist", "omega", "theta", "phi"]):
        npz[name] = np.clip(coords_6d[idx][msk == 1].reshape(L, L), -1, 1)
    npz["dist_abs"] = (npz["dist"] + 1) * 10
    npz["omega_abs"] = npz["omega"] * math.pi
    npz["theta_abs"] = npz["theta"] * math.pi
    npz["phi_abs"] = (npz["phi"] + 1) * math.pi / 2
    return L, npz
