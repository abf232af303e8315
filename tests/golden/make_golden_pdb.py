"""Golden vectors for the ``--pdb`` conditioning path, produced by the REFERENCE's own geometry code: the pure
functions ``get_coords6d`` / ``get_dihedrals`` / ``get_angles`` of /root/reference/dataset.py are extracted with `ast`
and executed (the module itself imports biotite, absent here) on a synthetic backbone.

    python tests/golden/make_golden_pdb.py        # writes tests/golden/pdb_coords6d.npz
"""
import math
import os
import sys

import numpy as np
import scipy
import scipy.spatial
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden_callers import REF, extract_functions  # noqa: E402
from tests.cfgs import synthetic_backbone  # noqa: E402


def main():
    ns = extract_functions(os.path.join(REF, "dataset.py"), {"get_coords6d", "get_dihedrals", "get_angles"})
    ns["scipy"] = scipy
    out = {}
    for k, (nres, seed) in enumerate([(37, 1), (64, 2), (9, 3)]):
        xyz = synthetic_backbone(nres, seed)
        for name in ("get_dihedrals", "get_angles"):
            ns["get_coords6d"].__globals__[name] = ns[name]
        ns["get_coords6d"].__globals__["scipy"] = scipy
        c6 = ns["get_coords6d"](xyz.copy(), dmax=20.0, normalize=True)
        out[f"xyz_{k}"] = xyz
        out[f"coords6d_{k}"] = np.nan_to_num(c6)
    np.savez_compressed(os.path.join(HERE, "pdb_coords6d.npz"), **out)
    print("wrote pdb_coords6d.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
