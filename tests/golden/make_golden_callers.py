"""Golden vectors for the callers either side of the loop, produced by the REFERENCE's own source:
the pure functions of /root/reference/utils.py are extracted with `ast` and executed (the module itself imports
biotite, absent here); the restraint arithmetic of sampling_rosetta.py:69-100 lives inside main(), so those
statements are sliced out of the file by line number and executed on synthetic samples.

    python tests/golden/make_golden_callers.py        # writes tests/golden/callers.npz
"""
import ast
import math
import os
import textwrap

import numpy as np
import torch

REF = os.environ.get("T2P_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def extract_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "math": math}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


class Cfg(dict):
    __getattr__ = dict.__getitem__


def main():
    ns = extract_functions(os.path.join(REF, "utils.py"), {"get_mask_all_lengths", "selected_mask_batch"})
    out = {}
    cfg = Cfg(data=Cfg(min_res_num=5, max_res_num=24), model=Cfg(condition=["length", "ss", "inpainting"]), device="cpu")
    out["all_lengths_5_24_b3"] = ns["get_mask_all_lengths"](cfg, batch_size=3).numpy()
    for i, info in enumerate(["1:5,10:15", "0,3:4,23", "7"]):
        batch = {"coords_6d": torch.zeros(2, 8, 24, 24)}
        out[f"selected_{i}"] = ns["selected_mask_batch"](batch, info, cfg)["mask_inpaint"].numpy()
        out[f"selected_{i}_info"] = np.array(info)

    # sampling_rosetta.py: the statements from `msk = np.round(coords_6d[-1])` to `npz["phi_abs"] = ...`
    lines = open(os.path.join(REF, "sampling_rosetta.py")).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "msk = np.round(coords_6d[-1])" in l)
    end = next(i for i, l in enumerate(lines) if 'npz["phi_abs"]' in l)
    body = lines[start:end + 1]
    # drop the `if args.pdb ... else ...` block about the sequence (needs pyrosetta); keep mask, L and npz arithmetic
    a = next(i for i, l in enumerate(body) if "if args.pdb is not None" in l)
    b = next(i for i, l in enumerate(body) if l.strip().startswith("npz = {}"))
    snippet = textwrap.dedent("\n".join(body[:a] + body[b:]))
    g = torch.Generator().manual_seed(99)
    for k, (N, L, C) in enumerate([(16, 11, 5), (32, 32, 8), (24, 1, 5)]):
        x = (torch.rand(C, N, N, generator=g) * 2.6 - 1.3)
        pad = torch.zeros(N, N)
        pad[:L, :L] = 1
        x[-1] = pad + (torch.rand(N, N, generator=g) - 0.5) * 0.8   # rounds back to the 0/1 mask
        env = {"np": np, "math": math, "coords_6d": x.numpy().copy()}
        exec(snippet, env)
        out[f"rosetta_{k}_in"] = x.numpy()
        out[f"rosetta_{k}_L"] = np.array(env["L"])
        for name, v in env["npz"].items():
            out[f"rosetta_{k}_{name}"] = v
    np.savez_compressed(os.path.join(HERE, "callers.npz"), **out)
    print("wrote callers.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
