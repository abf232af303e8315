"""Driver of the sampling path with the reference CLI (``sampling_6d.py``) and on-disk format.

    python -m text2protein_b200.sampling_6d CONFIG CHECKPOINT [--tag T] [--batch_size B] [--select_length True
        --length_index I] [--pdb X.pdb --chain A --mask_info 1:5,10:15 | --coords coords.pt]
        [--tokens tokens.pt --embed_table table.pt]

writes ``sampling/coords_6d/<config stem>/<run dir>/<tag>/sampled_<id>.pkl`` = a pickled ``torch.Tensor`` of shape
[1, C, N, N] per sample, exactly what the reference writes (sampling_6d.py:61,160-162) and what
``sampling_rosetta.py`` reads back.  What differs from the reference driver: the text encoder is not loaded here
(vicuna-7b weights and its tokenizer are an external dependency of the reference) -- captions arrive already
tokenised (``--tokens``: a dict {id: int64 [L]} or an int64 [n, L] tensor saved with torch.save) together with the
``embed_tokens`` table (``--embed_table``: [vocab, 4096] tensor), and the gather runs on the GPU.
"""
import argparse
import pickle as pkl
from pathlib import Path

import torch
import yaml

from . import utils as cond_utils
from .config import AttrDict
from .score_sde_pytorch import sampling, sde_lib
from .score_sde_pytorch.models.ema import ExponentialMovingAverage
from .score_sde_pytorch.models.ncsnpp import TokenContext
from .score_sde_pytorch.utils import get_model, restore_checkpoint


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("config", type=str)
    p.add_argument("checkpoint", type=str)
    p.add_argument("--pdb", type=str, default=None)
    p.add_argument("--chain", type=str, default="A")
    p.add_argument("--mask_info", type=str, default="1:5,10:15")
    p.add_argument("--tag", type=str, default="test")
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--n_iter", type=int, default=1)
    p.add_argument("--select_length", type=bool, default=False)
    p.add_argument("--length_index", type=int, default=1)  # index starts at 1
    # inputs the reference takes from its dataset / LLaMA stack
    p.add_argument("--tokens", type=str, default=None, help="torch-saved {id: int64[L]} dict or int64 [n, L] tensor")
    p.add_argument("--embed_table", type=str, default=None, help="torch-saved [vocab, context_dim] embedding table")
    p.add_argument("--coords", type=str, default=None, help="torch-saved [C, N, N] 6D map for --mask_info inpainting")
    p.add_argument("--out_root", type=str, default=".")
    p.add_argument("--num_iters", type=int, default=None, help="run only the first K PC iterations (smoke tests)")
    p.add_argument("--seed", type=int, default=None)
    return p


class _SamplingOptimizer:
    """Stand-in for ``losses.get_optimizer`` (training is out of scope): ``restore_checkpoint`` only needs an
    object that accepts the checkpoint's optimizer state."""

    def load_state_dict(self, sd):
        self.state = sd

    def state_dict(self):
        return getattr(self, "state", {})


def load_state(config, checkpoint, device):
    """sampling_6d.py:64-73: build the model, restore the checkpoint, copy the EMA weights in."""
    score_model = get_model(config)
    ema = ExponentialMovingAverage(score_model.parameters(), decay=config.model.ema_rate)
    state = dict(optimizer=_SamplingOptimizer(), model=score_model, ema=ema, step=0)
    state = restore_checkpoint(checkpoint, state, device)
    state["ema"].store(state["model"].parameters())
    state["ema"].copy_to(state["model"].parameters())
    return state


def make_sde(config):
    if config.training.sde == "vesde":
        return sde_lib.VESDE(sigma_min=config.model.sigma_min, sigma_max=config.model.sigma_max,
                             N=config.model.num_scales), 1e-5
    if config.training.sde == "vpsde":
        return sde_lib.VPSDE(beta_min=config.model.beta_min, beta_max=config.model.beta_max,
                             N=config.model.num_scales), 1e-3
    raise NotImplementedError(f"SDE {config.training.sde} unknown.")


def build_condition(args, config):
    """sampling_6d.py:145-152."""
    if args.select_length:
        mask = cond_utils.get_mask_all_lengths(config, batch_size=args.batch_size)[args.length_index - 1]
        return {"length": mask.to(config.device)}
    if args.coords is not None:
        coords = torch.load(args.coords).float()
        coords = coords[None].expand(args.batch_size, *coords.shape).contiguous()
        n_res = int(torch.round(coords[0, -1]).diagonal().sum().item())
        return cond_utils.get_condition_from_lengths(config, [n_res] * args.batch_size, coords_6d=coords,
                                                     mask_info=args.mask_info)
    if args.pdb is not None:
        # utils.py:108-137 get_conditions_from_pdb: the chain's own 6D map, replicated over the batch
        from .pdb_conditions import map_from_pdb
        coords, n_res = map_from_pdb(args.pdb, args.chain, config)
        coords = coords[None].expand(args.batch_size, *coords.shape).contiguous()
        return cond_utils.get_condition_from_lengths(config, [n_res] * args.batch_size, coords_6d=coords,
                                                     mask_info=args.mask_info)
    return {}


def save_samples(workdir, ids, samples):
    """sampling_6d.py:160-162: one pickled [1, C, N, N] tensor per sample."""
    paths = []
    for i, pid in enumerate(ids):
        path = workdir.joinpath(f"sampled_{pid}.pkl")
        with open(path, "wb") as f:
            pkl.dump(samples[i].unsqueeze(0), f)
        paths.append(path)
    return paths


def main(argv=None):
    args = build_parser().parse_args(argv)
    assert not (args.pdb is not None and args.select_length)
    with open(args.config, "r") as f:
        config = AttrDict(yaml.safe_load(f))
    config.device = args.device
    workdir = Path(args.out_root, "sampling", "coords_6d", Path(args.config).stem,
                   Path(args.checkpoint).parent.parent.stem, args.tag)
    workdir.mkdir(parents=True, exist_ok=True)

    state = load_state(config, args.checkpoint, args.device)
    sde, sampling_eps = make_sde(config)
    shape = (args.batch_size, config.data.num_channels, config.data.max_res_num, config.data.max_res_num)
    if args.num_iters is None and args.seed is None:
        sampling_fn = sampling.get_sampling_fn(config, sde, shape, sampling_eps)
    else:
        sampling_fn = sampling.get_pc_sampler(
            sde, shape, sampling.get_predictor(config.sampling.predictor.lower()),
            sampling.get_corrector(config.sampling.corrector.lower()), snr=config.sampling.snr,
            n_steps=config.sampling.n_steps_each, probability_flow=config.sampling.probability_flow,
            denoise=config.sampling.noise_removal, eps=sampling_eps, device=config.device, seed=args.seed,
            num_iters=args.num_iters)

    if args.tokens is None or args.embed_table is None:
        raise SystemExit("--tokens and --embed_table are required (the text encoder itself is not part of this path)")
    toks = torch.load(args.tokens)
    if isinstance(toks, dict):
        ids, rows = list(toks.keys()), torch.nn.utils.rnn.pad_sequence(
            [torch.as_tensor(v, dtype=torch.int64) for v in toks.values()], batch_first=True)
    else:
        rows = torch.as_tensor(toks, dtype=torch.int64)
        ids = [str(i) for i in range(rows.shape[0])]
    table = torch.load(args.embed_table).to(args.device)
    if table.dtype not in (torch.float32, torch.bfloat16):
        table = table.float()

    written = []
    total = (len(ids) + args.batch_size - 1) // args.batch_size
    for count in range(total):
        sl = slice(count * args.batch_size, (count + 1) * args.batch_size)
        pdb_id = ids[sl]
        if len(pdb_id) != args.batch_size:  # the reference skips the ragged last batch (sampling_6d.py:126-127)
            continue
        context = TokenContext(table, rows[sl].to(args.device))
        condition = build_condition(args, config)
        sample, n = sampling_fn(state["model"], condition=condition, context=context)
        written += save_samples(workdir, pdb_id, sample.cpu())
        print(f"[{count + 1} / {total}] save samples.")
    return written


if __name__ == "__main__":
    main()
