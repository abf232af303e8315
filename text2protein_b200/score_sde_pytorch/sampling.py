# API surface and closed-form expressions follow score_sde_pytorch as vendored by the reference
# (szhan227/text2protein, score_sde_pytorch/sampling.py):
# Copyright 2020 The Google Research Authors.
#
# Licensed under the Apache License, Version 2.0 (the "License");
# you may not use this file except in compliance with the License.
# You may obtain a copy of the License at
#
#     http://www.apache.org/licenses/LICENSE-2.0
#
# Unless required by applicable law or agreed to in writing, software
# distributed under the License is distributed on an "AS IS" BASIS,
# WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND, either express or implied.
# See the License for the specific language governing permissions and
# limitations under the License.
"""Predictor-corrector sampling -- drop-in for the reference ``score_sde_pytorch/sampling.py``.

Same registries, classes and signatures (``get_sampling_fn``, ``get_pc_sampler``, ``ReverseDiffusionPredictor``,
``LangevinCorrector``, ``shared_*_update_fn``).  Two execution paths, both through the C ABI of libt2p.so:

* fast path -- native score network + VESDE + the two stock update rules: the whole loop body of
  ``pc_sampler`` (reference :279-287) runs as ``t2p_pc_run``: per iteration 2 score-network forwards and 2
  fused step kernels, captured once into a CUDA graph and replayed; nothing returns to Python inside the loop.
* generic path -- any user score model / registered predictor or corrector class: the reference loop structure
  is kept, and the stock ``update_fn`` methods do their per-element work in the fused step kernels
  (``t2p_predictor_step`` / ``t2p_corrector_step``) on a torch-produced score.

Noise comes from the in-kernel Philox4x32-10 generator (oracle/philox_ref.py documents the stream layout),
seeded from torch's default CPU generator unless a seed is given, instead of ``torch.randn_like``.
"""
import abc
import ctypes as C
import functools
import itertools

import numpy as np
import torch

from text2protein_b200 import _lib
from . import sde_lib
from .models import utils as mutils
from .models.ncsnpp import UNetModel
from .models.utils import get_score_fn

_CORRECTORS = {}
_PREDICTORS = {}


def _make_register(table):
    def register(cls=None, *, name=None):
        def _register(c):
            key = c.__name__ if name is None else name
            if key in table:
                raise ValueError(f'Already registered model with name: {key}')
            table[key] = c
            return c

        return _register if cls is None else _register(cls)

    return register


register_predictor = _make_register(_PREDICTORS)
register_predictor.__doc__ = "A decorator for registering predictor classes (reference :32-49)."
register_corrector = _make_register(_CORRECTORS)
register_corrector.__doc__ = "A decorator for registering corrector classes (reference :52-69)."


def get_predictor(name):
    return _PREDICTORS[name]


def get_corrector(name):
    return _CORRECTORS[name]


def get_sampling_fn(config, sde, shape, eps):
    """Creates the sampling function from the config (reference :78-104).  ``config.sampling.symmetrize``
    (optional, default False = the reference's behaviour) turns on the symmetrisation of the distance / omega
    maps inside the step kernels."""
    extra = {}
    if getattr(config.sampling, "symmetrize", False):
        extra["symmetrize"] = True
    return get_pc_sampler(sde=sde,
                          shape=shape,
                          predictor=get_predictor(config.sampling.predictor.lower()),
                          corrector=get_corrector(config.sampling.corrector.lower()),
                          snr=config.sampling.snr,
                          n_steps=config.sampling.n_steps_each,
                          probability_flow=config.sampling.probability_flow,
                          denoise=config.sampling.noise_removal,
                          eps=eps,
                          device=config.device,
                          **extra)


# ---------------------------------------------------------------------------------------------- noise streams
class _Noise:
    """Seed / stream bookkeeping for update_fn calls made outside ``pc_sampler``."""
    seed = None
    counter = itertools.count(1 << 40)  # far from the stream ids a sampling run uses


class _Run:
    """State of the generic-path sampling run in progress: the stock update rules hand the condition mask (and the
    symmetrisation switch) to the step kernel, whose result is then already conditioned -- the loop's own
    ``torch.where`` (reference :283-287) becomes the idempotent re-application it is in the reference."""
    mask_u8 = None
    x_init = None
    symmetrize = False


def fresh_seed():
    """63-bit seed drawn from torch's default CPU generator (so torch.manual_seed makes runs repeatable)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def set_noise_seed(seed):
    _Noise.seed = int(seed)
    _Noise.counter = itertools.count(1 << 40)


def _next_stream():
    if _Noise.seed is None:
        _Noise.seed = fresh_seed()
    return _Noise.seed, next(_Noise.counter)


def philox_normal(shape, seed, stream, device, scale=1.0, sample_offset=0):
    """float32 normals of Philox stream ``stream`` for a [B, ...] tensor whose first sample has global index
    ``sample_offset`` (replaces torch.randn / randn_like)."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    per = out.numel() // shape[0]
    _lib.check(_lib.lib().t2p_philox_normal(C.c_uint64(seed), stream, sample_offset * per, out.numel(),
                                            C.c_float(scale), _lib.ptr(out), _lib.current_stream()))
    return out


# ---------------------------------------------------------------------------------------------- step kernels
def _step_args(x, score, seed, stream, sample_offset=0):
    """Arguments of one fused half-step.  Returns (args, new-state tensor, tensors to keep alive)."""
    if not x.is_cuda:
        raise _lib.NativeError("the fused PC-step kernels need CUDA tensors; there is no CPU path")
    xs = x.detach().to(torch.float32).contiguous()
    sc = score.detach()
    if sc.dtype not in (torch.float32, torch.float64):
        sc = sc.float()
    sc = sc.contiguous()
    out = torch.empty_like(xs)  # the caller's x is never mutated
    a = _lib.StepArgs()
    a.x = xs.data_ptr()
    a.x_out = out.data_ptr()
    a.score = sc.data_ptr()
    a.score_dtype = _lib.torch_dtype_code(sc.dtype)
    a.score_nhwc = 0
    a.seed = seed
    a.stream_id = stream
    a.sample_offset = sample_offset
    a.B, a.C, a.HW = xs.shape[0], xs.shape[1], xs.shape[2] * xs.shape[3]
    a.W = xs.shape[3]
    keep = [xs, sc]
    if _Run.mask_u8 is not None and _Run.mask_u8.shape == xs.shape:
        a.mask, a.x_init = _Run.mask_u8.data_ptr(), _Run.x_init.data_ptr()
    if _Run.symmetrize:
        if xs.shape[2] != xs.shape[3]:
            raise ValueError("symmetrize needs square maps")
        a.symmetrize = 1
    return a, out, keep


class Predictor(abc.ABC):
    """The abstract class for a predictor algorithm (reference :107-130)."""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__()
        self.sde = sde
        self.rsde = sde.reverse(score_fn, probability_flow)
        self.score_fn = score_fn
        self.probability_flow = probability_flow

    @abc.abstractmethod
    def update_fn(self, x, t, context=None):
        """Returns (x, x_mean): the next state and the next state without noise."""


class Corrector(abc.ABC):
    """The abstract class for a corrector algorithm (reference :133-157)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__()
        self.sde = sde
        self.score_fn = score_fn
        self.snr = snr
        self.n_steps = n_steps

    @abc.abstractmethod
    def update_fn(self, x, t, context=None):
        """Returns (x, x_mean)."""


@register_predictor(name='reverse_diffusion')
class ReverseDiffusionPredictor(Predictor):
    """x_mean = x - rev_f, x = x_mean + G z with (rev_f, G) the discretised reverse SDE (reference :157-167)."""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__(sde, score_fn, probability_flow)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE)) or isinstance(sde, sde_lib.subVPSDE):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, x, t, context=None):
        score = self.score_fn(x, t, context)
        seed, stream = _next_stream()
        with _lib.device_of(x):
            a, xs, keep = _step_args(x, score, seed, stream)
            if isinstance(self.sde, sde_lib.VESDE):
                G = self.sde.discretize_G(t).to(device=x.device, dtype=torch.float32).contiguous()
            else:
                ts = self.sde.timestep(t)
                G = torch.sqrt(self.sde.discrete_betas.to(x.device)[ts]).contiguous()
                sa = torch.sqrt(self.sde.alphas.to(x.device)[ts]).contiguous()
                a.sqrt_alpha = sa.data_ptr()
                keep.append(sa)
            keep.append(G)
            a.G = G.data_ptr()
            a.probability_flow = 1 if self.probability_flow else 0
            x_mean = torch.empty_like(xs)
            a.x_mean_out = x_mean.data_ptr()
            _lib.check(_lib.lib().t2p_predictor_step(C.byref(a), _lib.current_stream()))
        del keep
        # the reference returns float64 here (the score is float64) and rounds with .float() right
        # after; the kernel rounds once at the end of the same float64 arithmetic.
        return xs.double(), x_mean.double()


@register_corrector(name='langevin')
class LangevinCorrector(Corrector):
    """Langevin MCMC with the batch-mean signal-to-noise step size (reference :170-199)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE, sde_lib.subVPSDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, x, t, context=None):
        sde = self.sde
        alpha = None
        if isinstance(sde, (sde_lib.VPSDE, sde_lib.subVPSDE)):
            timestep = (t * (sde.N - 1) / sde.T).long()
            alpha = sde.alphas.to(t.device)[timestep].to(torch.float32).contiguous()
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, context)
            seed, stream = _next_stream()
            with _lib.device_of(x):
                a, xs, keep = _step_args(x, grad, seed, stream)
                a.snr = float(self.snr)
                if alpha is not None:
                    a.alpha = alpha.data_ptr()
                ws = torch.empty(max(1, _lib.lib().t2p_corrector_workspace_bytes(a.B, a.C * a.HW) // 8),
                                 dtype=torch.float64, device=xs.device)
                a.workspace = ws.data_ptr()
                xm = torch.empty_like(xs)
                a.x_mean_out = xm.data_ptr()
                _lib.check(_lib.lib().t2p_corrector_step(C.byref(a), _lib.current_stream()))
            del keep
            x, x_mean = xs.double(), xm.double()
        return x, x_mean


def shared_predictor_update_fn(x, t, context, sde, model, predictor, probability_flow):
    """Builds the predictor for this call and applies it (reference :201-205)."""
    score_fn = get_score_fn(sde, model, train=False)
    return predictor(sde, score_fn, probability_flow).update_fn(x, t, context)


def shared_corrector_update_fn(x, t, context, sde, model, corrector, snr, n_steps):
    """Builds the corrector for this call and applies it (reference :207-211)."""
    score_fn = get_score_fn(sde, model, train=False)
    return corrector(sde, score_fn, snr, n_steps).update_fn(x, t, context)


# ---------------------------------------------------------------------------------------------- the sampler
def apply_condition(x, condition):
    """Conditioning of the prior sample and the bool ``conditional_mask`` (True = free to evolve), with the
    reference's torch expressions in the reference's order -- dict ORDER matters (reference :260-275)."""
    conditional_mask = torch.ones_like(x).bool()
    if condition is not None:
        for k, v in condition.items():
            if k == "length":
                x = x * v.unsqueeze(1)
                conditional_mask = conditional_mask * v.unsqueeze(1)
                x[:, -1] = v
                conditional_mask[:, -1] = False
            elif k == "ss":
                x[:, 4:7] = v
                conditional_mask[:, 4:7] = False
            elif k == "inpainting":
                conditional_mask = conditional_mask * v["mask_inpaint"].unsqueeze(1)
                x = torch.where(conditional_mask, x, v["coords_6d"])
    return x, conditional_mask


def _unwrap(model):
    return model.module if isinstance(model, torch.nn.DataParallel) else model


def ve_tables(sde, eps, num_iters):
    """Per-iteration noise labels and diffusion coefficients, with the reference's float32 expressions
    (sampling.py:257,280-281; models/utils.py:166-169; sde_lib.py:237-245)."""
    timesteps = torch.linspace(sde.T, eps, sde.N)[:num_iters]
    labels = mutils.ve_labels(sde, timesteps.clone())
    G = sde.discretize_G(timesteps)
    return labels.to(torch.int64).contiguous(), G.to(torch.float32).contiguous()


def get_pc_sampler(sde, shape, predictor, corrector, snr, n_steps=1, probability_flow=False, denoise=True,
                   eps=1e-3, device='cuda', *, seed=None, num_iters=None, sample_offset=0, use_graph=True,
                   symmetrize=False, sync_step_size=None):
    """Creates a PC sampler (reference :213-291).  Keyword-only extras, all defaulting to the reference's behaviour:

    ``seed``            Philox seed (default: drawn from torch's generator per call).
    ``num_iters``       run only the first K of sde.N iterations (benchmarks and parity tests).
    ``sample_offset``   global index of this shard's first sample.  The NOISE is keyed by global sample index, so
                        every sample sees the same normals however the batch is sharded over GPUs; the samples
                        themselves still depend on the sharding through the Langevin step size, which is a mean
                        over the batch a call sees (reference :193-195) -- unless ``sync_step_size`` is given.
    ``sync_step_size``  a ``text2protein_b200.distributed.StepSizeSync`` (one per rank, same node): the step size
                        becomes the mean over the GLOBAL batch, exchanged inside the corrector kernel through
                        NVLink peer memory, and an n-way sharded run equals one reference run of the whole batch.
                        Native fast path only.
    ``symmetrize``      the step kernels replace channels 0 and 1 (Cb-Cb distance, omega) of the state by their
                        symmetric part after every half-step wherever (i, j) and (j, i) are both free.  False =
                        bit-identical to the reference, which has no such option.
    ``use_graph``       replay the iteration as a CUDA graph (fast path)."""
    predictor_update_fn = functools.partial(shared_predictor_update_fn, sde=sde, predictor=predictor,
                                            probability_flow=probability_flow)
    corrector_update_fn = functools.partial(shared_corrector_update_fn, sde=sde, corrector=corrector, snr=snr,
                                            n_steps=n_steps)
    K = sde.N if num_iters is None else int(num_iters)

    def pc_sampler(model, condition=None, context=None):
        """Returns (samples [B,C,N,N] float32 on ``device``, number of function evaluations)."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.NativeError("pc_sampler needs a CUDA device; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        with torch.no_grad(), torch.cuda.device(dev):
            run_seed = fresh_seed() if seed is None else int(seed)
            # prior: N(0, sigma_max^2) for VE / N(0, 1) for VP (sde_lib.py:136-137,229-230), Philox stream 0
            prior_scale = float(sde.sigma_max) if isinstance(sde, sde_lib.VESDE) else 1.0
            x = philox_normal(tuple(shape), run_seed, 0, dev, scale=prior_scale, sample_offset=sample_offset)
            x, conditional_mask = apply_condition(x, condition)
            x_initial = x.detach().clone()

            net = _unwrap(model)
            fast = (isinstance(net, UNetModel) and type(sde) is sde_lib.VESDE
                    and predictor is ReverseDiffusionPredictor and corrector is LangevinCorrector)
            if symmetrize and shape[2] != shape[3]:
                raise ValueError("symmetrize needs square maps")
            if sync_step_size is not None and not fast:
                raise NotImplementedError("sync_step_size is implemented for the native fast path only")
            if fast:
                labels, G = ve_tables(sde, eps, K)
                # weights written behind the version counters (p.data.copy_) are caught by a checksum that is read back
                # while the run is already queued; if it differs, the weights are pushed and the run repeated.  (With
                # a peer group every rank must make the same sequence of launches: checked up front there.)
                stale = net.sync_weights(check_data=True if sync_step_size is not None else "deferred")
                net.set_context(context)
                x = x.contiguous()
                x_mean = torch.empty_like(x)
                mask_u8 = conditional_mask.contiguous().view(torch.uint8)
                a = _lib.RunArgs()
                a.x, a.x_mean = x.data_ptr(), x_mean.data_ptr()
                a.mask, a.x_init = mask_u8.data_ptr(), x_initial.data_ptr()
                a.label_table, a.g_table = labels.data_ptr(), G.data_ptr()
                a.num_iters, a.n_steps = K, n_steps
                a.snr = float(snr)
                a.probability_flow = 1 if probability_flow else 0
                a.seed = run_seed
                a.sample_offset = sample_offset
                a.B = x.shape[0]
                a.use_graph = 1 if use_graph else 0
                a.symmetrize = 1 if symmetrize else 0
                if sync_step_size is not None:
                    a.peers = sync_step_size.handle
                _lib.check(_lib.lib().t2p_pc_run(net.native_handle, C.byref(a), _lib.current_stream()))
                if stale is not None and stale():
                    net.sync_weights(force=True)
                    net.set_context(context)
                    x.copy_(x_initial)
                    _lib.check(_lib.lib().t2p_pc_run(net.native_handle, C.byref(a), _lib.current_stream()))
                return (x_mean if denoise else x), K * (n_steps + 1)

            # generic path: reference loop structure, any model / registered update rule
            set_noise_seed(run_seed)
            timesteps = torch.linspace(sde.T, eps, sde.N, device=dev)
            x_mean = x
            _Run.mask_u8 = conditional_mask.contiguous().view(torch.uint8)
            _Run.x_init = x_initial.float().contiguous()
            _Run.symmetrize = bool(symmetrize)
            try:
                for i in range(K):
                    t = timesteps[i]
                    vec_t = torch.ones(shape[0], device=t.device) * t
                    x, x_mean = corrector_update_fn(x, vec_t, model=model, context=context)
                    x = torch.where(conditional_mask, x, x_initial).float()
                    x, x_mean = predictor_update_fn(x, vec_t, model=model, context=context)
                    x = torch.where(conditional_mask, x, x_initial).float()
            finally:
                _Run.mask_u8, _Run.x_init, _Run.symmetrize = None, None, False
            x_mean = torch.where(conditional_mask, x_mean, x_initial).float()
            return (x_mean if denoise else x), K * (n_steps + 1)

    return pc_sampler
