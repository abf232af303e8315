"""Score UNet -- drop-in for the reference ``score_sde_pytorch/models/ncsnpp.py::UNetModel``.

The module is a *parameter container* with exactly the reference's state_dict names, shapes and order (so
``.pth`` checkpoints, ``load_state_dict`` and the positional EMA list of models/ema.py keep working), and a
``forward(x, time_cond, text_emb)`` that runs the native sm_100a engine (csrc/unet.cu) through the C ABI.
The parameter tree itself comes from the engine (``t2p_unet_param_info``), which derives it from the config
the way ``UNetModel.__init__`` does (reference :74-217); there is no PyTorch implementation of the forward
pass and no CPU fallback.
"""
import ctypes as C
import itertools
import math

import numpy as np
import torch
import torch.nn as nn

from text2protein_b200 import _lib
from . import utils as mutils


def _engine_cfg(config, compute_dtype):
    m = config.model
    if m.resblock_type.lower() != "biggan":
        raise NotImplementedError("only resblock_type 'biggan' is supported (every shipped config uses it)")
    if m.embedding_type.lower() != "positional":
        raise NotImplementedError("only embedding_type 'positional' is supported")
    if m.nonlinearity.lower() != "swish":
        raise NotImplementedError("only nonlinearity 'swish' is supported")
    if not m.skip_rescale:
        raise NotImplementedError("skip_rescale=False is not supported")
    c = _lib.UnetCfg()
    c.num_channels = config.data.num_channels
    c.max_res_num = config.data.max_res_num
    c.nf = m.nf
    c.n_ch_mult = len(m.ch_mult)
    for i, v in enumerate(m.ch_mult):
        c.ch_mult[i] = v
    c.num_res_blocks = m.num_res_blocks
    c.n_attn_resolutions = len(m.attn_resolutions)
    for i, v in enumerate(m.attn_resolutions):
        c.attn_resolutions[i] = v
    c.n_heads = m.n_heads          # AttributeError if absent, like the reference (ncsnpp.py:94-95)
    c.context_dim = m.context_dim
    c.num_scales = m.num_scales
    c.scale_by_sigma = 1 if m.scale_by_sigma else 0
    c.compute_dtype = compute_dtype
    return c


class _Node(nn.Module):
    """Anonymous container; children / parameters are attached by dotted name."""


def _fan_avg_uniform_(t, scale, in_axis, out_axis):
    """DDPM 'fan_avg' uniform variance scaling (reference layers.py:44-80)."""
    shape = t.shape
    rf = np.prod(shape) / shape[in_axis] / shape[out_axis]
    fan_in, fan_out = shape[in_axis] * rf, shape[out_axis] * rf
    scale = 1e-10 if scale == 0 else scale
    bound = math.sqrt(3 * scale / ((fan_in + fan_out) / 2))
    with torch.no_grad():
        t.uniform_(-bound, bound)


class TokenContext:
    """Text context given as token ids [B, L] (int64) plus the text encoder's embedding table [vocab, context_dim]
    (fp32 or bf16), both on the device: pass it wherever the reference API takes ``context``."""

    def __init__(self, table, tokens):
        assert table.is_cuda and table.dim() == 2 and table.dtype in (torch.float32, torch.bfloat16)
        self.table = table.contiguous()
        self.tokens = tokens.to(device=table.device, dtype=torch.int64).contiguous()
        assert self.tokens.dim() == 2

    def version(self):
        return (self.table._version, self.tokens._version)

    @property
    def shape(self):
        return (self.tokens.shape[0], self.tokens.shape[1], self.table.shape[1])


class UNetModel(nn.Module):
    """``UNetModel(config)``; ``config.model.compute_dtype`` (optional, default 'bf16') selects the tcgen05
    path ('bf16') or the CUDA-core verification path ('fp32')."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        dt = str(config.model.get("compute_dtype", "bf16")).lower() if hasattr(config.model, "get") else "bf16"
        self._compute_dtype = {"bf16": _lib.BF16, "bfloat16": _lib.BF16, "fp32": _lib.F32, "float32": _lib.F32}[dt]
        self.nf = config.model.nf
        self.n_heads = config.model.n_heads
        self.context_dim = config.model.context_dim
        self.register_buffer('sigmas', torch.tensor(mutils.get_sigmas(config)))
        self._handle = C.c_void_p(0)
        self._create_engine()
        if hasattr(config.model, "get") and config.model.get("fused_groupnorm", False):
            self.set_fused_groupnorm(True)
        if hasattr(config.model, "get") and config.model.get("epilogue_groupnorm", None) is not None:
            self.set_epilogue_groupnorm(bool(config.model.get("epilogue_groupnorm")))
        self._build_tree()
        self._synced_version = None
        self._synced_checksum = None
        self._ctx_key = None
        self.reset_parameters()

    # ------------------------------------------------------------------ engine lifetime
    def _create_engine(self):
        cfg = _engine_cfg(self.config, self._compute_dtype)
        h = C.c_void_p(0)
        _lib.check(_lib.lib().t2p_unet_create(C.byref(cfg), C.byref(h)))
        self._handle = h

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                _lib.lib().t2p_unet_destroy(self._handle)
                self._handle = C.c_void_p(0)
        except Exception:
            pass

    def _engine_params(self):
        L = _lib.lib()
        n = L.t2p_unet_num_params(self._handle)
        out = []
        buf = C.create_string_buffer(256)
        shape = (C.c_int64 * 4)()
        ndim, dtype = C.c_int(0), C.c_int(0)
        for i in range(n):
            _lib.check(L.t2p_unet_param_info(self._handle, i, buf, 256, shape, C.byref(ndim), C.byref(dtype)))
            out.append((buf.value.decode(), tuple(shape[j] for j in range(ndim.value)), dtype.value))
        return out

    def _build_tree(self):
        self._leaf_slots = []  # (module, parameter name) in the reference's parameter order: a flat walk for sync_weights
        for name, shape, dtype in self._engine_params():
            if name == "sigmas":
                assert dtype == _lib.F64 and tuple(self.sigmas.shape) == shape
                continue
            parts = name.split(".")
            node = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            node.register_parameter(parts[-1], nn.Parameter(torch.empty(shape, dtype=torch.float32)))
            self._leaf_slots.append((node, parts[-1]))

    def reset_parameters(self):
        """Random init in the spirit of the reference (DDPM fan_avg-uniform convs / dense, zero biases and
        init_scale-0 output layers, zeroed proj_out, default torch init for the transformer linears)."""
        init_scale = self.config.model.init_scale
        for name, p in self.named_parameters():
            parts = name.split(".")
            leaf, parent = parts[-1], parts[-2] if len(parts) > 1 else ""
            with torch.no_grad():
                if parent.startswith("GroupNorm") or parent.startswith("norm") or name.startswith("out.0."):
                    p.fill_(1.0 if leaf == "weight" else 0.0)
                elif leaf in ("bias", "b"):
                    p.zero_()
                elif leaf == "W":  # NIN [in, out]
                    _fan_avg_uniform_(p, init_scale if parent == "NIN_3" else 0.1, in_axis=0, out_axis=1)
                elif parent == "proj_out":
                    p.zero_()
                elif parent == "Conv_1" or name.startswith("out.2."):
                    _fan_avg_uniform_(p, init_scale, in_axis=1, out_axis=0)
                elif parent in ("Conv_0", "Conv_2", "Dense_0") or parts[0] in ("pre_blocks", "pre_conv"):
                    _fan_avg_uniform_(p, 1.0, in_axis=1, out_axis=0)
                else:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))

    # ------------------------------------------------------------------ weights -> engine
    def _param_tensors(self):
        # the CURRENT parameter objects through the slots recorded at construction (same order as parameters()):
        # nn.Module.parameters() walks ~350 submodules with a de-duplication set, 4-5 ms of every sampler call
        return [m._parameters[n] for m, n in self._leaf_slots]

    def _weights_version(self):
        ts = self._param_tensors() + [self.sigmas]
        return tuple(t._version for t in ts) + tuple(t.data_ptr() for t in ts)

    def _weights_checksum_device(self):
        """One multi-tensor reduction over all parameters (a few hundred microseconds): catches writers that bypass
        the version counter (``p.data.copy_(...)``, the reference's own EMA idiom).  A device scalar."""
        norms = torch._foreach_norm([p.detach() for p in self._param_tensors()])
        return torch.stack(norms).double().sum()

    def _weights_checksum(self):
        return float(self._weights_checksum_device().item())

    def sync_weights(self, force=False, check_data=False):
        """Pushes the current parameters (e.g. after load_state_dict / ema.copy_to) to the engine and repacks
        them into kernel layout.  Called by ``forward`` and by the sampler; a change is detected through the
        tensors' version counters and storage pointers, which every in-place torch op on the parameter advances
        (``load_state_dict``, optimizers, this package's ``ExponentialMovingAverage``).  A write through ``p.data``
        does not: the sampler therefore also compares a checksum of the values once per run (``check_data=True``: at
        once, with a host wait; ``check_data="deferred"``: returns a callable the caller invokes AFTER launching its
        work -- True means the weights had changed and the work must be redone after ``sync_weights(force=True)``);
        around bare ``forward`` calls use ``sync_weights(force=True)`` after such a write."""
        ver = self._weights_version()
        if not force and ver == self._synced_version:
            if not check_data:
                return None
            if check_data == "deferred":
                # the checksum is computed on the stream and read back asynchronously: the caller goes on launching
                # its work and asks `stale()` afterwards (a host wait here would leave the GPU idle while the host
                # prepares the run -- 6 ms of the 13 ms a sampler call costs on top of its iterations)
                if getattr(self, "_chk_pin", None) is None:
                    self._chk_pin = torch.empty(1, dtype=torch.float64).pin_memory()
                self._chk_pin.copy_(self._weights_checksum_device().reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                want, pin = self._synced_checksum, self._chk_pin

                def stale():
                    ev.synchronize()
                    return float(pin[0]) != want
                return stale
            if self._weights_checksum() == self._synced_checksum:
                return None
        L = _lib.lib()
        st = _lib.current_stream()
        keep = []
        for name, t in self.state_dict(keep_vars=True).items():
            if not t.is_cuda:
                raise _lib.NativeError("UNetModel parameters must live on a CUDA device (model.to('cuda')); "
                                       "there is no CPU path")
            want = torch.float64 if name == "sigmas" else torch.float32
            src = t.detach().to(want).contiguous()
            keep.append(src)
            shape = (C.c_int64 * max(1, src.dim()))(*src.shape)
            _lib.check(L.t2p_unet_load(self._handle, name.encode(), _lib.ptr(src), shape, src.dim(),
                                       _lib.torch_dtype_code(want), st))
        _lib.check(L.t2p_unet_finalize(self._handle, st))
        torch.cuda.current_stream().synchronize()
        self._synced_version = ver
        self._synced_checksum = self._weights_checksum()
        self._ctx_key = None
        return None

    def set_context(self, text_emb):
        """Projects K|V of every cross-attention for this text context once (hoisted out of the loop).  ``text_emb``
        is the reference's [B, L, context_dim] embedding tensor or a ``TokenContext`` (token ids + embedding table:
        the gather of ``llm.model.embed_tokens``, sampling_6d.py:134-137, then runs on the device too)."""
        if text_emb is None:
            raise TypeError("context=None is not supported: the reference model crashes on it as well "
                            "(to_k expects context_dim inputs, model/attention.py:161-175)")
        if isinstance(text_emb, TokenContext):
            tc = text_emb
            if self._ctx_key is not None and self._ctx_key[0] is tc and self._ctx_key[1] == tc.version():
                return
            if tc.table.shape[1] != self.context_dim:
                raise ValueError(f"embedding table width {tc.table.shape[1]} != context_dim {self.context_dim}")
            _lib.check(_lib.lib().t2p_unet_set_context_tokens(
                self._handle, _lib.ptr(tc.table), _lib.torch_dtype_code(tc.table.dtype), tc.table.shape[0],
                _lib.ptr(tc.tokens), tc.tokens.shape[0], tc.tokens.shape[1], _lib.current_stream()))
            self._ctx_key = (tc, tc.version())
            return
        # identity + version of the tensor OBJECT (kept alive here): a pointer-based key would be fooled by the
        # caching allocator handing the same address to a new context tensor
        if self._ctx_key is not None and self._ctx_key[0] is text_emb and self._ctx_key[1] == text_emb._version:
            return
        key = (text_emb, text_emb._version)
        ctx = text_emb.detach().to(torch.float32).contiguous()
        if ctx.dim() != 3 or ctx.shape[2] != self.context_dim:
            raise ValueError(f"context must be [B, L, {self.context_dim}], got {tuple(ctx.shape)}")
        _lib.check(_lib.lib().t2p_unet_set_context(self._handle, _lib.ptr(ctx), ctx.shape[0], ctx.shape[1],
                                                   _lib.current_stream()))
        self._ctx_key = key

    # ------------------------------------------------------------------ forward
    def train(self, mode=True):
        if mode:
            raise NotImplementedError("the native score network is inference-only (sampling path)")
        return super().train(False)

    @torch.no_grad()
    def forward(self, x, time_cond, text_emb=None):
        """x [B,C,N,N] (any float dtype, cast to fp32 like ``x.float()``, reference :229), time_cond [B] noise
        labels (integer for VE models; a float tensor -- the VP branch of get_score_fn passes t (N - 1) -- is embedded
        as a float and truncated only for the sigma lookup, reference :221-223), text_emb [B,L,context_dim].
        Returns float64 [B,C,N,N] = h / sigmas[time_cond.long()] (the reference's promoted dtype, :259-261)."""
        if not x.is_cuda:
            raise _lib.NativeError("UNetModel.forward needs CUDA tensors; there is no CPU path")
        with _lib.device_of(x):
            self.sync_weights()
            self.set_context(text_emb)
            xf = x.detach().to(torch.float32).contiguous()
            tc = time_cond.detach().to(xf.device)
            labels = tc.long().contiguous()
            timesteps = tc.to(torch.float32).contiguous() if tc.is_floating_point() else None
            B = xf.shape[0]
            if text_emb.shape[0] != B:
                raise ValueError("context batch does not match x")
            out = torch.empty(xf.shape, dtype=torch.float64, device=xf.device)
            _lib.check(_lib.lib().t2p_unet_forward_t(self._handle, _lib.ptr(xf), _lib.ptr(labels), _lib.ptr(timesteps),
                                                     _lib.ptr(out), _lib.F64, B, _lib.current_stream()))
        return out

    def set_fused_groupnorm(self, on=True):
        """Option (default off): GroupNorm + SiLU inside the 3x3 convolutions on 128-pixel-wide images instead of
        a separate pass (include/t2p.h, t2p_unet_set_fused_groupnorm).  ``config.model.fused_groupnorm`` sets it
        at construction."""
        _lib.check(_lib.lib().t2p_unet_set_fused_groupnorm(self._handle, 1 if on else 0))

    def set_epilogue_groupnorm(self, on=True):
        """Option (default on): every ResBlock's Conv_0 applies GroupNorm_1 + SiLU to its own output in its epilogue
        (include/t2p.h, t2p_unet_set_epilogue_groupnorm).  ``config.model.epilogue_groupnorm`` sets it at
        construction."""
        _lib.check(_lib.lib().t2p_unet_set_epilogue_groupnorm(self._handle, 1 if on else 0))

    # ------------------------------------------------------------------ debugging aids
    def set_debug(self, on=True):
        _lib.check(_lib.lib().t2p_unet_set_debug(self._handle, 1 if on else 0))

    def tap(self, name, max_elems=1 << 28):
        shape = (C.c_int64 * 4)()
        probe = torch.empty(max_elems if max_elems < (1 << 22) else (1 << 22), dtype=torch.float32, device="cuda")
        L = _lib.lib()
        rc = L.t2p_unet_tap(self._handle, name.encode(), _lib.ptr(probe), probe.numel(), shape, _lib.current_stream())
        if rc != 0:
            msg = L.t2p_last_error().decode()
            if "too small" not in msg:
                raise _lib.NativeError(msg)
            probe = torch.empty(max_elems, dtype=torch.float32, device="cuda")
            _lib.check(L.t2p_unet_tap(self._handle, name.encode(), _lib.ptr(probe), probe.numel(), shape,
                                      _lib.current_stream()))
        n = shape[0] * shape[1] * shape[2] * shape[3]
        return probe[:n].reshape(shape[0], shape[1], shape[2], shape[3]).clone()

    @property
    def native_handle(self):
        return self._handle
