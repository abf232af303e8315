"""PC-iteration time of any BASELINE configuration (device-resident, graph-replayed), for the tables in DESIGN.md.
    python tools/bench_config.py <config> <batch> <ctx_len> [iters] [library file, e.g. libt2p_knobs.so | -] [gn_out=0|1]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests.cfgs import synthetic_condition  # noqa: E402
from text2protein_b200 import _lib, load_config  # noqa: E402
from text2protein_b200.score_sde_pytorch import sampling, sde_lib  # noqa: E402
from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel  # noqa: E402
from text2protein_b200.synthetic import rerandomize_  # noqa: E402

opts = dict(a.split("=", 1) for a in sys.argv[6:])
if len(sys.argv) > 5 and sys.argv[5] != "-":
    _lib.use_library(sys.argv[5])
name, B, Lctx = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
K = int(sys.argv[4]) if len(sys.argv) > 4 else 6
cfg = load_config(name, device="cuda")
cfg.model.compute_dtype = "bf16"
model = UNetModel(cfg).cuda()
rerandomize_(model.named_parameters(), 42)
if "gn_out" in opts:
    model.set_epilogue_groupnorm(opts["gn_out"] != "0")
Cc, N = cfg.data.num_channels, cfg.data.max_res_num
kinds = [k for k in cfg.model.condition]
cond = synthetic_condition(cfg, B, kinds)
dev_cond = {k: ({kk: vv.cuda() for kk, vv in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in cond.items()}
g = torch.Generator().manual_seed(1)
ctx = (torch.randn(B, Lctx, cfg.model.context_dim, generator=g) * 0.02).cuda()
sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
shape = (B, Cc, N, N)
x = sampling.philox_normal(shape, 7, 0, "cuda", scale=float(sde.sigma_max))
x, cmask = sampling.apply_condition(x, dev_cond)
x = x.contiguous()
x_init, x_mean = x.clone(), torch.empty_like(x)
mask_u8 = cmask.contiguous().view(torch.uint8)
model.sync_weights()
model.set_context(ctx)
labels, G = sampling.ve_tables(sde, 1e-5, K)
L = _lib.lib()


def run(k):
    a = _lib.RunArgs()
    a.x, a.x_mean, a.mask, a.x_init = x.data_ptr(), x_mean.data_ptr(), mask_u8.data_ptr(), x_init.data_ptr()
    a.label_table, a.g_table = labels.data_ptr(), G.data_ptr()
    a.num_iters, a.n_steps, a.snr, a.probability_flow = k, 1, float(cfg.sampling.snr), 0
    a.seed, a.sample_offset, a.B, a.use_graph = 7, 0, B, 1
    _lib.check(L.t2p_pc_run(model.native_handle, C.byref(a), _lib.current_stream()))


run(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run(K)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"config": name, "B": B, "N": N, "C": Cc, "L": Lctx, "conditions": kinds, "ms_per_pc_iteration": ms,
                  "maps_per_s_at_num_scales": B / (ms * 1e-3 * cfg.model.num_scales),
                  "launches_per_forward": int(L.t2p_unet_launches_per_forward(model.native_handle)),
                  "workspace_gb": L.t2p_unet_workspace_bytes(model.native_handle) / 1e9, "opts": opts}))
