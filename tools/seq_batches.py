import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from text2protein_b200 import load_config
from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel
from text2protein_b200.synthetic import rerandomize_device_
dev = torch.device("cuda", 0)
cfg = load_config("cond_length", device="cuda:0"); cfg.model.compute_dtype = "bf16"
with torch.device(dev):
    m = UNetModel(cfg)
rerandomize_device_(m.named_parameters(), 42); m.sync_weights()
for b in [int(x) for x in sys.argv[1:]]:
    ms = bench._time_loop(m, cfg, b, 4, 2, 256, ["length"], dev, sample_offset=0)
    torch.cuda.synchronize()
    print("B", b, "ms", ms, flush=True)
