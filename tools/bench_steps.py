"""Times the fused PC half-step kernels alone (graph-replayed over buffer sets larger than L2), for A/B runs.

python tools/bench_steps.py [--lib path/to/libt2p.so] [--B 64 --C 5 --N 128] [--check other.so]
Prints one JSON line per kernel.  --check runs the same inputs through a second build and reports the largest
relative difference of the updated state (the mask positions must agree exactly)."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200 import _lib  # noqa: E402


def load(path):
    L = C.CDLL(path)
    for name in ("t2p_predictor_step", "t2p_corrector_step", "t2p_corrector_workspace_bytes", "t2p_last_error"):
        res, args = _lib.SIGNATURES[name]
        getattr(L, name).restype, getattr(L, name).argtypes = res, args
    return L


def check(L, rc):
    if rc != 0:
        raise RuntimeError(L.t2p_last_error().decode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(os.path.dirname(_lib.__file__), "libt2p.so"))
    ap.add_argument("--check", default=None)
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--C", type=int, default=5)
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--mask", default="length", choices=["length", "none", "ones"])
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-mean", action="store_true", help="predictor without the x_mean output (13 B/element)")
    ap.add_argument("--in-place", action="store_true", help="conditioned positions already hold x_init (as inside t2p_pc_run)")
    a = ap.parse_args()
    dev = "cuda"
    B, Cc, N = a.B, a.C, a.N
    shape = (B, Cc, N, N)
    E = Cc * N * N
    L = load(a.lib)
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(40, N + 1, (B,), generator=g)
    ar = torch.arange(N)
    lm = (ar[None, :, None] < lengths[:, None, None]) & (ar[None, None, :] < lengths[:, None, None])
    mask = torch.ones(shape, dtype=torch.bool) * lm[:, None]
    mask[:, -1] = False
    if a.mask == "ones":
        mask[:] = True
    mask_u8 = mask.to(dev).contiguous().view(torch.uint8)
    free_frac = mask.float().mean().item()
    sets = max(2, int(8 * 64 * 5 * 128 * 128 / (B * E)))
    xs = [torch.randn(shape, device=dev) for _ in range(sets)]
    sc = [torch.randn(shape, device=dev) for _ in range(sets)]
    xi = [torch.randn(shape, device=dev) for _ in range(sets)]
    xm = [torch.empty(shape, device=dev) for _ in range(sets)]
    G = torch.full((B,), 0.3, device=dev)

    def make_args(Lx, i, pred, x=None, xmean=None):
        ws = torch.empty(max(1, Lx.t2p_corrector_workspace_bytes(B, E) // 8), dtype=torch.float64, device=dev)
        s = _lib.StepArgs()
        s.x, s.score = (x if x is not None else xs[i]).data_ptr(), sc[i].data_ptr()
        s.score_dtype, s.score_nhwc = 0, 0
        s.G, s.snr = G.data_ptr(), 0.17
        if a.mask != "none":
            s.mask, s.x_init = mask_u8.data_ptr(), xi[i].data_ptr()
            s.conditioned_in_place = 1 if a.in_place else 0
        s.x_mean_out = (xmean if xmean is not None else xm[i]).data_ptr() if (pred and not a.no_mean) else None
        s.seed, s.stream_id, s.sample_offset = 2024, 5, 0
        s.B, s.C, s.HW = B, Cc, N * N
        s.workspace = ws.data_ptr()
        s._keep = ws
        return s

    st = _lib.current_stream
    if a.check:
        L2 = load(a.check)
        for name, pred in (("predictor", True), ("corrector", False)):
            outs = []
            for Lx in (L, L2):
                x = xs[0].clone()
                xmean = torch.zeros(shape, device=dev)
                s = make_args(Lx, 0, pred, x, xmean)
                fn = Lx.t2p_predictor_step if pred else Lx.t2p_corrector_step
                check(Lx, fn(C.byref(s), st()))
                torch.cuda.synchronize()
                outs.append((x, xmean))
            d = ((outs[0][0] - outs[1][0]).abs().max() / outs[1][0].abs().max()).item()
            dm = ((outs[0][1] - outs[1][1]).abs().max() / outs[1][1].abs().max().clamp_min(1e-30)).item()
            exact = torch.equal(outs[0][0][~mask.to(dev)], outs[1][0][~mask.to(dev)]) if a.mask != "none" else True
            print(json.dumps({"check": name, "rel_diff_x": d, "rel_diff_x_mean": dm, "conditioned_exact": exact}))

    for name, fn, bytes_per in (("predictor_kernel", L.t2p_predictor_step, 17), ("corrector_kernel", L.t2p_corrector_step, 13)):
        args = [make_args(L, i, name == "predictor_kernel") for i in range(sets)]
        for s in args:
            check(L, fn(C.byref(s), st()))
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=dev)):
            for s in args:
                check(L, fn(C.byref(s), st()))
        graph.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            reps = 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / (reps * sets))
        n = B * E
        # bytes the update needs given the mask: a quad with a free position reads x, score, mask and writes x
        # (+ x_mean); a fully conditioned quad costs its mask bytes in place, else x_init in and x (+ x_mean) out;
        # the corrector reads the whole score once more for the norms (its second read hits L2)
        if a.mask == "none":
            need = n * (bytes_per - 1)
        else:
            q = mask.reshape(-1, 4).any(-1).float().mean().item()  # fraction of quads with a free position
            pred = name == "predictor_kernel"
            free_b = 17 if pred else 13
            cond_b = (1 if a.in_place else (13 if pred else 9)) + (0 if pred else 4)
            need = n * (q * free_b + (1 - q) * cond_b)
        print(json.dumps({"tag": a.tag, "in_place": a.in_place, "mask_aware_gbs": round(need / (best * 1e-3) / 1e9, 1),
                          "mask_aware_frac": round(need / (best * 1e-3) / 1e9 / 6553.0, 3), "lib": os.path.basename(a.lib), "kernel": name, "B": B, "C": Cc, "N": N,
                          "mask": a.mask, "free_frac": round(free_frac, 3), "us": round(best * 1e3, 2),
                          "nominal_gbs": round(n * bytes_per / (best * 1e-3) / 1e9, 1),
                          "frac_of_6553": round(n * bytes_per / (best * 1e-3) / 1e9 / 6553.0, 3)}))


if __name__ == "__main__":
    main()
