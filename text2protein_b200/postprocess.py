"""The step right after the sampling loop: reference ``sampling_rosetta.py:69-96`` turns a sampled
6D map into the restraint arrays handed to Rosetta -- round the padding channel, derive L, crop by the mask, clip to
[-1, 1] and undo the dataset scaling.  Here it is one kernel over the whole batch while the samples are still on
the device; only the L x L arrays travel to the host."""
import torch

from . import _lib

NAMES = ("dist", "omega", "theta", "phi")


def restraints_from_samples(samples):
    """samples: float32 [B, C, N, N] on a CUDA device.  Returns one dict per sample with the reference's ``npz``
    keys (dist, omega, theta, phi, dist_abs, omega_abs, theta_abs, phi_abs) as float32 [L, L] numpy arrays and
    "L"; raises ValueError("Terminated due to improper masking channel...") like the reference when the rounded
    padding channel does not hold a perfect square of ones."""
    assert samples.is_cuda and samples.dtype == torch.float32 and samples.dim() == 4
    x = samples.contiguous()
    B, C, N, _ = x.shape
    out = torch.empty(B, 8, N * N, dtype=torch.float32, device=x.device)
    L = torch.empty(B, dtype=torch.int32, device=x.device)
    _lib.check(_lib.lib().t2p_postprocess_6d(_lib.ptr(x), B, C, N, _lib.ptr(out), _lib.ptr(L), _lib.current_stream()))
    Ls = L.cpu().tolist()
    res = []
    for b, l in enumerate(Ls):
        if l < 0:
            raise ValueError("Terminated due to improper masking channel...")
        block = out[b, :, : l * l].reshape(8, l, l).cpu().numpy()
        d = {"L": l}
        for i, n in enumerate(NAMES):
            d[n] = block[i]
            d[n + "_abs"] = block[4 + i]
        res.append(d)
    return res
