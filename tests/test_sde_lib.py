"""Host-side SDE mirror (text2protein_b200/score_sde_pytorch/sde_lib.py) against outputs of the reference's own
classes (tests/golden/rsde.npz, written by make_golden_fullsize.py): forward drift / diffusion, marginals, prior
log-density and -- SURVEY row a7 -- ``SDE.reverse`` -> ``RSDE.sde`` / ``RSDE.discretize`` (sde_lib.py:66-103) for VE
and VP, with and without probability flow.  Pure torch on CPU: a user-registered predictor that calls
``self.rsde.discretize`` runs exactly this code."""
import os

import numpy as np
import pytest
import torch

from tests.cfgs import analytic_score, rsde_inputs
from text2protein_b200.score_sde_pytorch import sde_lib


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "rsde.npz"))


def _sdes():
    return {"ve": sde_lib.VESDE(0.01, 100.0, 50), "vp": sde_lib.VPSDE(0.1, 20.0, 50)}


@pytest.mark.parametrize("name", ["ve", "vp"])
@pytest.mark.parametrize("pf", [False, True])
def test_rsde_matches_reference(gold, name, pf):
    x, t = rsde_inputs()
    r = _sdes()[name].reverse(analytic_score, probability_flow=pf)
    assert r.N == 50 and r.T == 1 and r.probability_flow == pf
    tag = f"{name}_pf{int(pf)}"
    drift, diffusion = r.sde(x, t)
    f, G = r.discretize(x, t)
    # same torch expressions in the same order: bit-identical, not merely close
    assert np.array_equal(drift.numpy(), gold[tag + "_drift"])
    assert np.array_equal(np.asarray(diffusion if not torch.is_tensor(diffusion) else diffusion.numpy(),
                                     dtype=np.float64), gold[tag + "_diffusion"])
    assert np.array_equal(f.numpy(), gold[tag + "_f"])
    assert np.array_equal(G.numpy(), gold[tag + "_G"])
    if pf:
        assert float(np.abs(G.numpy()).max()) == 0.0


@pytest.mark.parametrize("name", ["ve", "vp"])
def test_forward_sde_marginal_and_prior_match_reference(gold, name):
    x, t = rsde_inputs()
    sde = _sdes()[name]
    drift, diffusion = sde.sde(x, t)
    mean, std = sde.marginal_prob(x, t)
    assert np.array_equal(drift.numpy(), gold[name + "_fwd_drift"])
    assert np.array_equal(diffusion.numpy(), gold[name + "_fwd_diffusion"])
    assert np.array_equal(mean.numpy(), gold[name + "_mean"])
    assert np.array_equal(std.numpy(), gold[name + "_std"])
    assert np.array_equal(sde.prior_logp(x).numpy(), gold[name + "_prior_logp"])


def test_user_predictor_through_rsde_discretize_equals_stock_arithmetic():
    """The shape of a user predictor built on ``self.rsde`` (what sampling.register_predictor is for):
    x_mean = x - f, with (f, G) from RSDE.discretize, equals the closed form the fused kernel evaluates."""
    x, t = rsde_inputs()
    sde = _sdes()["ve"]
    f, G = sde.reverse(analytic_score).discretize(x, t)
    g = sde.discretize_G(t)
    want = x + (g[:, None, None, None] ** 2) * analytic_score(x, t)
    assert torch.allclose(x - f, want, rtol=0, atol=0)
    assert torch.equal(G, g)


def test_subvpsde_raises_like_the_reference_cannot_run():
    with pytest.raises(NotImplementedError):
        sde_lib.subVPSDE().sde(torch.zeros(1, 1, 2, 2), torch.ones(1))
