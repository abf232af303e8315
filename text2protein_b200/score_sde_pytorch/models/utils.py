# API surface and closed-form expressions follow score_sde_pytorch as vendored by the reference
# (szhan227/text2protein, score_sde_pytorch/models/utils.py):
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
"""Score-function adapter -- mirror of the reference ``score_sde_pytorch/models/utils.py`` (same names and
signatures): noise-level tables, the model registry, and ``get_score_fn`` which turns the continuous time ``t`` of
the sampler into the integer noise label the score network is conditioned on."""
import numpy as np
import torch

from .. import sde_lib

_MODELS = {}


def register_model(cls=None, *, name=None):
    """Decorator registering a model class under ``name`` (reference :27-43)."""

    def _register(c):
        key = c.__name__ if name is None else name
        if key in _MODELS:
            raise ValueError(f'Already registered model with name: {key}')
        _MODELS[key] = c
        return c

    return _register if cls is None else _register(cls)


def get_model(name):
    return _MODELS[name]


def get_sigmas(config):
    """SMLD noise levels, float64, DESCENDING (reference :50-60)."""
    m = config.model
    return np.exp(np.linspace(np.log(m.sigma_max), np.log(m.sigma_min), m.num_scales))


def create_model(config):
    """Reference :88-94 (unused by the live path: nothing registers a model there either)."""
    model = get_model(config.model.name)(config).to(config.device)
    return torch.nn.DataParallel(model)


def get_model_fn(model, train=False):
    """Reference :97-123.  The native score network has no training mode; ``train=True`` raises there."""

    def model_fn(x, labels, context=None):
        if train:
            model.train()
        else:
            model.eval()
        return model(x, labels, context)

    return model_fn


def ve_labels(sde, t):
    """VE noise label of continuous time t (reference :166-169): round((T - t) * (N - 1)) in float32."""
    labels = sde.T - t
    labels = labels * (sde.N - 1)
    return torch.round(labels).long()


def get_score_fn(sde, model, train=False, continuous=False):
    """Wraps the model output into a time-dependent score (reference :126-176)."""
    model_fn = get_model_fn(model, train=train)

    if isinstance(sde, (sde_lib.VPSDE, sde_lib.subVPSDE)):
        def score_fn(x, t, context=None):
            if continuous or isinstance(sde, sde_lib.subVPSDE):
                labels = t * 999
                score = model_fn(x, labels, context)
                std = sde.marginal_prob(torch.zeros_like(x), t)[1]
            else:
                labels = t * (sde.N - 1)
                score = model_fn(x, labels, context)
                std = sde.sqrt_1m_alphas_cumprod.to(labels.device)[labels.long()]
            return -score / std[:, None, None, None]

    elif isinstance(sde, sde_lib.VESDE):
        def score_fn(x, t, context=None):
            if continuous:
                labels = sde.marginal_prob(torch.zeros_like(x), t)[1]
            else:
                labels = ve_labels(sde, t)
            return model_fn(x, labels, context)

    else:
        raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    return score_fn


def to_flattened_numpy(x):
    return x.detach().cpu().numpy().reshape((-1,))


def from_flattened_numpy(x, shape):
    return torch.from_numpy(x.reshape(shape))
