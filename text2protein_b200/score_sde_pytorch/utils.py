# API surface and closed-form expressions follow score_sde_pytorch as vendored by the reference
# (szhan227/text2protein, score_sde_pytorch/utils.py):
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
"""Config-driven model constructor and checkpoint IO -- mirror of the reference ``score_sde_pytorch/utils.py``."""
import torch

from .models import ncsnpp


def get_model(config):
    """``DataParallel(UNetModel(config).to(config.device))`` (reference :4-9).

    The wrapper is kept so that state_dict keys carry the ``module.`` prefix existing checkpoints have, but it is
    pinned to ONE device: scaling is one process per GPU with the batch sharded up front, never DataParallel's
    per-forward weight broadcast."""
    score_model = ncsnpp.UNetModel(config)
    score_model = score_model.to(config.device)
    dev = torch.device(config.device)
    if dev.type == "cuda":
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        return torch.nn.DataParallel(score_model, device_ids=[idx], output_device=idx)
    return torch.nn.DataParallel(score_model)


def restore_checkpoint(ckpt_dir, state, device):
    """Reference :11-17 (``strict=False`` on the model, EMA restored as a positional list)."""
    loaded_state = torch.load(ckpt_dir, map_location=device)
    state['optimizer'].load_state_dict(loaded_state['optimizer'])
    state['model'].load_state_dict(loaded_state['model'], strict=False)
    state['ema'].load_state_dict(loaded_state['ema'])
    state['step'] = loaded_state['step']
    return state


def save_checkpoint(ckpt_dir, state):
    torch.save({'optimizer': state['optimizer'].state_dict(), 'model': state['model'].state_dict(),
                'ema': state['ema'].state_dict(), 'step': state['step']}, ckpt_dir)


def recursive_to(obj, device):
    if isinstance(obj, torch.Tensor):
        return obj.cpu() if device == 'cpu' else obj.to(device, non_blocking=True)
    if isinstance(obj, list):
        return [recursive_to(o, device) for o in obj]
    if isinstance(obj, tuple):
        return tuple(recursive_to(o, device) for o in obj)
    if isinstance(obj, dict):
        return {k: recursive_to(v, device) for k, v in obj.items()}
    return obj
