# Modified from https://raw.githubusercontent.com/fadel/pytorch_ema/master/torch_ema/ema.py (MIT), as vendored by
# the reference (szhan227/text2protein, score_sde_pytorch/models/ema.py); partially based on
# https://github.com/tensorflow/tensorflow/blob/r1.13/tensorflow/python/training/moving_averages.py
"""Exponential moving average of parameters -- mirror of the reference ``score_sde_pytorch/models/ema.py``.
Sampling copies the EMA shadow list into the model by POSITION (reference :51-61), which is why
``UNetModel.parameters()`` keeps the reference's order.

``copy_to`` / ``restore`` write with ``Tensor.copy_`` under ``no_grad`` (the reference writes through ``.data``,
which does NOT advance the tensor version counter): the native ``UNetModel`` notices a changed parameter through
that counter and re-packs its kernel-layout weights before the next forward / sampling run."""
import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError('Decay must be between 0 and 1')
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.shadow_params = [p.clone().detach() for p in parameters if p.requires_grad]
        self.collected_params = []

    def update(self, parameters):
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        with torch.no_grad():
            live = [p for p in parameters if p.requires_grad]
            for shadow, p in zip(self.shadow_params, live):
                shadow.sub_((1.0 - decay) * (shadow - p))

    def copy_to(self, parameters):
        live = [p for p in parameters if p.requires_grad]
        with torch.no_grad():
            for shadow, p in zip(self.shadow_params, live):
                p.copy_(shadow)

    def store(self, parameters):
        self.collected_params = [p.clone() for p in parameters]

    def restore(self, parameters):
        with torch.no_grad():
            for saved, p in zip(self.collected_params, parameters):
                p.copy_(saved)

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

    def load_state_dict(self, state_dict):
        self.decay = state_dict['decay']
        self.num_updates = state_dict['num_updates']
        self.shadow_params = state_dict['shadow_params']
