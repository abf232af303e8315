"""Exponential moving average of parameters -- mirror of the reference ``score_sde_pytorch/models/ema.py``.
Sampling copies the EMA shadow list into the model by POSITION (reference :51-61), which is why
``UNetModel.parameters()`` keeps the reference's order."""
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
        for shadow, p in zip(self.shadow_params, live):
            p.data.copy_(shadow.data)

    def store(self, parameters):
        self.collected_params = [p.clone() for p in parameters]

    def restore(self, parameters):
        for saved, p in zip(self.collected_params, parameters):
            p.data.copy_(saved.data)

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

    def load_state_dict(self, state_dict):
        self.decay = state_dict['decay']
        self.num_updates = state_dict['num_updates']
        self.shadow_params = state_dict['shadow_params']
