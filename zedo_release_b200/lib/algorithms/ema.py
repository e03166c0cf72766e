"""``ExponentialMovingAverage`` with the reference's interface (lib/algorithms/ema.py).

The drivers only construct it and ``load_state_dict`` the checkpoint's ``'ema'`` entry
(``{'decay', 'num_updates', 'shadow_params'}``); the averaged weights are never copied into the
model (run/opt_main.py:79,135), so inference runs on the raw weights.  ``update`` / ``copy_to`` /
``store`` / ``restore`` are kept so training-side code that shares the object keeps working.
"""
import torch


def _trainable(parameters):
    return [p for p in parameters if p.requires_grad]


class ExponentialMovingAverage:
    """shadow <- shadow - (1 - d) (shadow - param), with the warm-up d = min(decay, (1 + n) / (10 + n))."""

    def __init__(self, parameters, decay, use_num_updates=True):
        if not 0.0 <= decay <= 1.0:
            raise ValueError('Decay must be between 0 and 1')
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.shadow_params = [p.detach().clone() for p in _trainable(parameters)]
        self.collected_params = []

    def _current_decay(self):
        if self.num_updates is None:
            return self.decay
        self.num_updates += 1
        return min(self.decay, (1 + self.num_updates) / (10 + self.num_updates))

    @torch.no_grad()
    def update(self, parameters):
        """Call after every optimiser step with the same parameters the object was built from."""
        weight = 1.0 - self._current_decay()
        for shadow, param in zip(self.shadow_params, _trainable(parameters)):
            shadow.sub_(weight * (shadow - param))

    @torch.no_grad()
    def copy_to(self, parameters):
        """Overwrite the trainable parameters with their moving averages.  Written through the parameter itself
        (not ``param.data``) so its version counter moves and the score modules re-pack their plan."""
        for shadow, param in zip(self.shadow_params, _trainable(parameters)):
            param.copy_(shadow)

    def store(self, parameters):
        """Remember the current parameters (to ``restore`` them after an evaluation with the EMA)."""
        self.collected_params = [p.clone() for p in parameters]

    @torch.no_grad()
    def restore(self, parameters):
        for saved, param in zip(self.collected_params, parameters):
            param.copy_(saved)

    def state_dict(self):
        return {'decay': self.decay, 'num_updates': self.num_updates, 'shadow_params': self.shadow_params}

    def load_state_dict(self, state_dict):
        self.decay, self.num_updates = state_dict['decay'], state_dict['num_updates']
        self.shadow_params = state_dict['shadow_params']
