"""Drop-in for the bf16 branch of ``MemoryEfficientFP16Optimizer`` (optim/fp16_optimizer_memory_efficent.py:12-144):
gradients stay in the parameters' dtype, ``multiply_grads`` is deferred into ``_multiply_factor`` (:92-94),
``clip_grad_norm`` folds the clip coefficient into it (:96-116) and ``step`` applies the product (:118-130).
Here the product lives in a device scalar and is applied INSIDE the fused Adam kernel, so the two extra
full passes over the gradients (unscale, clip) of the reference disappear and no host sync is needed."""
import torch


class MemoryEfficientBF16Optimizer:
    def __init__(self, wrapped_optimizer):
        if not getattr(wrapped_optimizer.optimizer, "supports_memory_efficient_fp16", False):
            raise ValueError("Unsupported optimizer: {}".format(wrapped_optimizer.__class__.__name__))
        self.wrapped_optimizer = wrapped_optimizer
        self._multiply_factor = 1.0
        self._grad_scale = None

    @property
    def optimizer(self):
        return self.wrapped_optimizer.optimizer

    @property
    def param_groups(self):
        return self.wrapped_optimizer.param_groups

    def backward(self, loss):
        loss.backward()

    def multiply_grads(self, c):
        self._multiply_factor *= float(c)

    def clip_grad_norm(self, max_norm, aggregate_norm_fn=None):
        """Returns the (pre-clip) gradient norm as a 0-dim device tensor (no host sync)."""
        out = self.optimizer.grad_norm_and_scale(self._multiply_factor, max_norm)
        if out is None:
            return torch.zeros(())
        norm = out[0]
        if aggregate_norm_fn is not None:
            raise NotImplementedError("sharded grad-norm aggregation (model parallel) is not used by ONE-PEACE")
        self._grad_scale = out[1:2]
        return norm

    def step(self, closure=None, groups=None):
        gs = self._grad_scale
        if gs is None and self._multiply_factor != 1.0:
            gs = torch.full((1,), self._multiply_factor, dtype=torch.float32, device=self._device())
        self.optimizer.step(closure, grad_scale=gs)
        self._multiply_factor = 1.0
        self._grad_scale = None

    def _device(self):
        return self.param_groups[0]["params"][0].device

    def zero_grad(self):
        self.wrapped_optimizer.zero_grad()
        self._multiply_factor = 1.0
        self._grad_scale = None

    def set_lr(self, lr):
        self.wrapped_optimizer.set_lr(lr)

    def get_lr(self):
        return self.wrapped_optimizer.get_lr()

    def state_dict(self):
        return self.wrapped_optimizer.state_dict()

    def load_state_dict(self, state_dict, optimizer_overrides=None):
        self.wrapped_optimizer.load_state_dict(state_dict, optimizer_overrides)
