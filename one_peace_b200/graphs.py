"""CUDA-graph execution of the hot loops (B200-first design point: "CUDA streams and graphs instead of a tracing compiler").

The encoder is ~5 kernel launches per layer through a Python / ctypes boundary.  At the benchmarked batch (64 x 197 tokens)
a layer is ~1 ms of GPU work and the launches hide behind it; at small batches (8 texts: 130 us of launches for 60 us of
work per layer) and in the training step (4.9 k launches, ~0.5 s of Python for ~0.33 s of GPU work) the CPU is the
bottleneck.  Both loops are static — same kernels, same shapes, same buffers every call — so they are captured once into
a CUDA graph and replayed with a single launch:

  * ``GraphedForward``   — ``model(**inputs)`` for fixed input shapes (the embedding API; hub_interface.py uses it when
                           constructed with ``cuda_graph=True``).  Inputs are copied into static buffers, the graph is
                           replayed, the static output is returned (cloned unless ``copy_out=False``).
  * ``GraphedTrainStep`` — criterion forward + backward of one training step (``loss, sample_size, logging = criterion(model,
                           sample); loss.backward()``) for a fixed sample shape.  Parameter gradients land in static ``.grad``
                           tensors; the optimizer step stays outside the graph (its bias correction and learning rate are host
                           scalars that change every step).  The kernel-ready parameter packs (components.PackCache) are rebuilt
                           INSIDE the graph — capture happens right after an optimizer step, when every cache is stale — so each
                           replay sees the current weights.

Capture follows the PyTorch whole-network recipe: warm-up iterations on a side stream, capture on that stream with the
allocator's private pool, replay on the caller's stream.  Nothing here falls back to eager execution silently: a sample
whose shapes differ from the captured ones raises.
"""
import torch


def _static_copy(t):
    return t.detach().clone() if torch.is_tensor(t) else t


class GraphedForward:
    def __init__(self, fn, example_inputs, warmup=3):
        """fn(**inputs) -> tensor; example_inputs: dict of CUDA tensors (shapes / dtypes are frozen)."""
        self.fn = fn
        self.static_in = {k: _static_copy(v) for k, v in example_inputs.items()}
        self.shapes = {k: (tuple(v.shape), v.dtype) for k, v in self.static_in.items() if torch.is_tensor(v)}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                fn(**self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(**self.static_in)

    def matches(self, inputs):
        return all(k in inputs and torch.is_tensor(inputs[k]) and (tuple(inputs[k].shape), inputs[k].dtype) == sd
                   for k, sd in self.shapes.items()) and len(inputs) == len(self.static_in)

    def __call__(self, copy_out=True, **inputs):
        if not self.matches(inputs):
            raise RuntimeError("GraphedForward: input shapes / dtypes differ from the captured ones")
        for k, v in inputs.items():
            if torch.is_tensor(v):
                self.static_in[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_out.clone() if copy_out else self.static_out


class GraphedTrainStep:
    """One captured criterion forward + backward.  Usage:

        step = GraphedTrainStep(model, criterion, sample, params)      # after at least one eager optimizer step
        for batch in loader:
            loss, sample_size, log = step(batch)                       # p.grad filled for every p in params
            optimizer.step()
    """

    def __init__(self, model, criterion, example_sample, params, warmup=2, optimizer_step=None):
        self.model, self.criterion, self.params = model, criterion, list(params)
        self.static_sample = self._clone_sample(example_sample)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for p in self.params:
                    p.grad = None
                loss = criterion(model, self.static_sample)[0]
                loss.backward()
                del loss
                if optimizer_step is not None:      # leaves every PackCache stale, as in a real training loop
                    optimizer_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self._invalidate_packs()
        # No autograd graph of an earlier EAGER step may be alive here (drop every reference to its loss / outputs): the
        # parameters' AccumulateGrad nodes live as long as such a graph does and are bound to the stream they were created on
        # — the legacy stream — which the engine would have to synchronise with during capture (illegal).  Once they have
        # expired, new ones are created inside the capture, on the capture stream.
        import gc
        gc.collect()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss, self.sample_size, log = criterion(model, self.static_sample)
            # torch.autograd.grad instead of loss.backward(): no AccumulateGrad nodes run inside the capture (they are bound to
            # the stream the parameters were first used on — the legacy stream after eager steps — and the engine's cross-stream
            # hand-off to it is illegal while capturing)
            grads = torch.autograd.grad(loss, self.params, allow_unused=True)
        self.static_loss = loss.detach()
        self.static_log = log
        self.static_grads = grads
        for p, g in zip(self.params, grads):
            p.grad = g

    def _invalidate_packs(self):
        """Every cached pack must be rebuilt inside the captured region so that replays track the optimizer's updates."""
        from .components import PackCache
        for m in self.model.modules():
            for v in list(vars(m).values()):
                if isinstance(v, PackCache):
                    v._key = None
                elif isinstance(v, dict):
                    for c in v.values():
                        if isinstance(c, PackCache):
                            c._key = None

    @staticmethod
    def _clone_sample(sample):
        out = {}
        for k, v in sample.items():
            if isinstance(v, dict):
                out[k] = GraphedTrainStep._clone_sample(v)
            else:
                out[k] = _static_copy(v)
        return out

    @staticmethod
    def _copy_sample(dst, src):
        for k, v in src.items():
            if isinstance(v, dict):
                GraphedTrainStep._copy_sample(dst[k], v)
            elif torch.is_tensor(v):
                if tuple(dst[k].shape) != tuple(v.shape) or dst[k].dtype != v.dtype:
                    raise RuntimeError(f"GraphedTrainStep: sample['{k}'] shape / dtype differs from the captured one")
                dst[k].copy_(v, non_blocking=True)

    def __call__(self, sample):
        self._copy_sample(self.static_sample, sample)
        self.graph.replay()
        for p, g in zip(self.params, self.static_grads):      # the optimizer may have dropped them (zero_grad(set_to_none))
            if p.grad is not g:
                p.grad = g
        return self.static_loss, self.sample_size, self.static_log
