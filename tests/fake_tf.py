"""A minimal stand-in for the TensorFlow entry points tf_bridge uses, so that the bridge's code is
executed end to end where TensorFlow itself is absent (this image).  It models the contracts the
bridge relies on, nothing more:

  tf.custom_gradient(f)      f(x) -> (outputs, grad_fn); the wrapper returns outputs and keeps grad_fn
                             (here: on the wrapper, attribute ``last_grad``) for the "tape" to call
  tf.py_function(fn, inp, Tout)   eager island: calls fn on the inputs, returns tensors of dtype Tout
  tf.experimental.dlpack     to_dlpack / from_dlpack capsules (exchanged with torch.utils.dlpack)
  tf.identity, tf.device, tf.float64, Tensor.shape / set_shape

Tensors are thin wrappers around torch CUDA tensors: the DLPack round trip is real."""
import contextlib
import types

import torch
from torch.utils import dlpack as tdl


class Tensor:
    def __init__(self, t):
        self._t = t
        self.shape = tuple(t.shape)
        self.dtype = "float64"

    def set_shape(self, shape):
        assert tuple(shape) == tuple(self._t.shape), (shape, self._t.shape)
        self.shape = tuple(shape)

    def numpy(self):
        return self._t.detach().cpu().numpy()


def make_module():
    tf = types.ModuleType("tensorflow")
    tf.float64 = "float64"
    tf.Tensor = Tensor
    tf.identity = lambda t: Tensor(t._t.clone())
    tf.device = lambda name: contextlib.nullcontext()
    tf.constant = lambda a, device="cuda:0": Tensor(torch.as_tensor(a, dtype=torch.float64, device=device))

    def py_function(fn, inp, Tout):
        out = fn(*inp)
        if isinstance(Tout, (list, tuple)):
            assert len(out) == len(Tout)
            return list(out)
        return out

    def custom_gradient(f):
        def wrapper(x):
            outs, grad = f(x)
            wrapper.last_grad = grad
            return outs
        wrapper.last_grad = None
        return wrapper

    tf.py_function = py_function
    tf.custom_gradient = custom_gradient
    tf.function = lambda f=None, **kw: (f if f is not None else (lambda g: g))
    dl = types.SimpleNamespace(to_dlpack=lambda t: tdl.to_dlpack(t._t),
                               from_dlpack=lambda cap: Tensor(tdl.from_dlpack(cap)))
    tf.experimental = types.SimpleNamespace(dlpack=dl)
    return tf
