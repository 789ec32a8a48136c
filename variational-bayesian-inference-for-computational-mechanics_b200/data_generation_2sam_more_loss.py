"""Forward-model adapter: the drop-in for the reference's
``MeasurementData`` FEM members (upstream
src/data_generation_2sam_more_loss.py:14-21,64-125,169-192).

``MeasurementData.fem_fh_fun_loop_rev(x[N,2]) -> [y[N,2], h[N,2]]`` is the
operator boundary that ``main_custom_training.py:191-196`` calls.  Here it is
one batched CUDA launch; it accepts

* a torch CUDA float64 tensor -> differentiable (``torch.autograd.Function``
  whose backward is the library's adjoint, the counterpart of the
  ``tf.custom_gradient`` wrapper in ``tf_bridge``),
* a torch CPU tensor or NumPy array -> the host-buffer C-ABI entry point
  (copies inside), returning the same kind of array.

Class attributes keep upstream's names and meaning.
"""
from __future__ import annotations

import math

import numpy as np

from . import fem_solver
from .fem_preprocess import PreProcessing


def _fem_function():
    import torch

    class FemFH(torch.autograd.Function):
        """y, h = FEM(x) with the discrete adjoint as backward (what
        tape.gradient derives upstream, main_custom_training.py:252-256)."""

        @staticmethod
        def forward(ctx, x, engine):
            # every autograd node owns its state: the 4x2 Jacobian per sample (64 B) is saved on ctx,
            # so several FEM calls may sit in one graph and backward in any order
            ctx.engine = engine
            ctx.n = x.shape[0]
            y, h, jac = engine.forward_jac(x.contiguous())
            ctx.save_for_backward(jac)
            return y, h

        @staticmethod
        def backward(ctx, gy, gh):
            (jac,) = ctx.saved_tensors
            z = lambda g: torch.zeros(ctx.n, 2, dtype=torch.float64, device=ctx.engine.device) if g is None \
                else g.contiguous()
            return ctx.engine.jac_vjp(jac, z(gy), z(gh)), None

    return FemFH


_FemFH = None


class MeasurementData:
    # same defaults as upstream (src/data_generation_2sam_more_loss.py:16-21);
    # main_custom_training.py:32-38 overrides theta_mean/theta_std.
    theta_mean = np.zeros((2,))
    theta_std = np.ones((2,))
    node_id = 231
    ele_id, nipt_id = 12, np.array([1, 3], dtype=int)
    num_parallel_cores = 10  # kept for signature compatibility; the batch is one launch
    device = None            # CUDA device index (None = torch's current device)

    def __init__(self, n_sam, ne_sam, d_y, d_z, d_theta, sig_e, sig_eta):
        # same attributes and initial values as upstream (src/data_generation_2sam_more_loss.py:22-39); no random
        # draws here, so a seeded generate_data_fem consumes the generator exactly like upstream's
        self.n_sam, self.ne_sam = n_sam, ne_sam
        self.d_y, self.d_theta, self.d_z = d_y, d_theta, d_z
        self.sig_e, self.sig_eta = sig_e, sig_eta
        self.e_data = np.zeros((n_sam, d_theta))
        self.y_data = np.zeros((n_sam, d_y))
        self.y_scaled_data = np.zeros((n_sam, d_y))
        self.z_data = np.zeros((n_sam, d_z))
        self.log_z_data = np.zeros((n_sam, d_z))
        self.z_scaled_data = np.zeros((n_sam, d_z))
        self.y_mean, self.y_std = np.zeros((1, d_y)), np.zeros((1, d_y))
        self.z_mean, self.z_std = np.zeros((1, d_z)), np.zeros((1, d_z))

    # ------------------------------------------------------------------ engine
    @classmethod
    def engine(cls):
        return fem_solver.default_engine(
            cls.device, theta_mean=tuple(float(v) for v in cls.theta_mean),
            theta_std=tuple(float(v) for v in cls.theta_std), node_id=int(cls.node_id), ele_id=int(cls.ele_id),
            nipt_id=tuple(int(v) for v in cls.nipt_id))

    # ---------------------------------------------------------------- boundary
    @staticmethod
    def fem_fh_fun_loop_rev(x):
        """x[N,2] -> [y[N,2], h[N,2]]  (src/data_generation_2sam_more_loss.py:169-175)."""
        global _FemFH
        import torch

        eng = MeasurementData.engine()
        if isinstance(x, torch.Tensor):
            if x.dim() != 2 or x.shape[-1] != 2:
                raise ValueError("x must have shape [N, 2]")
            if x.is_cuda:
                if _FemFH is None:
                    _FemFH = _fem_function()
                y, h = _FemFH.apply(x.to(torch.float64), eng)
                return [y, h]
            y, h = eng.forward_host(x.detach().numpy())
            return [torch.from_numpy(y), torch.from_numpy(h)]
        x = np.asarray(x, dtype=np.float64)
        if x.ndim != 2 or x.shape[-1] != 2:
            raise ValueError("x must have shape [N, 2]")
        y, h = eng.forward_host(x)
        return [y, h]

    @staticmethod
    def fem_fh_fun_one_loop(x):
        """Single sample (src/data_generation_2sam_more_loss.py:177-192)."""
        y, h = MeasurementData.fem_fh_fun_loop_rev(np.asarray(x, dtype=np.float64).reshape(1, 2))
        return [y[0], h[0]]

    @classmethod
    def fem_f_fun(cls, x):
        """Displacement of the observed node (src/data_generation_2sam_more_loss.py:112-125)."""
        return cls.fem_fh_fun_one_loop(x)[0]

    @classmethod
    def fem_h_fun(cls, x):
        """von Mises measure at the observed Gauss points (src/data_generation_2sam_more_loss.py:98-110)."""
        return cls.fem_fh_fun_one_loop(x)[1]

    # ------------------------------------------------------------ data generation
    def generate_data_fem(self, rng=None):
        """Synthetic observations (src/data_generation_2sam_more_loss.py:64-96).  The random draws come from
        NumPy's global generator in upstream's order -- theta, measurement noise, prediction noise, e_data -- so
        ``np.random.seed(k)`` reproduces upstream's draws; ``rng`` (a ``RandomState`` / ``Generator``) overrides."""
        randn = np.random.randn if rng is None else (lambda *s: rng.standard_normal(s))
        theta = randn(self.n_sam, self.d_theta)
        err = math.sqrt(self.sig_e) * randn(self.n_sam, self.d_y)
        eta = math.sqrt(self.sig_eta) * randn(self.n_sam, self.d_z)
        self.e_data = randn(self.ne_sam, self.d_theta)
        f, h = MeasurementData.fem_fh_fun_loop_rev(theta)   # one batched launch instead of tf.map_fn
        self.theta_data = theta
        self.y_data = f + err
        self.y_mean = np.mean(self.y_data, axis=0, keepdims=True)
        self.y_std = np.std(self.y_data, axis=0, keepdims=True)
        self.z_data = h + eta
        self.log_z_data = np.log(self.z_data)
        self.z_mean = np.mean(self.z_data, axis=0, keepdims=True)
        self.z_std = np.std(self.z_data, axis=0, keepdims=True)

    def save_data(self, file_path, file_name):
        """The ten datasets of src/data_generation_2sam_more_loss.py:256-268, written as the same kind of
        MATLAB-7.3 / HDF5 file ``hdf5storage.write`` produces (h5io.write; upstream passes path= and
        filename= the same way)."""
        from . import h5io
        data_dic = {"y_data": self.y_data, "y_scaled_data": self.y_data, "z_data": self.z_data,
                    "log_z_data": self.log_z_data, "z_scaled_data": self.z_data, "y_mean": self.y_mean,
                    "y_std": self.y_std, "z_mean": self.z_mean, "z_std": self.z_std, "e_data": self.e_data}
        h5io.write(data=data_dic, path=file_path, filename=file_name)

    @staticmethod
    def load_data(file_path, file_name):
        """``hdf5storage.read(path=..., filename=...)`` as main_custom_training.py:76 calls it: name -> array."""
        from . import h5io
        return h5io.read(path=file_path, filename=file_name)
