"""B200-native batched Cook's-membrane FEM forward + adjoint (the hot path of
nfeng2022/Variational-Bayesian-Inference-for-Computational-Mechanics) behind
the reference's own Python interface.  Module names follow upstream's
(fem_preprocess, fem_solver, fem_postprocess, data_generation_2sam_more_loss).

The directory name contains hyphens; import it with
``importlib.import_module("variational-bayesian-inference-for-computational-mechanics_b200")``
or through the ``vbfem_b200`` alias at the repository root.
"""
from . import _lib  # noqa: F401
from . import fem_preprocess, fem_postprocess, fem_solver, data_generation_2sam_more_loss, elbo, h5io, postprocess_lib  # noqa: F401
from ._lib import VbfemError, build, load  # noqa: F401
from .fem_preprocess import PreProcessing, cook_membrane_feap  # noqa: F401
from .fem_solver import CookFemEngine, FemSolver  # noqa: F401
from .fem_postprocess import PostProcessing  # noqa: F401
from .data_generation_2sam_more_loss import MeasurementData  # noqa: F401
