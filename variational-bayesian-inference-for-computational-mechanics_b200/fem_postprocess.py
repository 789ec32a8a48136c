"""Post-processing entry points that read the solver's output layout
(upstream src/fem_postprocess.py:163-185).  The batched von Mises recovery at
the observation points is fused into the CUDA kernel; this module keeps the
host-side single-result function that ``fem_test.py`` calls."""
from __future__ import annotations

import numpy as np

from .fem_preprocess import PreProcessing


class PostProcessing:
    @staticmethod
    def von_mises_stress(step_id, ele_id, nipt_id):
        """sqrt(0.5 * sum((P6 sigma)^2)) at Gauss points ``nipt_id`` (1-based) of
        element ``ele_id`` -- the reference's measure, with its truncated
        deviatoric projector (src/fem_postprocess.py:163-170)."""
        s = PreProcessing.out_data["ele_stress"][:, :, ele_id - 1, step_id - 1]
        s = s[:, np.asarray(nipt_id) - 1]
        pick = [0, 4, 8, 3, 7, 2]
        P6 = PreProcessing.Pdevs[pick, :][:, pick]
        return np.sqrt(0.5 * np.sum((P6 @ s) ** 2, axis=0))
