"""Test double for CookFemEngine.elbo_step1_partials backed by the CPU oracle
(test infrastructure: lets the sharding / all-reduce host logic run on CPU)."""
import os

import numpy as np
import torch

import fem_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    def __init__(self):
        g = np.load(os.path.join(ROOT, "tests", "golden", "ref_numpy_twin.npz"))
        mesh = {"nnodes": 231, "nele": 200, "coord": g["coord"], "conn": g["IEN"]}
        dof = {"IEN": g["IEN"], "LM": g["LM"], "free_dof": g["free_dof"], "ndof": 462, "Pf": g["Pf"]}
        self.oracle = fo.TorchOracle(mesh, dof)

    def elbo_step1_partials(self, mu, sig2, e_data, y_batch, sig_e, j_begin=0, j_end=None, want_f=False):
        B, S = mu.shape[0], e_data.shape[0]
        j_end = B * S if j_end is None else j_end
        with torch.enable_grad():  # called from inside autograd.Function.forward
            mu = mu.clone().requires_grad_(True)
            sig2 = sig2.clone().requires_grad_(True)
            theta = (e_data * sig2.sqrt().unsqueeze(1) + mu.unsqueeze(1)).reshape(-1, 2)[j_begin:j_end]
            f, _ = self.oracle.fem_fh(theta)
        fd = f.detach()
        sums = torch.stack([fd[:, 0].sum(), fd[:, 1].sum(), (fd ** 2).sum()])
        g = (B * fd - y_batch.sum(0)) / (sig_e * B * (B * S))  # d(loss)/d f_j through -term2
        if j_end > j_begin:
            with torch.enable_grad():
                gmu, gsig2 = torch.autograd.grad((f * g).sum(), [mu, sig2])
        else:
            gmu, gsig2 = torch.zeros_like(mu), torch.zeros_like(sig2)
        return sums, gmu, gsig2, (fd if want_f else None)

    def elbo_step2_partials(self, mu, sig2, e_data, j_begin=0, j_end=None, want_h=False):
        B, S = mu.shape[0], e_data.shape[0]
        j_end = B * S if j_end is None else j_end
        theta = (e_data * sig2.sqrt().unsqueeze(1) + mu.unsqueeze(1)).reshape(-1, 2)[j_begin:j_end]
        with torch.no_grad():
            _, h = self.oracle.fem_fh(theta)
        sums = torch.cat([h.sum(0), (h ** 2).sum(0)]) if j_end > j_begin else torch.zeros(4, dtype=torch.float64)
        return sums, (h if want_h else None)


class OracleEngineFused(OracleEngine):
    """Adds the test double of CookFemEngine.elbo_step1_loss (the whole step-1 loss and its gradient w.r.t. the nets'
    outputs, main_custom_training.py:183-235), from the oracle's own statement-by-statement loss and torch autograd with
    (mu, sig2, log_sig2) as independent leaves -- what vbfem_elbo_step1_loss returns."""

    def elbo_step1_loss(self, mu, sig2, log_sig2, e_data, y_batch, sig_e, j_begin, j_end, allreduce):
        import math
        B, S = mu.shape[0], e_data.shape[0]
        assert not allreduce and j_begin == 0 and j_end == B * S
        with torch.enable_grad():
            m = mu.clone().requires_grad_(True)
            s2 = sig2.clone().requires_grad_(True)
            ls = log_sig2.clone().requires_grad_(True)
            _, _, t2, t3 = fo.elbo_step1_torch(self.oracle, y_batch, m, s2, e_data, sig_e)
            t1 = -0.5 * ls.sum(dim=-1).mean(dim=0) - 0.5 * 2 * math.log(2.0 * math.pi) - 0.5 * 2
            loss = t1 - t2 - t3
            dmu, dsig2, dls = torch.autograd.grad(loss, [m, s2, ls])
        return loss.detach(), dmu, dsig2, dls

