"""The case-4 (Cook's membrane) evaluators of upstream's ``src/postprocess_lib.py`` that call the FEM seam,
on the CUDA path.  Two access patterns:

* ONE sample per call -- the Metropolis log-posterior ``logp_y_2d`` (src/postprocess_lib.py:78-89) evaluated
  num_mc_sam + burn times in sequence (src/postprocess_lib.py:91-103).  Each call is one launch of the library on
  mapped pinned memory (``vbfem_forward_host``, batches <= 64: no staging copies, one synchronise): tens of
  microseconds instead of a traced ``tf.map_fn`` round trip.
* LARGE batches -- the reference-sample generators of the KDE comparisons (src/postprocess_lib.py:1025-1044,
  1087-1165, 1245-1252): ``num_data * num_sam`` forward solves in one call.

Only the numerical parts are mirrored (plotting stays upstream's).  ``sampyl`` is not in this image: a
plain random-walk Metropolis (``metropolis_chain``) stands in for ``sampyl.Metropolis`` when it is missing.
"""
from __future__ import annotations

import numpy as np

from .data_generation_2sam_more_loss import MeasurementData


def _fem(theta):
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, 2)
    f, h = MeasurementData.fem_fh_fun_loop_rev(theta)
    return np.asarray(f), np.asarray(h)


def metropolis_chain(logp, start, n_samples, burn=0, thin=1, scale=1.0, rng=None):
    """Random-walk Metropolis with a Gaussian proposal (what ``sampyl.Metropolis(logp, start).sample(n, burn,
    thin)`` does): returns the chain [ceil((n_samples - burn) / thin), d]."""
    rng = np.random.default_rng() if rng is None else rng
    x = np.array(start, dtype=np.float64)
    lp = logp(x)
    out = []
    for i in range(int(n_samples)):
        prop = x + scale * rng.standard_normal(x.shape)
        lpp = logp(prop)
        if np.log(rng.uniform()) < lpp - lp:
            x, lp = prop, lpp
        if i >= burn and (i - burn) % thin == 0:
            out.append(x.copy())
    return np.asarray(out)


class PostProcess:
    @staticmethod
    def logp_y_2d(y_data, sig_e):
        """Log-posterior of theta given one observation y (src/postprocess_lib.py:78-89): Gaussian likelihood
        around the FEM displacement f(theta), standard-normal prior.  One single-sample solve per call."""
        y_data = np.asarray(y_data, dtype=np.float64)

        def logp(theta):
            theta = np.asarray(theta, dtype=np.float64).reshape(1, -1)
            f, _ = _fem(theta)
            like = -0.5 / sig_e * np.sum((y_data - f) ** 2) - np.log(2 * np.pi * sig_e)
            prior = -0.5 * np.sum(theta ** 2) - np.log(2 * np.pi)
            return like + prior

        return logp

    @staticmethod
    def zpdf_2d_example_more_loss_mcmc(z_data, y_data, sig_e, sig_eta, num_mc_sam, burn_num, thin_num, rng=None,
                                       scale=1.0):
        """Reference predictive density by MCMC (src/postprocess_lib.py:91-103): Metropolis chain of theta | y,
        one batched forward solve over the chain, z = h + eta, Gaussian KDE evaluated at z_data."""
        from scipy import stats
        rng = np.random.default_rng() if rng is None else rng
        d = z_data.shape[-1]
        logp = PostProcess.logp_y_2d(y_data, sig_e)
        try:
            import sampyl
            chain = sampyl.Metropolis(logp, {"theta": np.zeros((d,))}).sample(num_mc_sam + burn_num, burn=burn_num,
                                                                             thin=thin_num)
            theta_sam = chain["theta"]
        except ImportError:
            theta_sam = metropolis_chain(logp, np.zeros((d,)), num_mc_sam + burn_num, burn_num, thin_num, scale, rng)
        eta = np.sqrt(sig_eta) * rng.standard_normal((theta_sam.shape[0], d))
        _, h = _fem(theta_sam)
        z_sam = h + eta
        kde = stats.gaussian_kde(z_sam.T)
        return kde(z_data.T), [np.mean(np.log(z_sam)), np.std(np.log(z_sam))]

    @staticmethod
    def zpdf_2d_case4_ref(z_data, sig_eta, num_sam, theta_mean, theta_sig, rng=None):
        """KDE of z = h(theta) + eta for theta ~ N(theta_mean, diag(theta_sig)) at the points z_data
        (src/postprocess_lib.py:1025-1030)."""
        from scipy import stats
        rng = np.random.default_rng() if rng is None else rng
        theta = rng.standard_normal((num_sam, 2)) * np.sqrt(theta_sig) + theta_mean
        eta = np.sqrt(sig_eta) * rng.standard_normal((num_sam, 2))
        z_sam = _fem(theta)[1] + eta
        return stats.gaussian_kde(z_sam.T)(z_data.T)

    @staticmethod
    def zpdf_2d_case4_method1(theta_mean, theta_sig, sig_eta, mf, num_points, num_sam, rng=None):
        """Predictive density of the one-step method on a grid around the sample mean
        (src/postprocess_lib.py:1032-1044): returns [z_data, pdf], [z_mu, z_std], [x_grid, y_grid]."""
        from scipy import stats
        rng = np.random.default_rng() if rng is None else rng
        theta = rng.standard_normal((num_sam, 2)) * np.sqrt(theta_sig) + theta_mean
        eta = np.sqrt(sig_eta) * rng.standard_normal((num_sam, 2))
        z_sam = _fem(theta)[1] + eta
        kde = stats.gaussian_kde(z_sam.T)
        z_mu, z_std = z_sam.mean(axis=0), z_sam.std(axis=0)
        axes = [np.linspace(z_mu[k] - mf * z_std[k], z_mu[k] + mf * z_std[k], num_points) for k in range(2)]
        x_grid, y_grid = np.meshgrid(*axes)
        z_data = np.stack((x_grid.ravel(), y_grid.ravel()), axis=1)
        return [z_data, kde(z_data.T)], [z_mu, z_std], [x_grid, y_grid]

    @staticmethod
    def _yz_samples(theta_mean, theta_sig, sig_eta, y_data, num_sam, rng):
        """theta[b, s] = std[b] * e[s] + mean[b] with ONE draw e[s] shared by all b, eta[s] likewise
        (np.kron of the reference), z = h(theta) + eta, paired with the repeated observation y[b]."""
        nb = y_data.shape[0]
        theta = np.sqrt(theta_sig)[:, None, :] * rng.standard_normal((num_sam, 2)) + np.asarray(theta_mean)[:, None, :]
        eta = np.sqrt(sig_eta) * rng.standard_normal((num_sam, 2))
        z = _fem(theta.reshape(nb * num_sam, 2))[1] + np.tile(eta, (nb, 1))
        y = np.repeat(np.asarray(y_data, dtype=np.float64), num_sam, axis=0)
        return y, z

    @staticmethod
    def moments_2d_case4_method1(theta_mean, theta_sig, sig_eta, num_sam, rng=None):
        """Mean and variance of z per observation for the one-step method (src/postprocess_lib.py:1236-1252):
        num_data * num_sam forward solves in one call."""
        rng = np.random.default_rng() if rng is None else rng
        nb = np.asarray(theta_mean).shape[0]
        _, z = PostProcess._yz_samples(theta_mean, theta_sig, sig_eta, np.zeros((nb, 2)), num_sam, rng)
        z = z.reshape(nb, num_sam, 2)
        return z.mean(axis=1), z.var(axis=1)

    @staticmethod
    def kld_2d_example_case4_method1(theta_mean, theta_sig, sig_eta, y_data, num_sam, kde_ref, rng=None):
        """|E_q[log q(z|y) - log p_ref(z|y)]| per observation with KDE densities (bandwidth factor 1) for both
        (src/postprocess_lib.py:1127-1165).  kde_ref = [joint KDE of (y, z), marginal KDE of y]."""
        from scipy import stats
        rng = np.random.default_rng() if rng is None else rng
        nb = y_data.shape[0]
        y_sam, z_sam = PostProcess._yz_samples(theta_mean, theta_sig, sig_eta, y_data, num_sam, rng)
        yz = np.concatenate((y_sam, z_sam), axis=1)
        log_q = stats.gaussian_kde(yz.T, bw_method=1.).logpdf(yz.T) - stats.gaussian_kde(y_sam.T, bw_method=1.).logpdf(y_sam.T)
        log_ref = kde_ref[0].logpdf(yz.T) - kde_ref[1].logpdf(y_sam.T)
        return np.abs(np.mean((log_q - log_ref).reshape(nb, num_sam), axis=1))

    @staticmethod
    def kld_reference_kdes_case4(theta_mean, theta_sig, sig_eta, y_data, num_sam, rng=None):
        """The reference KDE pair of src/postprocess_lib.py:1104-1122 (joint of (y, z_ref) and marginal of y,
        bandwidth factor 1) from the FEM samples of the reference posterior N(theta_mean, theta_sig)."""
        from scipy import stats
        rng = np.random.default_rng() if rng is None else rng
        y_sam, z_sam = PostProcess._yz_samples(theta_mean, theta_sig, sig_eta, y_data, num_sam, rng)
        return [stats.gaussian_kde(np.concatenate((y_sam, z_sam), axis=1).T, bw_method=1.),
                stats.gaussian_kde(y_sam.T, bw_method=1.)]
