"""Step-1 ELBO of the variational Bayesian network around the fused CUDA op
(upstream main_custom_training.py:183-235, 252-258).

loss = term1 - term2 - term3  (= -ELBO), where term2 -- the Monte-Carlo data
term -- needs one FEM solve per reparameterised sample
theta[b,s] = e[s] * sqrt(sig2[b]) + mu[b].  The library evaluates, in one
launch, reparameterisation -> FEM forward -> data-term cotangent -> FEM adjoint
-> reduction to d(loss)/d(mu, sig2) for a contiguous shard of the flattened
[B*S] sample axis.  Ranks (one per GPU) own disjoint shards; the only
collective is one all-reduce(sum) of 3 + 4B doubles per step.

Upstream's data term averages (y_b - f_j)^2 over ALL B x (B*S) pairs because
f_data is not reshaped back to [B, S, .] (main_custom_training.py:205,210-214);
that is reproduced here through the sufficient statistics sum_j f_j and
sum_j |f_j|^2.
"""
from __future__ import annotations

import math


def shard_range(total: int, rank: int, world: int):
    """Contiguous split of ``total`` samples over ``world`` ranks (first
    ``total % world`` ranks get one extra)."""
    q, r = divmod(int(total), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def term1(log_theta_sig, theta_dim=2):
    """main_custom_training.py:183-185."""
    return (-0.5 * log_theta_sig.sum(dim=-1).mean(dim=0) - 0.5 * theta_dim * math.log(2.0 * math.pi)
            - 0.5 * theta_dim)


def term3(theta_mean, theta_sig, theta_dim=2):
    """main_custom_training.py:226-229."""
    return (-0.5 * theta_dim * math.log(2.0 * math.pi)
            - 0.5 * (theta_sig + theta_mean ** 2).sum(dim=-1).mean(dim=0))


def term2_from_sums(sums, y_batch, n_samples, sig_e, y_dim=2):
    """term2 (main_custom_training.py:199-214) from sum_j f_j (sums[0:2]) and
    sum_j |f_j|^2 (sums[2]) over all ``n_samples`` = B*S samples."""
    B = y_batch.shape[0]
    ysum = y_batch.sum(dim=0)
    ysq = (y_batch ** 2).sum()
    tot = B * sums[2] - 2.0 * (sums[0] * ysum[0] + sums[1] * ysum[1]) + n_samples * ysq
    l1 = -0.5 * y_dim * math.log(2.0 * math.pi * sig_e)
    return l1 - 0.5 / sig_e * tot / (B * n_samples)


def _data_term_function():
    import torch

    class NegTerm2(torch.autograd.Function):
        """-term2(mu, sig2) with gradients from the fused FEM adjoint."""

        @staticmethod
        def forward(ctx, mu, sig2, loss_obj, y_batch):
            B, S = mu.shape[0], loss_obj.e_data.shape[0]
            lo, hi = shard_range(B * S, loss_obj.rank, loss_obj.world)
            if loss_obj.world > 1 and loss_obj.peer:
                # the exchange is part of the reduction kernel: P2P stores into the peers' mailboxes over NVLink
                buf = loss_obj.engine.elbo_step1_totals(
                    mu.detach().contiguous(), sig2.detach().contiguous(), loss_obj.e_data, y_batch.contiguous(),
                    loss_obj.sig_e, lo, hi)
            else:
                sums, gmu, gsig2, _ = loss_obj.engine.elbo_step1_partials(
                    mu.detach().contiguous(), sig2.detach().contiguous(), loss_obj.e_data, y_batch.contiguous(),
                    loss_obj.sig_e, lo, hi)
                buf = torch.cat([sums, gmu.reshape(-1), gsig2.reshape(-1)])
                if loss_obj.world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=loss_obj.group)
            ctx.save_for_backward(buf[3:3 + 2 * B].reshape(B, 2), buf[3 + 2 * B:].reshape(B, 2))
            return -term2_from_sums(buf[:3], y_batch, B * S, loss_obj.sig_e)

        @staticmethod
        def backward(ctx, g):
            gmu, gsig2 = ctx.saved_tensors
            return g * gmu, g * gsig2, None, None

    class Step1Fused(torch.autograd.Function):
        """The whole step-1 loss (term1 - term2 - term3) and its gradient w.r.t. (mu, sig2, log_sig2) from the
        library: reparameterisation, FEM, adjoint, reductions, the ranks' exchange and the closed-form KL terms
        without any framework arithmetic in between (main_custom_training.py:183-235)."""

        @staticmethod
        def forward(ctx, mu, sig2, log_sig2, loss_obj, y_batch):
            B, S = mu.shape[0], loss_obj.e_data.shape[0]
            lo, hi = shard_range(B * S, loss_obj.rank, loss_obj.world)
            loss, dmu, dsig2, dls = loss_obj.engine.elbo_step1_loss(
                mu.detach().contiguous(), sig2.detach().contiguous(), log_sig2.detach().contiguous(), loss_obj.e_data,
                y_batch.contiguous(), loss_obj.sig_e, lo, hi, loss_obj.world > 1)
            ctx.save_for_backward(dmu, dsig2, dls)
            return loss.clone()

        @staticmethod
        def backward(ctx, g):
            dmu, dsig2, dls = ctx.saved_tensors
            return g * dmu, g * dsig2, g * dls, None, None

    return NegTerm2, Step1Fused


_NegTerm2 = _Step1Fused = None


class Step1Loss:
    """vi_pred_loss_step1 (main_custom_training.py:231-235) on the fused op.

    ``engine`` needs one method, ``elbo_step1_partials`` (CookFemEngine has it);
    ``group``/``rank``/``world`` describe the torch.distributed process group
    whose ranks share the Monte-Carlo samples."""

    def __init__(self, engine, e_data, sig_e, group=None, rank=0, world=1, peer=None, fused=True):
        self.engine, self.e_data, self.sig_e = engine, e_data.contiguous(), float(sig_e)
        self.group, self.rank, self.world = group, int(rank), int(world)
        # peer mailboxes connected for this world (engine.peer_connect_group): exchange inside the reduction kernel;
        # otherwise (or with peer=False) one torch.distributed all-reduce after it
        connected = int(getattr(engine, "peer_world", 0)) == self.world and self.world > 1
        if peer and not connected:
            raise ValueError("peer exchange requested but the engine has no peer mailboxes for this world size")
        self.peer = connected if peer is None else bool(peer)
        # the KL terms and the loss value in the library too (one ranks' exchange or a single rank); fused=False keeps
        # them in the framework (the form the tests compare against)
        self.fused = bool(fused)

    def __call__(self, y_batch, theta_mean, theta_sig, log_theta_sig):
        global _NegTerm2, _Step1Fused
        if _NegTerm2 is None:
            _NegTerm2, _Step1Fused = _data_term_function()
        if self.fused and (self.world == 1 or self.peer) and y_batch.shape[0] <= 128 \
                and hasattr(self.engine, "elbo_step1_loss"):
            return _Step1Fused.apply(theta_mean, theta_sig, log_theta_sig, self, y_batch)
        neg_t2 = _NegTerm2.apply(theta_mean, theta_sig, self, y_batch)
        return term1(log_theta_sig) + neg_t2 - term3(theta_mean, theta_sig)


def make_step1_model(num_neuron=20, num_layers=3, y_dim=2, theta_dim=2, device=None, seed=0):
    """The two float64 MLPs of main_custom_training.py:130-153,176 (Dense+ReLU
    x3, linear head): returns a module mapping y[B,2] ->
    (theta_mean, theta_sig = exp(log_theta_sig), log_theta_sig)."""
    import torch
    from torch import nn

    def mlp():
        layers, d = [], y_dim
        for _ in range(num_layers):
            layers += [nn.Linear(d, num_neuron), nn.ReLU()]
            d = num_neuron
        layers.append(nn.Linear(d, theta_dim))
        return nn.Sequential(*layers)

    class Step1Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.mean_net, self.logsig_net = mlp(), mlp()
            self.parallel_nets, self._side = True, None

        def forward(self, y):
            if y.is_cuda and self.parallel_nets:
                # the two nets are independent: the log-variance net runs on a side stream (forward here, and --
                # autograd replays each op on the stream of its forward -- backward too).  In a captured training
                # step the ~80 small kernels of the nets then form two parallel branches of the graph.
                if self._side is None:
                    self._side = torch.cuda.Stream(device=y.device)
                cur = torch.cuda.current_stream(y.device)
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    ls = self.logsig_net(y)
                    sig = torch.exp(ls)
                m = self.mean_net(y)
                cur.wait_stream(self._side)
                return m, sig, ls
            m, ls = self.mean_net(y), self.logsig_net(y)
            return m, torch.exp(ls), ls

    g = torch.Generator().manual_seed(seed)
    model = Step1Model().double()
    with torch.no_grad():  # Keras default init: Glorot-uniform kernels, zero biases
        for mod in model.modules():
            if isinstance(mod, nn.Linear):
                lim = math.sqrt(6.0 / (mod.in_features + mod.out_features))
                mod.weight.copy_((torch.rand(mod.weight.shape, generator=g, dtype=torch.float64) * 2 - 1) * lim)
                mod.bias.zero_()
    return model.to(device) if device is not None else model


def make_step1_optimizer(model, lr=1e-3):
    """Adam as configured at main_custom_training.py:243."""
    import torch

    return torch.optim.Adam(model.parameters(), lr=lr, betas=(0.99, 0.999), eps=1e-10)


def make_step1_optimizer_capturable(model, lr=1e-3):
    """Same Adam (main_custom_training.py:243) with its step counter on the device, so that
    the whole training step can be captured in a CUDA graph."""
    import torch

    params = list(model.parameters())
    if params and params[0].is_cuda:
        try:   # one multi-tensor kernel for all 16 parameter tensors instead of a dozen foreach launches
            return torch.optim.Adam(params, lr=lr, betas=(0.99, 0.999), eps=1e-10, capturable=True, fused=True)
        except (RuntimeError, ValueError):
            pass
    return torch.optim.Adam(params, lr=lr, betas=(0.99, 0.999), eps=1e-10, capturable=True)


class GraphedStep1:
    """One step-1 training step (main_custom_training.py:252-258: forward, loss, tape.gradient,
    Adam) captured ONCE in a CUDA graph and replayed per batch:

        pinned host batch --H2D--> nets --> reparameterisation + FEM + adjoint (libvbfem) --> loss
        --> backward through the nets --> Adam

    The ~100 small launches of the two MLPs and of Adam otherwise cost about as much as the
    6400 finite-element solves of a step.  ``step(y_batch_host)`` copies the batch into the
    pinned staging buffer, replays the graph and returns the loss tensor (device); reading
    it (``float``) is the step's only synchronisation.  Falls back to eager execution when
    capture is not possible (``self.graphed`` tells which)."""

    def __init__(self, model, optimizer, loss_fn, batch_size, device, use_graph=True, warmup=3):
        import torch

        self.torch, self.model, self.opt, self.loss_fn = torch, model, optimizer, loss_fn
        self.pin = torch.empty(batch_size, 2, dtype=torch.float64).pin_memory()
        self.yb = torch.empty(batch_size, 2, dtype=torch.float64, device=device)
        self.loss = torch.zeros((), dtype=torch.float64, device=device)
        self.graph, self.graphed = None, False
        self._warm = warmup
        self._use_graph = use_graph
        self._device = device

    def _body(self):
        self.opt.zero_grad(set_to_none=True)
        mu, sig, ls = self.model(self.yb)
        loss = self.loss_fn(self.yb, mu, sig, ls)
        loss.backward()
        self.opt.step()
        self.loss.copy_(loss.detach())

    def _capture(self):
        torch = self.torch
        side = torch.cuda.Stream(device=self._device)
        side.wait_stream(torch.cuda.current_stream(self._device))
        with torch.cuda.stream(side):
            self.yb.copy_(self.pin, non_blocking=True)
            for _ in range(self._warm):   # also sizes the library's scratch buffers before capture
                self._body()
        torch.cuda.current_stream(self._device).wait_stream(side)
        torch.cuda.synchronize(self._device)
        try:
            g = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(g):
                self._body()
            self.graph, self.graphed = g, True
        except Exception as exc:  # capture unsupported (e.g. an uncapturable collective): run eagerly
            self.graph, self.graphed = None, False
            self.capture_error = repr(exc)
            torch.cuda.synchronize(self._device)

    def step(self, y_batch_host):
        """One training step on the batch; returns the loss tensor (device): reading it is the only synchronisation."""
        self.pin.copy_(self.torch.as_tensor(y_batch_host, dtype=self.torch.float64))
        if self._use_graph and self.graph is None and not hasattr(self, "capture_error"):
            self._capture()
        self.yb.copy_(self.pin, non_blocking=True)   # H2D on the launch stream, then the graph finds the batch on the device
        if self.graphed:
            self.graph.replay()
        else:
            self._body()
        return self.loss

    # ---- pipelined stepping: the host does not wait for a step's loss before it launches the next one
    def step_async(self, y_batch_host, depth=4):
        """Like ``step`` but without a host synchronisation per step: the batch goes through a ring of ``depth``
        pinned staging buffers (H2D on the launch stream right before the graph replay), the loss comes back through
        a ring of pinned scalars (D2H right after the replay).  Returns a ticket for ``loss_of``; a staging slot is
        reused only after the step that used it has finished on the GPU.  Every step still does its own H2D copy and
        D2H read -- the host just runs ahead of the GPU by up to ``depth`` steps, like any training loop that does not
        print the loss every step."""
        torch = self.torch
        if not hasattr(self, "_ring"):
            self._ring = [{"pin": torch.empty_like(self.pin).pin_memory(),
                           "loss": torch.zeros((), dtype=torch.float64).pin_memory(),
                           "done": torch.cuda.Event()} for _ in range(depth)]
            self._tick = 0
        if self._use_graph and self.graph is None and not hasattr(self, "capture_error"):
            self.pin.copy_(torch.as_tensor(y_batch_host, dtype=torch.float64))
            self._capture()       # warm-up steps + capture on the first call (they train on this batch, like step())
        slot = self._ring[self._tick % len(self._ring)]
        if self._tick >= len(self._ring):
            slot["done"].synchronize()   # the step that used this slot last has finished
        slot["pin"].copy_(torch.as_tensor(y_batch_host, dtype=torch.float64))
        self.yb.copy_(slot["pin"], non_blocking=True)
        if self.graphed:
            self.graph.replay()
        else:
            self._body()
        slot["loss"].copy_(self.loss, non_blocking=True)
        slot["done"].record()
        self._tick += 1
        return self._tick - 1

    def loss_of(self, ticket):
        """The loss of the step ``ticket`` (blocks until that step has finished)."""
        if self._tick - ticket > len(self._ring):
            raise ValueError("ticket too old: its staging slot has been reused")
        slot = self._ring[ticket % len(self._ring)]
        slot["done"].synchronize()
        return float(slot["loss"])


# ------------------------------------------------------------------------------------- step 2
def term4(z_mean, log_z_sig, z_dim=2):
    """main_custom_training.py:338-340."""
    return ((-0.5 * log_z_sig.sum(dim=-1) - z_mean.sum(dim=-1)).mean()
            - 0.5 * z_dim * math.log(2.0 * math.pi) - 0.5 * z_dim)


def term5_from_sums(sums, z_mean, z_sig, n_samples, sig_eta, z_dim=2):
    """term5 (main_custom_training.py:347-364) from sum_j h_j (sums[0:2]) and sum_j h_j^2
    (sums[2:4]) over all ``n_samples`` = B*S samples.  Upstream broadcasts h_data[B*S, 2] against
    z_mean_point[B, 1, 2], so l2 averages over ALL B x (B*S) pairs:
        mean_{b,j} l2 = -0.5/sig_eta * sum_k( -2 mean_j(h_jk) mean_b exp(zm_bk + zs_bk/2) + mean_j(h_jk^2) )."""
    l1 = -0.5 / sig_eta * torch_exp(2.0 * z_mean + 2.0 * z_sig).sum(dim=-1)
    hm, h2m = sums[0:2] / n_samples, sums[2:4] / n_samples
    em = torch_exp(z_mean + 0.5 * z_sig).mean(dim=0)
    l2 = -0.5 / sig_eta * (-2.0 * hm * em + h2m).sum()
    l3 = -0.5 * z_dim * math.log(2.0 * math.pi * sig_eta)
    return l1.mean() + l2 + l3


def torch_exp(t):
    import torch

    return torch.exp(t)


def add_loss(z_mean, z_sig, logz_mean_post, logz_sig_post):
    """main_custom_training.py:373-375."""
    return ((z_mean - logz_mean_post) ** 2).mean() + ((z_sig - logz_sig_post) ** 2).mean()


class Step2Loss:
    """vi_pred_loss_step2 (main_custom_training.py:381-384) on the fused forward-only op:
    (term4 - term5) * alpha + add_loss.  The theta nets are frozen in step 2
    (main_custom_training.py:305), so the FEM needs no adjoint here: the library returns the
    sufficient statistics of h over this rank's sample shard, the ranks all-reduce 4 doubles."""

    def __init__(self, engine, e_data, sig_eta, alpha=1.0, group=None, rank=0, world=1):
        self.engine, self.e_data, self.sig_eta, self.alpha = engine, e_data.contiguous(), float(sig_eta), float(alpha)
        self.group, self.rank, self.world = group, int(rank), int(world)

    def __call__(self, theta_mean, theta_sig, z_mean, z_sig, log_z_sig, logz_mean_post, logz_sig_post):
        B, S = theta_mean.shape[0], self.e_data.shape[0]
        lo, hi = shard_range(B * S, self.rank, self.world)
        if self.world > 1 and int(getattr(self.engine, "peer_world", 0)) == self.world:
            sums = self.engine.elbo_step2_totals(theta_mean.detach().contiguous(), theta_sig.detach().contiguous(),
                                                 self.e_data, lo, hi)
        else:
            sums, _ = self.engine.elbo_step2_partials(theta_mean.detach().contiguous(),
                                                      theta_sig.detach().contiguous(), self.e_data, lo, hi)
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        t5 = term5_from_sums(sums, z_mean, z_sig, B * S, self.sig_eta)
        return (term4(z_mean, log_z_sig) - t5) * self.alpha + add_loss(z_mean, z_sig, logz_mean_post, logz_sig_post)


def logz_posterior_moments(engine, theta_mean, theta_sig, e_data, eta_err, chunk=1 << 18):
    """The pre-pass of step 2 (main_custom_training.py:311-328): forward FEM over all
    num_data * ne_sam reparameterised samples, z = h + eta_err, moments of log z over the samples.
    theta_mean/theta_sig [D, 2], e_data [S, 2], eta_err [S, 2] (device tensors) ->
    (logz_mean_post [D, 2], logz_sig_post [D, 2])."""
    import torch

    D, S = theta_mean.shape[0], e_data.shape[0]
    theta = (e_data[None, :, :] * torch.sqrt(theta_sig)[:, None, :] + theta_mean[:, None, :]).reshape(-1, 2)
    hs = []
    for i in range(0, theta.shape[0], chunk):
        hs.append(engine.forward(theta[i:i + chunk].contiguous())[1])
    h = torch.cat(hs).reshape(D, S, 2)
    logz = torch.log(h + eta_err[None, :, :])
    return logz.mean(dim=1), logz.var(dim=1, unbiased=False)
