"""CUDA-graph capture of the batched step for the reference's own (small) configurations.

At the reference's sizes -- rotated MNIST b = 256, m = 32, L = 16; SPRITES b = 500, M = 72, L = 64 -- one ELBO +
gradient step is ~550 kernel launches of a few microseconds each: the step is launch-bound, not compute-bound.
``GraphedElboStep`` captures the forward and the backward of ``mainSVGP.elbo_step`` into two CUDA graphs
(``torch.cuda.make_graphed_callables``: static input / output buffers, autograd-aware, the decoder still runs between
the two as ordinary PyTorch) and replays them; shapes, ``clip_pv`` and the step's keyword arguments are fixed at
capture time.  Replaces: the per-step ``sess.run`` of the reference's TF-1.15 graph (MNIST_experiment.py:318-330).

Under capture no host synchronisation is allowed, so the Cholesky factorisations record a non-positive pivot in a static
device word instead of raising (ops.pd_flag); ``pd_check()`` reads it (one small device->host copy) and raises
``ops.NotPositiveDefinite`` -- call it every few steps, or pass ``pd_check_every=k`` to have ``__call__`` do so.
"""
import torch

from .step import elbo_terms


class _StepModule(torch.nn.Module):
    def __init__(self, svgp, clip_pv, kw):
        super().__init__()
        self.svgp, self.clip_pv, self.kw = svgp, clip_pv, kw

    def forward(self, aux, qnet_mu, qnet_var):
        res = self.svgp.elbo_step(aux, qnet_mu, qnet_var, clip_pv=self.clip_pv, **self.kw)
        return res["p_m"], res["p_v"], res["recon_l"], res["kl_l"], res["ce_l"]


class GraphedElboStep:
    """``step = GraphedElboStep(svgp, aux, qnet_mu, qnet_var); res = step(aux, qnet_mu, qnet_var)`` -- same dict as
    ``svgp.elbo_step`` (without mu_hat / A_hat), differentiable w.r.t. qnet_mu, qnet_var and the module's parameters."""

    def __init__(self, svgp, aux, qnet_mu, qnet_var, clip_pv=False, pd_check_every=0, **kw):
        if getattr(svgp, "titsias", False):
            raise ValueError("the Titsias branch runs the reference's per-channel loop; it is not captured")
        if kw.get("group") is not None:
            raise ValueError("the sharded step synchronises with the host (global batch size); capture the single-process step")
        from . import ops
        self.svgp = svgp
        self.b = float(aux.shape[0])
        self._device, self._every, self._calls = aux.device, int(pd_check_every), 0
        ops.pd_flag(aux.device)                   # before capture: the captured factorisations write their status here
        self._eager = _StepModule(svgp, clip_pv, dict(kw, return_A_hat=False))
        sample = (aux.detach().clone(), qnet_mu.detach().clone().requires_grad_(True), qnet_var.detach().clone().requires_grad_(True))
        self._graphed = torch.cuda.make_graphed_callables(self._eager, sample, num_warmup_iters=3, allow_unused_input=True)

    def __call__(self, aux, qnet_mu, qnet_var):
        pm, pv, recon, kl, ce = self._graphed(aux, qnet_mu, qnet_var)
        self._calls += 1
        if self._every and self._calls % self._every == 0:
            self.pd_check()
        res = dict(p_m=pm, p_v=pv, recon_l=recon, kl_l=kl, ce_l=ce)
        res.update(elbo_terms(res, self.b, self.svgp.N_train))
        return res

    def pd_check(self):
        """Raise ops.NotPositiveDefinite if a replayed step met a non-positive Cholesky pivot (reads one device word)."""
        from . import ops
        ops.pd_flag_check(self._device)

    def check(self, aux, qnet_mu, qnet_var):
        """One eager step with the positive-definiteness checks on (raises ops.NotPositiveDefinite)."""
        with torch.no_grad():
            self._eager(aux, qnet_mu, qnet_var)
