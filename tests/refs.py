"""Shared reference computations: literal float64 oracle vs the product on the same inputs.

Objective used for gradient parity: J = KL_term + <g_m, p_m> + <g_v, p_v> with fixed
random g (seed 0), standing in for the decoder's upstream gradients.
"""
import torch

from oracle import svgp_literal as lit
import svgp_vae_b200 as pkg

F64 = torch.float64


def upstream(shape, device="cpu"):
    g = torch.Generator().manual_seed(0)
    gm = torch.randn(*shape, generator=g, dtype=F64)
    gv = torch.randn(*shape, generator=g, dtype=F64)
    return gm.to(device), gv.to(device)


def make_pair(kind, cfg, device):
    """(oracle object, product object on `device`, differentiable parameters of each, same order)."""
    ctor = cfg["ctor"]
    if kind == "mnist":
        o = lit.MnistSVGP(name="o", **ctor)
        s = pkg.mnistSVGP(name="p", **ctor).to(device)
        op = [o.inducing_index_points, o.object_vectors, o.amplitude, o.l_GP]
        sp = [s.inducing_index_points, s.object_vectors, s.amplitude, s.l_GP]
    elif kind == "sprites":
        o = lit.SpritesSVGP(name="o", **ctor)
        s = pkg.spritesSVGP(name="p", **ctor).to(device)
        op = [o.inducing_index_points, o.GPLVM_action]
        sp = [s.inducing_index_points, s.GPLVM_action]
        if ctor.get("K_SE"):
            op += [o.sigma_action, o.l_action, o.sigma_character, o.l_character]
            sp += [s.sigma_action, s.l_action, s.sigma_character, s.l_character]
    elif kind == "sweep":
        # the sweep kernel is spritesSVGP's K_SE form on continuous [4 | 4] features: restated with the literal
        # SpritesSVGP by treating every row as an "inducing-layout" row
        o = _ContinuousSE(ctor["initial_inducing_points"], ctor["jitter"], ctor["N_train"], ctor["L"])
        s = pkg.productSVGP(**ctor).to(device)
        op = [o.inducing_index_points, o.hyp]
        sp = [s.inducing_index_points, s._hyp()]
    else:
        raise ValueError(kind)
    return o, s, op, sp


class _ContinuousSE(lit.SpritesSVGP):
    def __init__(self, Z, jitter, N_train, L):
        super().__init__(False, False, Z, "o", jitter, N_train, 4, torch.zeros(1, 4), 4, L, K_SE=True)
        self.hyp = torch.ones(4, dtype=F64)

    def kernel_matrix(self, x, y, x_inducing=True, y_inducing=True, diag_only=False):
        self.sigma_action, self.l_action, self.sigma_character, self.l_character = self.hyp
        return super().kernel_matrix(x, y, True, True, diag_only)


def oracle_objective(o, params, aux, y, noise, clip_pv=False):
    y = y.detach().to(F64).cpu().clone().requires_grad_(True)
    noise = noise.detach().to(F64).cpu().clone().requires_grad_(True)
    for t in params:
        t.requires_grad_(True)
    res = lit.minibatch_glue(o, aux.detach().to(F64).cpu(), y, noise, clip_pv=clip_pv)
    gm, gv = upstream(tuple(y.shape))
    J = res["KL_term"] + (gm * res["p_m"]).sum() + (gv * res["p_v"]).sum()
    grads = torch.autograd.grad(J, [y, noise] + list(params), allow_unused=True)
    return res, J, grads


def product_objective(s, params, aux, y, noise, clip_pv=False, **kw):
    y = y.detach().clone().requires_grad_(True)
    noise = noise.detach().clone().requires_grad_(True)
    res = s.elbo_step(aux, y, noise, clip_pv=clip_pv, **kw)
    gm, gv = upstream(tuple(y.shape), y.device)
    J = res["KL_term"] + (gm.to(res["p_m"].dtype) * res["p_m"]).sum().double() + (gv.to(res["p_v"].dtype) * res["p_v"]).sum().double()
    grads = torch.autograd.grad(J, [y, noise] + list(params), allow_unused=True)
    return res, J, grads
