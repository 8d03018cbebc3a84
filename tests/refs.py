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


def streamlined_objective(o, params, aux, y, noise, clip_pv=False):
    """oracle_objective through the COLLAPSED float64 restatement (oracle/svgp_streamlined.py: no (b, m, m) tensor, all
    channels at once; agrees with the literal one to 1e-12, tests/test_oracle.py) on the literal object's own kernel
    matrices -- for the shapes where the literal per-channel loop takes minutes (SPRITES M = 500 at L = 64)."""
    from oracle import svgp_streamlined as st
    y = y.detach().to(F64).cpu().clone().requires_grad_(True)
    noise = noise.detach().to(F64).cpu().clone().requires_grad_(True)
    for t in params:
        t.requires_grad_(True)
    x = aux.detach().to(F64).cpu()
    Z = o.inducing_index_points
    K_nm = o.kernel_matrix(x, Z, x_inducing=False, y_inducing=True)
    K_mm = o.kernel_matrix(Z, Z)
    kappa = o.kernel_matrix(x, x, x_inducing=False, y_inducing=False, diag_only=True)
    t = st.streamlined_terms(K_nm, K_mm, kappa, y, noise, o.N_train, o.jitter, clip_pv=clip_pv)
    res = dict(t)
    res.update(st.glue_from_terms(t, float(x.shape[0]), o.N_train))
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


# ---- SVIGP_Hensman (tests/golden/make_svigp_golden.py uses the same seeded variational parameters) -----------------
def svigp_case(normalize, device, fixture):
    """-> (product SVIGP_Hensman on `device`, aux, upstream g (b, L)) for the inputs of svigp_golden.npz."""
    from svgp_vae_b200 import configs
    from svgp_vae_b200.svigp import SVIGP_Hensman
    L = 3
    cfg = configs.mnist_inputs(fixture, L=L, b=192, normalize=normalize)
    c = cfg["ctor"]
    s = SVIGP_Hensman(fixed_inducing_points=False, initial_inducing_points=c["initial_inducing_points"], name="p",
                      jitter=c["jitter"], N_train=c["N_train"], dtype=torch.float64, L=L, fixed_gp_params=False,
                      object_vectors_init=c["object_vectors_init"], K_obj_normalize=normalize)
    m = s.nr_inducing
    g = torch.Generator().manual_seed(3)
    mu = 0.5 * torch.randn(L, m, generator=g, dtype=F64)
    A = torch.eye(m, dtype=F64).repeat(L, 1, 1) * 0.7 + 0.05 * torch.tril(torch.randn(L, m, m, generator=g, dtype=F64))
    with torch.no_grad():
        s.GP_var_params_mu.copy_(mu)
        s.GP_var_params_A.copy_(A)
        s.Hensman_likelihood_noise.fill_(0.3)
    g2 = torch.Generator().manual_seed(5)
    gm = torch.randn(192, L, generator=g2, dtype=F64)
    return s.to(device), cfg["aux"].to(F64).to(device), gm.to(device)


def svigp_check(s, aux, gm, gold, key, tol, rel_err, batched=True):
    """Values and gradients of SVIGP_Hensman against the reference-source golden vectors."""
    if batched:
        rec, kl, means = s.variational_loss_all(aux)
    else:
        outs = [s.variational_loss(aux, None, l) for l in range(s.L)]
        rec, kl, means = torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]), torch.stack([o[2] for o in outs], 1)
    g = lambda n: torch.from_numpy(gold[key + "/" + n])
    assert rel_err(rec, g("L3")) < tol and rel_err(kl, g("KL")) < tol and rel_err(means, g("mean")) < tol
    b = float(aux.shape[0])
    J = rec.sum() - (b / s.N_train) * kl.sum() + (gm * means).sum()
    assert abs(float(J) - float(g("J"))) < tol * abs(float(g("J")))
    leaves = [s.inducing_index_points, s.object_vectors, s.amplitude, s.l_GP, s.noise, s.GP_var_params_mu, s.GP_var_params_A]
    grads = torch.autograd.grad(J, leaves)
    for n, gr in zip(["Z", "table", "amplitude", "length", "noise", "mu", "A"], grads):
        assert rel_err(gr, g("grad_" + n)) < (max(tol, 1e-4) if gr.numel() == 1 else tol), n
    test_aux = aux[:40].clone()
    test_aux[:, 1] = test_aux[:, 1] + 0.3
    mv, B = s.approximate_posterior_params(test_aux, 1)
    assert rel_err(mv, g("post_mean")) < tol and rel_err(B, g("post_B")) < tol
