"""Streamlined, jitter-exact form of the mini-batched SVGP step -- TEST INFRASTRUCTURE.

Same mathematics as ``svgp_literal.MiniBatchSVGP`` + ``minibatch_glue``
(SVGPVAE_model.py:220-343, :865-898, utils.py:483-504) for all L channels at once, but
organised the way the CUDA path computes it (DESIGN.md "collapsed sums"):

  * channel-independent factors once:  K, Kinv = inv(K + jI), logdet(K + jI)
  * the only reductions over datapoints are
        A_l = sum_i p_il k_i k_i^T   (L, M, M)      v_l = sum_i p_il y_il k_i   (L, M)
    and three scalar row sums per channel;
  * every *sum over datapoints* in L3 and in the cross-entropy term is rewritten as a
    trace against A_l / v_l (e.g. sum_i p_il k_i^T W k_i = <W, A_l>), so the (b, m, m)
    tensor of :286-294 never exists and the per-row O(M^2) work is only the predictive
    variance  p_v,il = kappa_i - k_i^T Kinv k_i + k_i^T S_l k_i;
  * the reference's jitter asymmetry is kept literally: S_l = inv(K + c A_l + jI) is built
    from the un-jittered K, mu_hat = c K S v and A_hat = K S K use the un-jittered K,
    a_l = Kinv mu_hat uses the jittered inverse (SVGPVAE_model.py:319,328,331,339-341).

float64 torch (CPU, or CUDA for the full-size probes); differentiable.  Pinning: see oracle/__init__.py.
"""
import math

import torch

from .svgp_literal import add_jitter, recip_no_nan

LOG_2PI = 1.8378770664093453


def streamlined_terms(K_nm, K_mm, kappa, y, noise, N_train, jitter, b_global=None, clip_pv=False,
                      A_extra=None):
    """All-channel SVGP step.

    K_nm (b, M), K_mm (M, M), kappa (b,) = diag K(x, x), y / noise (b, L).
    Returns dict(p_m, p_v (b, L), recon_l, kl_l, ce_l (L,), mu_hat (L, M), A_hat (L, M, M)).
    """
    b, M = K_nm.shape
    L = y.shape[1]
    bg = float(b if b_global is None else b_global)
    c = N_train / bg
    K = K_mm
    Kinv = torch.linalg.inv(add_jitter(K, jitter))
    ldK = 2 * torch.log(torch.diagonal(torch.linalg.cholesky(add_jitter(K, jitter)))).sum()

    p = recip_no_nan(noise)                                   # (b, L)
    A = torch.einsum('il,ia,ib->lab', p, K_nm, K_nm)          # (L, M, M)
    v = torch.einsum('il,ia->la', p * y, K_nm)                # (L, M)
    s_pk = (p * kappa[:, None]).sum(0)                        # (L,)
    s_pyy = (p * y * y).sum(0)
    s_log = torch.log(noise).sum(0)

    S = torch.linalg.inv(add_jitter(K + c * A, jitter))       # (L, M, M)
    w = c * torch.einsum('lab,lb->la', S, v)                  # S v scaled: p_m = K_nm w
    mu_hat = w @ K.T                                          # K w
    a = mu_hat @ Kinv.T                                       # Kinv mu_hat
    A_hat = K @ S @ K
    ld_Ahat = 2 * torch.log(torch.diagonal(torch.linalg.cholesky(add_jitter(A_hat, jitter)),
                                           dim1=-2, dim2=-1)).sum(-1)
    tr_KinvAhat = (Kinv * A_hat).sum((-1, -2))                # both symmetric
    kl_l = 0.5 * (ldK - ld_Ahat - M + tr_KinvAhat + (mu_hat * a).sum(-1))

    W = Kinv @ A_hat @ Kinv
    s_ph = (Kinv * A).sum((-1, -2))                           # sum_i p_il h_i,  h_i = k_i^T Kinv k_i
    s_t = (W * A).sum((-1, -2))                               # sum_i p_il k_i^T W_l k_i
    aAa = torch.einsum('la,lab,lb->l', a, A, a)
    recon_l = -0.5 * (s_pk - s_ph + s_t + s_log + bg_or(b, b_global) * LOG_2PI_exact()
                      + s_pyy - 2 * (a * v).sum(-1) + aAa)

    # per-row predictive moments (the only O(b M^2 L) work left)
    h = torch.einsum('ia,ab,ib->i', K_nm, Kinv, K_nm)
    q1 = torch.einsum('ia,lab,ib->il', K_nm, S, K_nm)
    p_m = K_nm @ w.T
    p_v_raw = kappa[:, None] - h[:, None] + q1
    p_v = torch.clamp(p_v_raw, 1e-4, 100.0) if clip_pv else p_v_raw

    # cross entropy, collapsed:  sum_i p (p_v + p_m^2 - 2 p_m y + y^2)
    s_ppv = s_pk - s_ph + (S * A).sum((-1, -2))
    if clip_pv:
        s_ppv = s_ppv + (p * (p_v - p_v_raw)).sum(0)
    wAw = torch.einsum('la,lab,lb->l', w, A, w)
    ce_l = -0.5 * (bg_or(b, b_global) * LOG_2PI + s_log + s_ppv + wAw - 2 * (w * v).sum(-1) + s_pyy)
    return dict(p_m=p_m, p_v=p_v, recon_l=recon_l, kl_l=kl_l, ce_l=ce_l, mu_hat=mu_hat, A_hat=A_hat,
                A=A, v=v, S=S, w=w)


def bg_or(b, b_global):
    return float(b if b_global is None else b_global)


def LOG_2PI_exact():
    # the reference's variational_loss uses tf.log(2*np.pi) (:298) while gauss_cross_entropy
    # hard-codes 1.8378770664093453 (utils.py:498); they agree to the last float64 digit.
    return math.log(2 * math.pi)


def glue_from_terms(t, b, N_train):
    """SVGPVAE_model.py:880-898 on top of ``streamlined_terms``."""
    recon, kl, ce = t['recon_l'].sum(), t['kl_l'].sum(), t['ce_l'].sum()
    inside = recon - (b / N_train) * kl
    return dict(inside_elbo_recon=recon, inside_elbo_kl=kl, inside_elbo=inside, ce_term=ce,
                KL_term=-ce + inside)
