"""Prediction-time entry points of the SVGP object, all latent channels at once (no gradients).

Reference call sites (the per-channel Python loops these replace):

  precompute_GP_params_SVGPVAE            SVGPVAE_model.py:989-1023   (L x m) mean terms, (L x m x m) Sigma_l^-1,
                                                                      built from ALL training encodings; note :1014 inverts
                                                                      Sigma_l WITHOUT jitter and without the N_train/b factor
  predict_from_precomputed                :610-635 called per channel at :1165-1168
  posterior_predict                       :1048-1050 (bacthing_predict_SVGPVAE_rotated_mnist): approximate_posterior_params
                                          (:303-343) with test index points against the full training set

Same kernels as the training step (K1 builder, K2 SYRK on tcgen05, float64 Cholesky stage, K4 row quadratic forms).
"""
import torch

from . import ops
from .backend import get_backend


def _recip_no_nan(x):
    safe = torch.where(x == 0, torch.ones_like(x), x)
    return torch.where(x == 0, torch.zeros_like(x), 1.0 / safe)


def _operands(svgp, aux_data):
    be = get_backend()
    spec = svgp._spec()
    hyp = svgp._hyp().detach().float().contiguous()
    Fx = svgp._features(aux_data, False).detach().float().contiguous()
    Fz = svgp._features(svgp.inducing_index_points, True).detach().float().contiguous()
    kop = be.kernel_fwd(spec, Fx, Fz, hyp, tc=be.want_tc(Fx.shape[0], Fz.shape[0]))
    return be, spec, hyp, Fx, Fz, kop


@torch.no_grad()
def precompute_GP_params_SVGPVAE(means, vars, aux_data, svgp):
    """(mean_terms (L, m), inv_Sigma_l (L, m, m)) -- SVGPVAE_model.py:989-1023, every channel in one pass."""
    be, spec, hyp, Fx, Fz, kop = _operands(svgp, aux_data)
    K_mm = be.kernel_fwd(spec, Fz, Fz, hyp, tc=False).K.double()
    p = _recip_no_nan(vars.detach().float()).contiguous()
    A = be.syrk(kop, p)                                                   # K_mn diag(1/var_l) K_nm          :1013
    V = be.gemm_tn(kop, (p * means.detach().float()).contiguous())        # K_mn (means_l / var_l)           :1015-1016
    Sigma_inv, _, _ = ops.spd_inverse_logdet(K_mm.unsqueeze(0) + A)       # :1014 -- no jitter
    mean_terms = ops.bmv64(Sigma_inv, V)
    return mean_terms.to(svgp.dtype), Sigma_inv.to(svgp.dtype)


@torch.no_grad()
def predict_from_precomputed(svgp, index_points, mean_terms, sigma_terms, K_mm_inv=None):
    """All channels of approximate_posterior_params_precomputed_GP_posterior_params (:610-635):
    mean (b, L) = K_bm mean_terms_l,  B (b, L) = K_bb - diag(K_bm K_mm^-1 K_mb) + diag(K_bm sigma_term_l K_mb)."""
    be, spec, hyp, Fx, Fz, kop = _operands(svgp, index_points)
    if K_mm_inv is None:
        K_mm = be.kernel_fwd(spec, Fz, Fz, hyp, tc=False).K.double()
        eye = torch.eye(K_mm.shape[0], dtype=K_mm.dtype, device=K_mm.device)
        K_mm_inv = ops.spd_inverse_logdet((K_mm + svgp.jitter * eye).unsqueeze(0))[0][0]      # :623-625 (with jitter)
    K_bb = be.kernel_diag_fwd(spec, Fx, Fx, hyp)
    mean = be.gemm_nn(kop, mean_terms.detach().float().contiguous())
    h = be.rowquad(kop, K_mm_inv.detach().double().unsqueeze(0).contiguous())
    q = be.rowquad(kop, sigma_terms.detach().double().contiguous())
    B = K_bb[:, None] - h + q
    return mean.to(svgp.dtype), B.to(svgp.dtype)


@torch.no_grad()
def posterior_predict(svgp, index_points_test, index_points_train, means, vars):
    """(p_m (x, L), p_v (x, L)): approximate_posterior_params (:303-343) for every channel, test points against the
    given training encodings -- the body of the L-loop of bacthing_predict_SVGPVAE_rotated_mnist (:1048-1050)."""
    be, spec, hyp, Fn, Fz, kop_n = _operands(svgp, index_points_train)
    Fx = svgp._features(index_points_test, False).detach().float().contiguous()
    kop_x = be.kernel_fwd(spec, Fx, Fz, hyp, tc=be.want_tc(Fx.shape[0], Fz.shape[0]))
    K_mm = be.kernel_fwd(spec, Fz, Fz, hyp, tc=False).K.double()
    eye = torch.eye(K_mm.shape[0], dtype=K_mm.dtype, device=K_mm.device)
    c = svgp.N_train / float(index_points_train.shape[0])                 # :316, :328
    p = _recip_no_nan(vars.detach().float()).contiguous()
    A = be.syrk(kop_n, p)
    V = be.gemm_tn(kop_n, (p * means.detach().float()).contiguous())
    K_mm_inv = ops.spd_inverse_logdet((K_mm + svgp.jitter * eye).unsqueeze(0))[0]              # :319
    S = ops.spd_inverse_logdet(K_mm.unsqueeze(0) + c * A + svgp.jitter * eye)[0]             # :328-331
    w = c * ops.bmv64(S, V)
    K_xx = be.kernel_diag_fwd(spec, Fx, Fx, hyp)
    p_m = be.gemm_nn(kop_x, w.float().contiguous())                                          # :332-334
    p_v = K_xx[:, None] - be.rowquad(kop_x, K_mm_inv.contiguous()) + be.rowquad(kop_x, S.contiguous())   # :336-337
    return p_m.to(svgp.dtype), p_v.to(svgp.dtype)
