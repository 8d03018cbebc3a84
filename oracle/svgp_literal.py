"""Literal float64 restatement of the reference SVGP objects -- TEST INFRASTRUCTURE.

Follows, operation for operation (explicit inverses, the (b, m, m) lambda tensor, the
asymmetric jitter, the ball-KL quirks), these reference ranges:

  SVGPVAE_model.py:13-14     add_jitter
  SVGPVAE_model.py:17-171    BallSVGP            (class ``SVGP``)
  SVGPVAE_model.py:174-378   MiniBatchSVGP       (class ``mainSVGP``)
  SVGPVAE_model.py:381-484   MnistSVGP           (class ``mnistSVGP``)
  SVGPVAE_model.py:487-635   SpritesSVGP         (class ``spritesSVGP``)
  utils.py:483-504           gauss_cross_entropy
  SVGPVAE_model.py:865-898   minibatch_glue      (L-loop of forward_pass_SVGPVAE)
  SVGPVAE_model.py:674-697   ball_glue           (build_SVGPVAE_elbo_graph)

Everything is torch-CPU float64 so that gradients of any output can be taken with
torch autograd (the reference uses tf.gradients, MNIST_experiment.py:202-208).
Parameters are plain tensors; set ``requires_grad_`` on the ones to differentiate.

Pinned to the reference source under the TF shim; unpinned w.r.t. TensorFlow itself (see oracle/__init__.py).
"""
import math

import numpy as np
import torch

from . import tfp_kernels as tfk

F64 = torch.float64
LOG_2PI = 1.8378770664093453      # utils.py:498


def as64(x):
    if isinstance(x, torch.Tensor):
        return x.to(F64)
    return torch.as_tensor(x, dtype=F64)


def add_jitter(mat, jitter):
    """SVGPVAE_model.py:13-14 (set_diag(A, diag(A) + j))."""
    m = mat.shape[-1]
    return mat + jitter * torch.eye(m, dtype=mat.dtype, device=mat.device)


def recip_no_nan(x):
    """tf.math.reciprocal_no_nan: 1/x, and 0 where x == 0 (SVGPVAE_model.py:78,282,330)."""
    safe = torch.where(x == 0, torch.ones_like(x), x)
    return torch.where(x == 0, torch.zeros_like(x), 1.0 / safe)


def matvec(a, v):
    """tf.linalg.matvec with broadcasting batch dims: out[..., i] = sum_j a[..., i, j] v[..., j]."""
    return (a * v.unsqueeze(-2)).sum(-1)


def gauss_cross_entropy(mu1, var1, mu2, var2):
    """utils.py:483-504: E_{N(mu1,var1)}[log N(z | mu2, var2)], element-wise."""
    quad = (var1 + mu1 ** 2 - 2 * mu1 * mu2 + mu2 ** 2) / var2
    return -0.5 * (LOG_2PI + torch.log(var2) + quad)


# --------------------------------------------------------------------------------------
# moving ball: one GP per video, 1-D RBF over time           SVGPVAE_model.py:17-171
# --------------------------------------------------------------------------------------
class BallSVGP:
    def __init__(self, titsias, num_inducing_points, fixed_inducing_points, tmin, tmax, vidlt,
                 fixed_gp_params, name, jitter, ip_min, ip_max, GP_init):
        # :36-58 -- inducing times are linspace(tmin,tmax,m) when fixed, else linspace(ip_min,ip_max,m)
        self.titsias = titsias
        self.num_inducing_points = num_inducing_points
        self.jitter = jitter
        lo, hi = (tmin, tmax) if fixed_inducing_points else (ip_min, ip_max)
        # :46 / :49 -- the reference builds the grid with np.linspace(..., dtype=np.float32): the inducing times ARE
        # float32 numbers (3.0714285...), so the float64 oracle starts from the same rounded values
        self.inducing_index_points = torch.from_numpy(
            np.linspace(lo, hi, num_inducing_points, dtype=np.float32).astype(np.float64))
        self.l_GP = as64(vidlt if fixed_gp_params else GP_init)
        # :60 amplitude=None
        self.kernel = tfk.ExponentiatedQuadratic(amplitude=None, length_scale=self.l_GP)

    def _mats(self, x):
        z = self.inducing_index_points
        self.kernel.length_scale = self.l_GP
        K_mm = self.kernel.matrix(z[:, None], z[:, None])                       # :81 / :152
        K_mm_inv = torch.linalg.inv(add_jitter(K_mm, self.jitter))              # :83 / :154
        K_nn = self.kernel.matrix(x[..., None], x[..., None])                   # :84 / :155
        K_nm = self.kernel.matrix(x[..., None], z[:, None])                     # :86 / :157
        return K_mm, K_mm_inv, K_nn, K_nm, K_nm.transpose(-1, -2)

    def approximate_posterior_params(self, index_points, y=None, noise=None):
        """:141-171 -> mean (B,T), cov (B,T,T), mu_hat (B,m), A_hat (B,m,m)."""
        K_mm, K_mm_inv, K_nn, K_nm, K_mn = self._mats(index_points)
        prec = recip_no_nan(noise)
        sigma_l = K_mm + K_mn @ (torch.diag_embed(prec) @ K_nm)                 # :160 (no N/b factor)
        sigma_l_inv = torch.linalg.inv(add_jitter(sigma_l, self.jitter))        # :161
        KSK = K_nm @ (sigma_l_inv @ K_mn)                                       # :162
        mean_vector = matvec(KSK, prec * y)                                     # :164
        B = K_nn - K_nm @ (K_mm_inv @ K_mn) + KSK                               # :165
        mu_hat = matvec(K_mm @ (sigma_l_inv @ K_mn), prec * y)                  # :167
        A_hat = K_mm @ (sigma_l_inv @ K_mm)                                     # :169
        return mean_vector, B, mu_hat, A_hat

    def variational_loss(self, x, y, noise, mu_hat, A_hat):
        """:62-139 (note the argument order differs from the mini-batched class)."""
        T = float(x.shape[1])
        m = float(self.inducing_index_points.shape[0])
        prec = recip_no_nan(noise)                                              # :78
        K_mm, K_mm_inv, K_nn, K_nm, K_mn = self._mats(x)
        Q_nn = K_nm @ (K_mm_inv @ K_mn)

        if self.titsias:                                                        # :89-101
            cov = torch.diag_embed(noise) + Q_nn
            cov_j = add_jitter(cov, self.jitter)
            cov_inv = torch.linalg.inv(cov_j)
            chol = torch.linalg.cholesky(cov_j)
            logdet = 2 * torch.log(torch.diagonal(chol, dim1=-2, dim2=-1)).sum(1)
            trace_term = prec * torch.diagonal(K_nn - Q_nn, dim1=-2, dim2=-1)
            L2 = -0.5 * (T * math.log(2 * math.pi) + logdet
                         + (y * matvec(cov_inv, y)).sum(1) + trace_term.sum(1))
            return L2, torch.zeros((), dtype=F64)

        mean_vector = matvec(K_nm, matvec(K_mm_inv, mu_hat))                    # :106
        K_tilde = prec * torch.diagonal(K_nn - Q_nn, dim1=-2, dim2=-1)          # :109
        lam = K_nm.unsqueeze(3) @ K_nm.unsqueeze(3).transpose(-1, -2)           # :112  (B,T,m,m)
        lam = K_mm_inv @ (lam @ K_mm_inv)                                       # :116
        A_rep = A_hat.unsqueeze(1).expand(-1, x.shape[1], -1, -1)               # :120
        trace_terms = prec * torch.diagonal(A_rep @ lam, dim1=-2, dim2=-1).sum(-1)   # :121
        L3 = -0.5 * (K_tilde.sum(1) + trace_terms.sum(1) + torch.log(noise).sum(1)
                     + T * math.log(2 * math.pi) + (prec * (y - mean_vector) ** 2).sum(1))   # :124-126

        K_chol = torch.linalg.cholesky(add_jitter(K_mm, self.jitter))           # :129
        S_chol = torch.linalg.cholesky(add_jitter(A_hat, self.jitter))          # :130
        K_logdet = 2 * torch.log(torch.diagonal(K_chol)).sum()                  # :131
        S_logdet = 2 * torch.log(torch.diagonal(S_chol, dim1=-2, dim2=-1)).sum()    # :132 (whole batch!)
        # :134-137 -- the quadratic term is taken on A_hat (not mu_hat) and summed over the batch
        quirk = (A_hat * matvec(K_mm_inv, A_hat)).sum()
        KL = 0.5 * (K_logdet - S_logdet - m
                    + torch.diagonal(K_mm_inv @ A_hat, dim1=-2, dim2=-1).sum(-1) + quirk)
        return L3, KL


# --------------------------------------------------------------------------------------
# mini-batched SVGP                                            SVGPVAE_model.py:174-378
# --------------------------------------------------------------------------------------
class MiniBatchSVGP:
    def __init__(self, titsias, fixed_inducing_points, initial_inducing_points, name, jitter,
                 N_train, dtype, L, K_obj_normalize=False):
        self.jitter = jitter
        self.titsias = titsias
        self.nr_inducing = len(initial_inducing_points)
        self.N_train = N_train
        self.L = L
        self.K_obj_normalize = K_obj_normalize
        self.inducing_index_points = as64(initial_inducing_points).clone()

    def kernel_matrix(self, x, y, x_inducing=True, y_inducing=True, diag_only=False):
        raise NotImplementedError

    def approximate_posterior_params(self, index_points_test, index_points_train=None, y=None, noise=None):
        """:303-343."""
        Z = self.inducing_index_points
        b = float(index_points_train.shape[0])
        c = self.N_train / b
        K_mm = self.kernel_matrix(Z, Z)                                                   # :318
        K_mm_inv = torch.linalg.inv(add_jitter(K_mm, self.jitter))                        # :319
        K_xx = self.kernel_matrix(index_points_test, index_points_test, False, False, True)   # :320
        K_xm = self.kernel_matrix(index_points_test, Z, x_inducing=False)                 # :322
        K_mx = K_xm.T
        K_nm = self.kernel_matrix(index_points_train, Z, x_inducing=False)                # :325
        K_mn = K_nm.T
        prec = recip_no_nan(noise)
        sigma_l = K_mm + c * (K_mn @ (K_nm * prec[:, None]))                              # :328-330
        sigma_l_inv = torch.linalg.inv(add_jitter(sigma_l, self.jitter))                  # :331
        mean_vector = c * matvec(K_xm, matvec(sigma_l_inv, matvec(K_mn, prec * y)))       # :332-334
        KSK = K_xm @ (sigma_l_inv @ K_mx)                                                 # :336
        B = K_xx + torch.diagonal(-(K_xm @ (K_mm_inv @ K_mx)) + KSK)                      # :337
        mu_hat = c * matvec(K_mm @ (sigma_l_inv @ K_mn), prec * y)                        # :339-340
        A_hat = K_mm @ (sigma_l_inv @ K_mm)                                               # :341
        return mean_vector, B, mu_hat, A_hat

    def variational_loss(self, x, y, mu_hat, A_hat, noise=None):
        """:220-301."""
        Z = self.inducing_index_points
        b = float(x.shape[0])
        m = float(Z.shape[0])
        K_mm = self.kernel_matrix(Z, Z)                                                   # :238
        K_mm_inv = torch.linalg.inv(add_jitter(K_mm, self.jitter))                        # :239
        K_nn = self.kernel_matrix(x, x, False, False, True)                               # :241
        K_nm = self.kernel_matrix(x, Z, x_inducing=False)                                 # :243
        K_mn = K_nm.T

        if self.titsias:                                                                  # :246-259
            Q = K_nm @ (K_mm_inv @ K_mn)
            cov_j = add_jitter(torch.diag(noise) + Q, self.jitter)
            trace_term = recip_no_nan(noise) * (K_nn - torch.diagonal(Q))
            cov_inv = torch.linalg.inv(cov_j)
            logdet = 2 * torch.log(torch.diagonal(torch.linalg.cholesky(cov_j))).sum()
            L2 = -0.5 * (b * math.log(2 * math.pi) + logdet + (y * matvec(cov_inv, y)).sum()
                         + trace_term.sum())
            return L2, torch.zeros((), dtype=F64)

        mean_vector = matvec(K_nm, matvec(K_mm_inv, mu_hat))                              # :264-265
        K_chol = torch.linalg.cholesky(add_jitter(K_mm, self.jitter))                     # :270
        S_chol = torch.linalg.cholesky(add_jitter(A_hat, self.jitter))                    # :271-272
        K_logdet = 2 * torch.log(torch.diagonal(K_chol)).sum()
        S_logdet = 2 * torch.log(torch.diagonal(S_chol)).sum()
        KL = 0.5 * (K_logdet - S_logdet - m + torch.trace(K_mm_inv @ A_hat)
                    + (mu_hat * matvec(K_mm_inv, mu_hat)).sum())                          # :276-279
        prec = recip_no_nan(noise)                                                        # :282
        K_tilde = prec * (K_nn - torch.diagonal(K_nm @ (K_mm_inv @ K_mn)))                # :284
        lam = K_nm.unsqueeze(2) @ K_nm.unsqueeze(2).transpose(1, 2)                       # :287  (b,m,m)
        lam = K_mm_inv @ (lam @ K_mm_inv)                                                 # :291
        trace_terms = prec * torch.diagonal(A_hat @ lam, dim1=-2, dim2=-1).sum(-1)        # :294
        L3 = -0.5 * (K_tilde.sum() + trace_terms.sum() + torch.log(noise).sum()
                     + b * math.log(2 * math.pi) + (prec * (y - mean_vector) ** 2).sum())  # :297-299
        return L3, KL

    def mean_vector_bias_analysis(self, index_points, y=None, noise=None):
        """:345-370."""
        Z = self.inducing_index_points
        c = self.N_train / float(index_points.shape[0])
        K_mm = self.kernel_matrix(Z, Z)
        K_bm = self.kernel_matrix(index_points, Z, x_inducing=False)
        prec = recip_no_nan(noise)
        sigma_l = K_mm + c * (K_bm.T @ (torch.diag(prec) @ K_bm))
        sigma_l_inv = torch.linalg.inv(add_jitter(sigma_l, self.jitter))
        return c * matvec(K_mm @ (sigma_l_inv @ K_bm.T), prec * y)


def _row_norm(v):
    return torch.sqrt((v * v).sum(1))          # tf.math.reduce_euclidean_norm(axis=1)


class MnistSVGP(MiniBatchSVGP):
    """:381-484.  aux row = [id, angle, 8 PCA dims]; inducing row = [unused id, angle, 8 dims]."""

    def __init__(self, titsias, fixed_inducing_points, initial_inducing_points, fixed_gp_params,
                 object_vectors_init, name, jitter, N_train, L, K_obj_normalize):
        super().__init__(titsias, fixed_inducing_points, initial_inducing_points, name, jitter,
                         N_train, F64, L, K_obj_normalize)
        self.l_GP = torch.tensor(1.0, dtype=F64)                                          # :409-413
        self.amplitude = torch.tensor(1.0, dtype=F64)
        self.object_vectors = None if object_vectors_init is None else as64(object_vectors_init).clone()

    def kernel_matrix(self, x, y, x_inducing=True, y_inducing=True, diag_only=False):
        view = tfk.ExpSinSquared(amplitude=self.amplitude, length_scale=self.l_GP, period=2 * math.pi)  # :416
        lin = tfk.Linear()                                                                # :417
        x_view, y_view = x[:, 1], y[:, 1]
        if self.object_vectors is None:                                                   # :444-445
            x_obj, y_obj = x[:, 2:], y[:, 2:]
        else:                                                                             # :447-455
            x_obj = x[:, 2:] if x_inducing else self.object_vectors[x[:, 0].long()]
            y_obj = y[:, 2:] if y_inducing else self.object_vectors[y[:, 0].long()]
        if diag_only:                                                                     # :458-467
            view_k = view.apply(x_view[:, None], y_view[:, None])
            obj_k = lin.apply(x_obj, y_obj)
            if self.K_obj_normalize:
                obj_k = obj_k / (_row_norm(x_obj) * _row_norm(y_obj))
        else:                                                                             # :461,469-474
            view_k = view.matrix(x_view[:, None], y_view[:, None])
            obj_k = lin.matrix(x_obj, y_obj)
            if self.K_obj_normalize:
                obj_k = obj_k * (1.0 / (_row_norm(x_obj)[:, None] @ _row_norm(y_obj)[None, :]))
        return view_k * obj_k                                                             # :476


class SpritesSVGP(MiniBatchSVGP):
    """:487-635.  aux row = [action id, 16-d character vector]; inducing row = [8-d action, 16-d character]."""

    def __init__(self, titsias, fixed_inducing_points, initial_inducing_points, name, jitter, N_train,
                 L_action, initial_GPLVM_action, L_character, L, fixed_GP_params=False,
                 fixed_GPLVM=False, K_obj_normalize=False, K_SE=False):
        super().__init__(titsias, fixed_inducing_points, initial_inducing_points, name, jitter,
                         N_train, F64, L, K_obj_normalize)
        self.L_action, self.L_character, self.K_SE = L_action, L_character, K_SE
        self.GPLVM_action = as64(initial_GPLVM_action).clone()
        if K_SE:                                                                          # :530-544
            self.l_action = torch.tensor(1.0, dtype=F64)
            self.sigma_action = torch.tensor(0.1, dtype=F64)
            self.l_character = torch.tensor(1.0, dtype=F64)
            self.sigma_character = torch.tensor(0.1, dtype=F64)

    def _kernels(self):
        if self.K_SE:
            return (tfk.ExponentiatedQuadratic(self.sigma_action, self.l_action),
                    tfk.ExponentiatedQuadratic(self.sigma_character, self.l_character))
        return tfk.Linear(), tfk.Linear()                                                 # :547-548

    def kernel_matrix(self, x, y, x_inducing=True, y_inducing=True, diag_only=False):
        k_act, k_chr = self._kernels()
        La = self.L_action
        if x_inducing:                                                                    # :562-565
            xa, xc = x[:, :La], x[:, La:]
        else:
            xa, xc = self.GPLVM_action[x[:, 0].long()], x[:, 1:]
        if y_inducing:                                                                    # :567-570
            ya, yc = y[:, :La], y[:, La:]
        else:
            ya, yc = self.GPLVM_action[y[:, 0].long()], y[:, 1:]
        normalise = (not self.K_SE) and self.K_obj_normalize
        if diag_only:                                                                     # :572-583
            ka, kc = k_act.apply(xa, ya), k_chr.apply(xc, yc)
            if normalise:
                ka = ka / (_row_norm(xa) * _row_norm(ya))
                kc = kc / (_row_norm(xc) * _row_norm(yc))
        else:                                                                             # :585-598
            ka, kc = k_act.matrix(xa, ya), k_chr.matrix(xc, yc)
            if normalise:
                ka = ka * (1.0 / (_row_norm(xa)[:, None] @ _row_norm(ya)[None, :]))
                kc = kc * (1.0 / (_row_norm(xc)[:, None] @ _row_norm(yc)[None, :]))
        return ka * kc                                                                    # :600

    def approximate_posterior_params_precomputed_GP_posterior_params(self, index_points, mean_term,
                                                                     sigma_term, K_mm_inv=None):
        """:610-635."""
        Z = self.inducing_index_points
        if K_mm_inv is None:
            K_mm_inv = torch.linalg.inv(add_jitter(self.kernel_matrix(Z, Z), self.jitter))
        K_bb = self.kernel_matrix(index_points, index_points, False, False, True)
        K_bm = self.kernel_matrix(index_points, Z, x_inducing=False)
        mean_vector = matvec(K_bm, mean_term)
        B = K_bb + torch.diagonal(-(K_bm @ (K_mm_inv @ K_bm.T)) + K_bm @ (sigma_term @ K_bm.T))
        return mean_vector, B


# --------------------------------------------------------------------------------------
# call sites
# --------------------------------------------------------------------------------------
def minibatch_glue(svgp, aux_data, qnet_mu, qnet_var, clip_pv=False):
    """SVGPVAE_model.py:865-898: the per-latent-channel loop and the ELBO bookkeeping.

    ``clip_pv`` mirrors ``if repr_NN: p_v = clip(p_v, 1e-4, 100)`` (:891-892, SPRITES only).
    Returns a dict with p_m, p_v (b, L) and the scalars the caller logs / differentiates.
    """
    b = float(aux_data.shape[0])
    recon, kl, p_m, p_v, mu_hats, A_hats = [], [], [], [], [], []
    for l in range(qnet_mu.shape[1]):                                                     # :868
        pm_l, pv_l, mu_hat_l, A_hat_l = svgp.approximate_posterior_params(
            aux_data, aux_data, qnet_mu[:, l], qnet_var[:, l])
        rec_l, kl_l = svgp.variational_loss(x=aux_data, y=qnet_mu[:, l], noise=qnet_var[:, l],
                                            mu_hat=mu_hat_l, A_hat=A_hat_l)
        recon.append(rec_l); kl.append(kl_l); p_m.append(pm_l); p_v.append(pv_l)
        mu_hats.append(mu_hat_l); A_hats.append(A_hat_l)
    recon_l, kl_l = torch.stack(recon), torch.stack(kl)
    inside_elbo_recon, inside_elbo_kl = recon_l.sum(), kl_l.sum()                         # :880-881
    if svgp.titsias:
        inside_elbo = inside_elbo_recon - inside_elbo_kl                                  # :884
    else:
        inside_elbo = inside_elbo_recon - (b / svgp.N_train) * inside_elbo_kl             # :886
    p_m, p_v = torch.stack(p_m, 1), torch.stack(p_v, 1)                                   # :888-889
    if clip_pv:
        p_v = torch.clamp(p_v, 1e-4, 100.0)                                               # :891-892
    ce_term = gauss_cross_entropy(p_m, p_v, qnet_mu, qnet_var).sum()                      # :895-896
    KL_term = -ce_term + inside_elbo                                                      # :898
    return dict(p_m=p_m, p_v=p_v, inside_elbo_recon=inside_elbo_recon, inside_elbo_kl=inside_elbo_kl,
                inside_elbo=inside_elbo, ce_term=ce_term, KL_term=KL_term,
                recon_l=recon_l, kl_l=kl_l, mu_hat=torch.stack(mu_hats), A_hat=torch.stack(A_hats))


def ball_glue(svgp_x, svgp_y, qnet_mu, qnet_var, tmax=None):
    """SVGPVAE_model.py:663-664, 674-697, 709 with qnet_* (batch, tmax, 2) standing in for the encoder."""
    batch, tmax = qnet_mu.shape[0], qnet_mu.shape[1]
    batch_T = (torch.arange(tmax, dtype=F64) + 1.0).repeat(batch, 1)                      # :663-664
    out = {}
    recon, kl, pms, pvs = 0.0, 0.0, [], []
    for ch, svgp in enumerate((svgp_x, svgp_y)):
        pm, B, mu_hat, A_hat = svgp.approximate_posterior_params(
            batch_T, y=qnet_mu[:, :, ch], noise=qnet_var[:, :, ch])                       # :674-677
        rec, k = svgp.variational_loss(batch_T, qnet_mu[:, :, ch], qnet_var[:, :, ch],
                                       mu_hat=mu_hat, A_hat=A_hat)                        # :680-683
        recon = recon + rec; kl = kl + k
        pms.append(pm); pvs.append(torch.diagonal(B, dim1=-2, dim2=-1))                   # :692-693
        out["B_%d" % ch] = B; out["mu_hat_%d" % ch] = mu_hat; out["A_hat_%d" % ch] = A_hat
    inside_elbo = recon - kl                                                              # :684-686
    full_p_mu, full_p_var = torch.stack(pms, 2), torch.stack(pvs, 2)
    ce = -gauss_cross_entropy(full_p_mu, full_p_var, qnet_mu, qnet_var).sum((1, 2))       # :696-697
    out.update(p_m=full_p_mu, p_v=full_p_var, inside_elbo_recon=recon, inside_elbo_kl=kl,
               inside_elbo=inside_elbo, ce_term=ce, KL_term=ce + inside_elbo)             # :709
    return out
