"""SVIGP_Hensman: the free-form q(u) = N(m_l, A_l A_l^T) sparse GP of Hensman et al. (2013), the reference's
``SVIGP_Hensman_model.py:14-227``.  Same constructor, methods, argument order and return tuples;
the kernel is the rotated-MNIST ExpSinSquared(view) x Linear(object) product of ``mnistSVGP`` (:78-127 restates
SVGPVAE_model.py:427-476), so K1 is shared.  Everything runs on the library primitives of ops.py:

  K_mm, K_nm, diag K_nn          svgp_kernel_fwd / svgp_kernel_diag_fwd                     :150-156
  inv(K_mm + jI), log-dets       svgp_chol_f64 / svgp_trinv_f64 (ops.spd_inverse_logdet)     :151, :168-173
  k_i^T Kinv k_i                 svgp_rowquad                                                :182
  sum_i k_i^T Kinv S Kinv k_i    <Kinv S Kinv, sum_i k_i k_i^T> with ONE unit-weight svgp_syrk  (the reference forms
                                 the (b, m, m) tensor :184-192; the SYRK is channel-independent) :184-192

``variational_loss_all`` is the batched entry (all L channels in one call); the per-channel
``variational_loss(x, z, lat_channel)`` is an L = 1 slice of the same computation.
"""
import numpy as np
import torch

from . import ops
from ._lib import SVGP_K_COSINE, SVGP_K_EXPSIN, SVGP_K_LINEAR
from .svgp import _KernelBase, _add_diagonal_jitter, _as_param_or_buffer, _torch_dtype


class SVIGP_Hensman(_KernelBase):
    def __init__(self, fixed_inducing_points, initial_inducing_points, name, jitter, N_train, dtype, L, fixed_gp_params,
                 object_vectors_init, K_obj_normalize=False):
        super().__init__()
        self.dtype = _torch_dtype(dtype)
        self.jitter = jitter
        self.nr_inducing = len(initial_inducing_points)
        self.N_train = N_train
        self.L = L
        self.K_obj_normalize = K_obj_normalize
        self._ip_name = _as_param_or_buffer(self, "Sparse_GP_inducing_points_{}".format(name), initial_inducing_points,
                                            self.dtype, not fixed_inducing_points)                       # :40-45
        self._l_name = _as_param_or_buffer(self, "GP_length_scale_{}".format(name), 1.0, self.dtype, not fixed_gp_params)
        self._amp_name = _as_param_or_buffer(self, "GP_amplitude_{}".format(name), 1.0, self.dtype, not fixed_gp_params)
        if object_vectors_init is not None:                                                              # :58-64
            self._ov_name = _as_param_or_buffer(self, "GP_object_vectors_{}".format(name), object_vectors_init, self.dtype, True)
        else:
            self._ov_name = None
        m = self.nr_inducing
        # (inner) variational parameters, one (m,) mean and one (m, m) scale per latent channel :66-74; stored stacked
        # (the reference keeps python lists of per-channel Variables named GP_var_params_mu_{l+1} / GP_var_params_A_{l+1})
        self.GP_var_params_mu = torch.nn.Parameter(torch.zeros(L, m, dtype=self.dtype))
        self.GP_var_params_A = torch.nn.Parameter(torch.eye(m, dtype=self.dtype).repeat(L, 1, 1))
        self.Hensman_likelihood_noise = torch.nn.Parameter(torch.tensor(0.1, dtype=self.dtype))          # :76

    inducing_index_points = property(lambda self: getattr(self, self._ip_name))
    l_GP = property(lambda self: getattr(self, self._l_name))
    amplitude = property(lambda self: getattr(self, self._amp_name))
    object_vectors = property(lambda self: None if self._ov_name is None else getattr(self, self._ov_name))
    noise = property(lambda self: self.Hensman_likelihood_noise)

    @property
    def variational_inducing_observations_loc(self):
        return [self.GP_var_params_mu[l] for l in range(self.L)]

    @property
    def variational_inducing_observations_scale(self):
        return [self.GP_var_params_A[l] for l in range(self.L)]

    @property
    def variational_inducing_observations_cov_mat(self):                                                 # :72-73
        return [self.GP_var_params_A[l] @ self.GP_var_params_A[l].t() for l in range(self.L)]

    # ---- kernel (same product kernel as mnistSVGP) ----------------------------------------------------------------
    def _spec(self):
        d_obj = self.inducing_index_points.shape[1] - 2
        return (SVGP_K_EXPSIN, 1, SVGP_K_COSINE if self.K_obj_normalize else SVGP_K_LINEAR, d_obj)

    def _hyp(self):
        one = torch.ones((), dtype=self.dtype, device=self.l_GP.device)
        return torch.stack([self.amplitude, self.l_GP, one, one])

    def _features(self, x, inducing):
        if inducing or self.object_vectors is None:                                                      # :92-103
            return x[:, 1:]
        obj = ops.gather_rows(self.object_vectors, x[:, 0].long())
        return torch.cat([x[:, 1:2].to(obj.dtype), obj], dim=1)

    def variable_summary(self):
        return self.l_GP, self.amplitude, self.object_vectors, self.inducing_index_points

    # ---- L_H ------------------------------------------------------------------------------------------------------
    def _loss_channels(self, x, channels):
        """(L_3 sum terms (c,), KL terms (c,), mean vectors (b, c)) for the listed latent channels."""
        hyp, spec = self._hyp(), self._spec()
        m = float(self.nr_inducing)
        Fz = self._features(self.inducing_index_points, True)
        Fx = self._features(x, False)
        K_mm = ops.kernel_matrix(Fz, Fz, hyp, spec).double()
        K_mm_inv, ldK, _ = ops.spd_inverse_logdet(_add_diagonal_jitter(K_mm, self.jitter).unsqueeze(0))  # :151, :168
        K_nn = ops.kernel_diag(Fx, Fx, hyp, spec).double()                                               # :153
        K_nm = ops.kernel_matrix(Fx, Fz, hyp, spec)                                                      # :155
        idx = torch.as_tensor(channels, device=K_nm.device)
        mu = self.GP_var_params_mu.double()[idx]                                                         # (c, m)
        Ach = self.GP_var_params_A.double()[idx]
        S = ops.bmm64(Ach, Ach, False, True)                                                             # A A^T  :72-73
        a = ops.bmm64(mu.unsqueeze(0), K_mm_inv)[0]                                                      # Kinv m (Kinv symmetric)
        mean_vector = ops.k_matmul(K_nm, a).double()                                                     # :161-162  (b, c)
        ld_S = ops.spd_logdet(_add_diagonal_jitter(S, self.jitter))                                      # :169-173
        KL = 0.5 * (ldK[0] - ld_S - m + (K_mm_inv * S).sum((-1, -2)) + (mu * a).sum(-1))                 # :175-178
        precision = 1.0 / self.noise.double()                                                            # :181
        K_tilde = precision * (K_nn - ops.rowquad(K_nm, K_mm_inv)[:, 0].double())                        # :183
        # sum_i tr(S Kinv k_i k_i^T Kinv) = <Kinv S Kinv, sum_i k_i k_i^T>: one unit-weight SYRK for all channels
        G = ops.syrk(K_nm, torch.ones(K_nm.shape[0], 1, dtype=K_nm.dtype, device=K_nm.device))            # (1, m, m)
        W = ops.bmm64(ops.bmm64(K_mm_inv, S), K_mm_inv)                                                  # (c, m, m)
        trace_sum = precision * (W * G).sum((-1, -2))                                                    # :185-195
        L3 = -0.5 * (K_tilde.sum() + trace_sum)                                                          # :198
        dt = self.dtype
        return L3.to(dt), KL.to(dt), mean_vector.to(dt)

    def variational_loss(self, x, z, lat_channel):
        """:135-200 -> (L_3_sum_term, KL_term, mean_vector (b,)) for one latent channel (z is unused, as in the reference)."""
        L3, KL, mean = self._loss_channels(x, [int(lat_channel)])
        return L3[0], KL[0], mean[:, 0]

    def variational_loss_all(self, x):
        """All L channels in one call: (L_3 (L,), KL (L,), mean_vectors (b, L)) -- the loop of
        forward_pass_deep_SVIGP_Hensman :246-252 as one batched computation."""
        return self._loss_channels(x, list(range(self.L)))

    def approximate_posterior_params(self, index_points_test, lat_channel):
        """:202-227 -> (mean_vector (x,), B).  As in the reference, B = K_xx - A (K_mm - S) A^T subtracts an (x, x)
        matrix from the (x,) vector of prior variances by broadcasting, i.e. B[i, j] = K_xx[j] - Q[i, j] (:225)."""
        hyp, spec = self._hyp(), self._spec()
        Fz = self._features(self.inducing_index_points, True)
        Fx = self._features(index_points_test, False)
        K_mm = ops.kernel_matrix(Fz, Fz, hyp, spec).double()
        K_mm_inv, _, _ = ops.spd_inverse_logdet(_add_diagonal_jitter(K_mm, self.jitter).unsqueeze(0))
        K_xx = ops.kernel_diag(Fx, Fx, hyp, spec).double()
        K_xm = ops.kernel_matrix(Fx, Fz, hyp, spec).double()
        l = int(lat_channel)
        mu = self.GP_var_params_mu[l].double()
        Al = self.GP_var_params_A[l].double()
        S = Al @ Al.t()
        A = ops.bmm64(K_xm.unsqueeze(0), K_mm_inv)                                                       # (1, x, m)  :221
        mean_vector = (A[0] @ mu)
        mid = (K_mm - S).unsqueeze(0)
        Q = ops.bmm64(ops.bmm64(A, mid), A, False, True)[0]                                              # (x, x)
        B = K_xx - Q                                                                                     # :225 (broadcast)
        return mean_vector.to(self.dtype), B.to(self.dtype)


def forward_pass_deep_SVIGP_Hensman(data_batch, vae, svgp):
    """SVIGP_Hensman_model.py:230-283 with the L-loop replaced by ``variational_loss_all``.  ``vae`` needs ``decode``
    and ``dtype``; returns the reference's tuple."""
    images, aux_data = data_batch
    _, w, h, c = images.shape
    K = float(w * h * c)
    b = float(images.shape[0])
    recon_l, kl_l, mean_vectors = svgp.variational_loss_all(aux_data[:, 1:])                             # :247 passes aux_data[:, 1:]
    inside_elbo_recon, inside_elbo_kl = recon_l.sum(), kl_l.sum()
    inside_elbo = inside_elbo_recon - (b / svgp.N_train) * inside_elbo_kl                                # :257
    KL_term = inside_elbo
    recon_images = vae.decode(mean_vectors)
    recon_loss = ((images - recon_images) ** 2).sum()                                                    # :267
    elbo = (-b * K * torch.log(svgp.noise) - 0.5 * b * K * float(np.log(2 * np.pi))
            - (0.5 * svgp.noise ** (-2)) * recon_loss + inside_elbo)                                     # :276-277
    recon_loss = recon_loss / K
    return elbo, recon_loss, KL_term, inside_elbo, recon_images, inside_elbo_recon, inside_elbo_kl, mean_vectors
