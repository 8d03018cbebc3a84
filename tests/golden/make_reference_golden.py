"""Golden vectors from the REFERENCE SOURCE ITSELF, executed here under a TensorFlow-API shim.

    python tests/golden/make_reference_golden.py          # needs /root/reference (this container only)

TensorFlow 1.15 / TFP 0.8 cannot be installed in this image, so the reference cannot run as shipped.
Its SVGP path is pure tensor algebra through ~40 tf.* symbols, though: tests/golden/tf_shim.py provides
those symbols on torch float64 (eager), the unmodified /root/reference/SVGPVAE_model.py and utils.py are
imported on top of it, and the reference's own classes / functions are called:

  mnist*    forward_pass_SVGPVAE (SVGPVAE_model.py:823-936) with a stub VAE whose encoder returns the test's
            (qnet_mu, qnet_var) -- i.e. the L-loop :868-878, the glue :880-898, mnistSVGP.kernel_matrix :427-476,
            mainSVGP.approximate_posterior_params :303-343 and variational_loss :220-301 all run as written
  sprites*  spritesSVGP (:487-635) per channel exactly as :868-878 calls it, then :880-898 (with the p_v clip
            :891-892) re-typed below because forward_pass_SVGPVAE only clips when a representation network is
            attached (:846-848), which is outside the SVGP path
  ball      SVGP (:17-171) as build_SVGPVAE_elbo_graph calls it (:674-683) + the glue :685-697, re-typed below
            because that function builds its own randomly initialised encoder (:667)

Gradients: torch autograd differentiates through the reference code (stand-in for tf.gradients) for the
objective J = KL_term + <g_m, p_m> + <g_v, p_v> of tests/refs.py.  Output: tests/golden/reference_golden.npz,
checked by tests/test_oracle.py::test_oracle_matches_reference_source and by the GPU parity tests.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REFERENCE = os.environ.get("SVGP_REFERENCE_DIR", "/root/reference")

import tf_shim  # noqa: E402

tf = tf_shim.install()
sys.path.insert(0, REFERENCE)
import SVGPVAE_model as ref  # noqa: E402  (the reference, unmodified)
from utils import gauss_cross_entropy  # noqa: E402  (reference utils.py:483-504)

import refs  # noqa: E402
from svgp_vae_b200 import configs  # noqa: E402

F64 = torch.float64


class StubVAE:
    """forward_pass_SVGPVAE only needs .dtype, .encode (-> qnet_mu, qnet_var) and .decode."""
    dtype = np.float64

    def __init__(self, mu, var):
        self.mu, self.var = mu, var

    def encode(self, images):
        return self.mu, self.var

    def decode(self, z):
        return torch.zeros(z.shape[0], 28, 28, 1, dtype=F64)


def _leaf(t):
    return torch.as_tensor(np.asarray(t)).to(F64).clone().requires_grad_(True)


def _pack(out, name, res, J, grads, gnames):
    out[name + "/p_m"] = res["p_m"].detach().numpy()
    out[name + "/p_v"] = res["p_v"].detach().numpy()
    out[name + "/scalars"] = np.array([float(res[k]) for k in ("inside_elbo_recon", "inside_elbo_kl", "ce_term", "KL_term")] + [float(J)])
    for g, n in zip(grads, gnames):
        out[name + "/grad_" + n] = (torch.zeros(1) if g is None else g).detach().numpy()


def mnist_case(out, name, cfg):
    ctor = dict(cfg["ctor"])
    svgp = ref.mnistSVGP(name="ref", **ctor)
    aux = cfg["aux"].to(F64)
    mu, var = _leaf(cfg["y"]), _leaf(cfg["noise"])
    images = torch.zeros(aux.shape[0], 28, 28, 1, dtype=F64)
    r = ref.forward_pass_SVGPVAE((images, aux), beta=1.0, vae=StubVAE(mu, var), svgp=svgp, C_ma=0.0, lagrange_mult=1.0,
                                 alpha=0.99, kappa=0.02, clipping_qs=False, GECO=False)
    elbo, recon_loss, KL_term, inside_elbo, ce_term, p_m, p_v, _, _, _, rec, kl = r[:12]
    res = dict(p_m=p_m, p_v=p_v, inside_elbo_recon=rec, inside_elbo_kl=kl, ce_term=ce_term, KL_term=KL_term)
    gm, gv = refs.upstream(tuple(mu.shape))
    J = KL_term + (gm * p_m).sum() + (gv * p_v).sum()
    leaves = [mu, var, svgp.inducing_index_points, svgp.object_vectors, svgp.amplitude, svgp.l_GP]
    grads = torch.autograd.grad(J, leaves, allow_unused=True)
    _pack(out, name, res, J, grads, ["y", "noise", "Z", "table", "amplitude", "length"])
    # the kernel builder on its own (:427-476)
    out[name + "/K_nm"] = svgp.kernel_matrix(aux, svgp.inducing_index_points, x_inducing=False).detach().numpy()
    out[name + "/K_mm"] = svgp.kernel_matrix(svgp.inducing_index_points, svgp.inducing_index_points).detach().numpy()
    out[name + "/K_nn_diag"] = svgp.kernel_matrix(aux, aux, x_inducing=False, y_inducing=False, diag_only=True).detach().numpy()
    # one channel of the per-channel API (:303-343, :220-301)
    m, B, mu_hat, A_hat = svgp.approximate_posterior_params(aux, aux, mu[:, 0], var[:, 0])
    L3, KL = svgp.variational_loss(x=aux, y=mu[:, 0], noise=var[:, 0], mu_hat=mu_hat, A_hat=A_hat)
    out[name + "/ch0_mu_hat"], out[name + "/ch0_A_hat"] = mu_hat.detach().numpy(), A_hat.detach().numpy()
    out[name + "/ch0_L3_KL"] = np.array([float(L3), float(KL)])
    out[name + "/ch0_bias_mean"] = svgp.mean_vector_bias_analysis(aux, mu[:, 0], var[:, 0]).detach().numpy()
    # conditional generation: test index points against training encodings, the L-loop of
    # bacthing_predict_SVGPVAE_rotated_mnist (:1048-1050); test rows = the first 48 rows with the view angle shifted
    with torch.no_grad():
        test_aux = aux[:48].clone()
        test_aux[:, 1] = test_aux[:, 1] + 0.3
        pm, pv = [], []
        for l in range(mu.shape[1]):
            m_l, v_l, _, _ = svgp.approximate_posterior_params(test_aux, aux, mu[:, l], var[:, l])
            pm.append(m_l); pv.append(v_l)
        out[name + "/cgen_p_m"], out[name + "/cgen_p_v"] = tf.stack(pm, axis=1).numpy(), tf.stack(pv, axis=1).numpy()


def sprites_case(out, name, cfg, clip=True):
    ctor = dict(cfg["ctor"])
    svgp = ref.spritesSVGP(name="ref", **ctor)
    aux = cfg["aux"].to(F64)
    qnet_mu, qnet_var = _leaf(cfg["y"]), _leaf(cfg["noise"])
    b = tf.cast(tf.shape(aux)[0], dtype=np.float64)
    # ---- SVGPVAE_model.py:865-898 (re-typed; see the module docstring) ----
    inside_elbo_recon, inside_elbo_kl = [], []
    p_m, p_v = [], []
    for l in range(qnet_mu.get_shape()[1]):
        p_m_l, p_v_l, mu_hat_l, A_hat_l = svgp.approximate_posterior_params(aux, aux, qnet_mu[:, l], qnet_var[:, l])
        rec_l, kl_l = svgp.variational_loss(x=aux, y=qnet_mu[:, l], noise=qnet_var[:, l], mu_hat=mu_hat_l, A_hat=A_hat_l)
        inside_elbo_recon.append(rec_l)
        inside_elbo_kl.append(kl_l)
        p_m.append(p_m_l)
        p_v.append(p_v_l)
    inside_elbo_recon = tf.reduce_sum(inside_elbo_recon)
    inside_elbo_kl = tf.reduce_sum(inside_elbo_kl)
    inside_elbo = inside_elbo_recon - (b / svgp.N_train) * inside_elbo_kl
    p_m = tf.stack(p_m, axis=1)
    p_v = tf.stack(p_v, axis=1)
    if clip:
        p_v = tf.clip_by_value(p_v, 1e-4, 100)
    ce_term = tf.reduce_sum(gauss_cross_entropy(p_m, p_v, qnet_mu, qnet_var))
    KL_term = -ce_term + inside_elbo
    # -----------------------------------------------------------------------
    res = dict(p_m=p_m, p_v=p_v, inside_elbo_recon=inside_elbo_recon, inside_elbo_kl=inside_elbo_kl, ce_term=ce_term,
               KL_term=KL_term)
    gm, gv = refs.upstream(tuple(qnet_mu.shape))
    J = KL_term + (gm * p_m).sum() + (gv * p_v).sum()
    leaves = [qnet_mu, qnet_var, svgp.inducing_index_points, svgp.GPLVM_action]
    names = ["y", "noise", "Z", "table"]
    if ctor.get("K_SE"):
        leaves += [svgp.sigma_action, svgp.l_action, svgp.sigma_character, svgp.l_character]
        names += ["sigma_action", "l_action", "sigma_character", "l_character"]
    grads = torch.autograd.grad(J, leaves, allow_unused=True)
    _pack(out, name, res, J, grads, names)
    out[name + "/K_nm"] = svgp.kernel_matrix(aux, svgp.inducing_index_points, x_inducing=False).detach().numpy()
    out[name + "/K_nn_diag"] = svgp.kernel_matrix(aux, aux, x_inducing=False, y_inducing=False, diag_only=True).detach().numpy()
    # prediction-time entry (:610-635) on channel 0, fed the way precompute_GP_params_SVGPVAE feeds it (:1004-1019)
    with torch.no_grad():
        K_mm = svgp.kernel_matrix(svgp.inducing_index_points, svgp.inducing_index_points)
        K_nm = svgp.kernel_matrix(aux, svgp.inducing_index_points, x_inducing=False)
        prec = tf.math.reciprocal_no_nan(qnet_var[:, 0])
        sigma_l = K_mm + tf.matmul(tf.transpose(K_nm, perm=[1, 0]), tf.multiply(K_nm, prec[:, tf.newaxis]))
        sigma_l_inv = tf.linalg.inv(sigma_l)                                   # :1014 -- no jitter at prediction time
        mean_term = tf.linalg.matvec(sigma_l_inv, tf.linalg.matvec(tf.transpose(K_nm, perm=[1, 0]), prec * qnet_mu[:, 0]))
        mean_vec, Bdiag = svgp.approximate_posterior_params_precomputed_GP_posterior_params(aux, mean_term, sigma_l_inv)
    out[name + "/pred_mean_term"], out[name + "/pred_sigma_term"] = mean_term.numpy(), sigma_l_inv.numpy()
    out[name + "/pred_mean"], out[name + "/pred_B"] = mean_vec.numpy(), Bdiag.numpy()
    # the reference's own precompute over all channels (:989-1023) and the per-channel predictions from it (:1165-1168)
    with torch.no_grad():
        mean_terms, inv_sigmas = ref.precompute_GP_params_SVGPVAE(qnet_mu, qnet_var, aux, svgp)
        pm, pv = [], []
        for l in range(qnet_mu.shape[1]):
            m_l, v_l = svgp.approximate_posterior_params_precomputed_GP_posterior_params(aux[:100], mean_terms[l], inv_sigmas[l])
            pm.append(m_l); pv.append(v_l)
    out[name + "/precomp_mean_terms"], out[name + "/precomp_inv_sigma"] = mean_terms.numpy(), inv_sigmas.numpy()
    out[name + "/precomp_p_m"], out[name + "/precomp_p_v"] = tf.stack(pm, axis=1).numpy(), tf.stack(pv, axis=1).numpy()


def ball_case(out, cfg):
    ctor = dict(cfg["ctor"])
    svgp_x, svgp_y = ref.SVGP(name="x", **ctor), ref.SVGP(name="y", **ctor)
    qnet_mu, qnet_var = _leaf(cfg["y"]), _leaf(cfg["noise"])
    batch, tmax = qnet_mu.shape[0], qnet_mu.shape[1]
    # ---- SVGPVAE_model.py:663-664, 674-697, 709 (re-typed; see the module docstring) ----
    T = tf.range(tmax, dtype=np.float64) + 1.0
    batch_T = tf.tile(tf.expand_dims(T, 0), (batch, 1))
    p_m_x, p_v_x, mu_hat_x, A_hat_x = svgp_x.approximate_posterior_params(index_points=batch_T, y=qnet_mu[:, :, 0], noise=qnet_var[:, :, 0])
    p_m_y, p_v_y, mu_hat_y, A_hat_y = svgp_y.approximate_posterior_params(index_points=batch_T, y=qnet_mu[:, :, 1], noise=qnet_var[:, :, 1])
    recon_x, kl_x = svgp_x.variational_loss(x=batch_T, y=qnet_mu[:, :, 0], noise=qnet_var[:, :, 0], mu_hat=mu_hat_x, A_hat=A_hat_x)
    recon_y, kl_y = svgp_y.variational_loss(x=batch_T, y=qnet_mu[:, :, 1], noise=qnet_var[:, :, 1], mu_hat=mu_hat_y, A_hat=A_hat_y)
    inside_elbo_recon = recon_x + recon_y
    inside_elbo_kl = kl_x + kl_y
    inside_elbo = inside_elbo_recon - inside_elbo_kl
    full_p_mu = tf.stack([p_m_x, p_m_y], axis=2)
    full_p_var = tf.stack([tf.linalg.diag_part(p_v_x), tf.linalg.diag_part(p_v_y)], axis=2)
    ce_term = gauss_cross_entropy(full_p_mu, full_p_var, qnet_mu, qnet_var)
    ce_term = -tf.reduce_sum(ce_term, (1, 2))
    KL_term = ce_term + inside_elbo
    # -------------------------------------------------------------------------------------
    gm, gv = refs.upstream(tuple(qnet_mu.shape))
    J = KL_term.sum() + (gm * full_p_mu).sum() + (gv * full_p_var).sum()
    gy, gn = torch.autograd.grad(J, [qnet_mu, qnet_var])
    out["ball/p_m"], out["ball/p_v"] = full_p_mu.detach().numpy(), full_p_var.detach().numpy()
    out["ball/B_x"] = p_v_x.detach().numpy()
    out["ball/mu_hat_x"], out["ball/A_hat_x"] = mu_hat_x.detach().numpy(), A_hat_x.detach().numpy()
    out["ball/KL_term"], out["ball/recon"], out["ball/kl"] = KL_term.detach().numpy(), inside_elbo_recon.detach().numpy(), inside_elbo_kl.detach().numpy()
    out["ball/J"] = np.array([float(J)])
    out["ball/grad_y"], out["ball/grad_noise"] = gy.numpy(), gn.numpy()


class GlueVAE(StubVAE):
    """Deterministic stand-in decoder so that elbo / recon_loss / GECO state depend on the latent samples."""

    def decode(self, z):
        base = torch.linspace(-1.0, 1.0, 28 * 28, dtype=F64).reshape(1, 28, 28, 1)
        return torch.tanh(z.sum(1)).reshape(-1, 1, 1, 1) * base + 0.1 * z[:, :1].reshape(-1, 1, 1, 1)


def glue_images(b):
    i = torch.arange(b, dtype=F64).reshape(-1, 1, 1, 1)
    base = torch.linspace(0.0, 1.0, 28 * 28, dtype=F64).reshape(1, 28, 28, 1)
    return torch.sin(0.1 * i + 3.0 * base)


def glue_cases(out, fx):
    """The rest of forward_pass_SVGPVAE (:900-936: sampling, decoder, beta-ELBO and GECO) and aux_data_SVGPVAE_sprites
    (:1086-1115), for svgp_vae_b200/glue.py."""
    cfg = configs.mnist_inputs(fx, L=4)
    aux = cfg["aux"].to(F64)
    images = glue_images(aux.shape[0])
    for geco in (False, True):
        svgp = ref.mnistSVGP(name="ref", **dict(cfg["ctor"]))
        mu, var = _leaf(cfg["y"]), _leaf(cfg["noise"])
        r = ref.forward_pass_SVGPVAE((images, aux), beta=0.7, vae=GlueVAE(mu, var), svgp=svgp, C_ma=0.3, lagrange_mult=1.5,
                                     alpha=0.99, kappa=0.02, clipping_qs=True, GECO=geco)
        tag = "glue_geco" if geco else "glue_beta"
        out[tag + "/elbo"] = np.array([float(r[0])])
        out[tag + "/recon_loss"] = np.array([float(r[1])])
        out[tag + "/latent_samples"] = r[12].detach().numpy()
        out[tag + "/C_ma"] = np.array([float(r[13])])
        out[tag + "/lagrange_mult"] = np.array([float(r[14])])
        g = torch.autograd.grad(r[0], [mu, var, svgp.inducing_index_points])
        out[tag + "/grad_y"], out[tag + "/grad_noise"], out[tag + "/grad_Z"] = (t.numpy() for t in g)
    out["glue/epsilon"] = tf.random.normal(shape=(aux.shape[0], 4)).numpy()

    class Repr:
        def repr_nn(self, images):
            return images.reshape(images.shape[0], -1)[:, :16] * 2.0 + 0.5
    b = 12
    imgs = glue_images(b)
    action_ids = torch.tensor([3, 7, 1, 0, 5, 5, 2, 71, 9, 4, 6, 8])
    seg, rep = [0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2], [5, 3, 4]
    out["sprites_aux/aux"] = ref.aux_data_SVGPVAE_sprites((imgs, action_ids), Repr(), seg, rep).numpy()


def titsias_cases(out, fx):
    """The L_2 (Titsias) branch of both classes: SVGPVAE_model.py:246-259 through forward_pass_SVGPVAE, :89-101 direct."""
    cfg = configs.mnist_inputs(fx, L=2, b=64)
    ctor = dict(cfg["ctor"], titsias=True)
    svgp = ref.mnistSVGP(name="ref", **ctor)
    aux = cfg["aux"].to(F64)
    mu, var = _leaf(cfg["y"]), _leaf(cfg["noise"])
    images = torch.zeros(aux.shape[0], 28, 28, 1, dtype=F64)
    r = ref.forward_pass_SVGPVAE((images, aux), beta=1.0, vae=StubVAE(mu, var), svgp=svgp, C_ma=0.0, lagrange_mult=1.0,
                                 alpha=0.99, kappa=0.02, clipping_qs=False, GECO=False)
    elbo, recon_loss, KL_term, inside_elbo, ce_term, p_m, p_v, _, _, _, rec, kl = r[:12]
    res = dict(p_m=p_m, p_v=p_v, inside_elbo_recon=rec, inside_elbo_kl=kl, ce_term=ce_term, KL_term=KL_term)
    gm, gv = refs.upstream(tuple(mu.shape))
    J = KL_term + (gm * p_m).sum() + (gv * p_v).sum()
    grads = torch.autograd.grad(J, [mu, var, svgp.inducing_index_points, svgp.object_vectors, svgp.amplitude, svgp.l_GP], allow_unused=True)
    _pack(out, "mnist_titsias", res, J, grads, ["y", "noise", "Z", "table", "amplitude", "length"])
    cfgb = configs.ball_inputs()
    sb = ref.SVGP(name="x", **dict(cfgb["ctor"], titsias=True))
    y, nz = _leaf(cfgb["y"][:, :, 0]), _leaf(cfgb["noise"][:, :, 0])
    batch_T = tf.tile(tf.expand_dims(tf.range(30, dtype=np.float64) + 1.0, 0), (y.shape[0], 1))
    _, _, mu_hat, A_hat = sb.approximate_posterior_params(index_points=batch_T, y=y, noise=nz)
    L2, zero = sb.variational_loss(x=batch_T, y=y, noise=nz, mu_hat=mu_hat, A_hat=A_hat)
    gy, gn = torch.autograd.grad(L2.sum(), [y, nz])
    out["ball_titsias/L2"] = L2.detach().numpy()
    out["ball_titsias/grad_y"], out["ball_titsias/grad_noise"] = gy.numpy(), gn.numpy()


def main():
    fx = os.path.join(HERE, "mnist_aux.npz")
    out = {}
    mnist_case(out, "mnist", configs.mnist_inputs(fx, L=4))
    mnist_case(out, "mnist_norm", configs.mnist_inputs(fx, L=4, normalize=True))
    mnist_case(out, "mnist_train_last", configs.mnist_inputs(fx, L=2, b=210, rows="train", batch_index=15))
    sprites_case(out, "sprites72", configs.sprites_inputs(M=72, L=4))
    sprites_case(out, "sprites72_raw", configs.sprites_inputs(M=72, L=4, normalize=False))
    sprites_case(out, "sprites72_se", configs.sprites_inputs(M=72, L=3, K_SE=True))
    ball_case(out, configs.ball_inputs())
    titsias_cases(out, fx)
    glue_cases(out, fx)
    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays;", os.path.getsize(path), "bytes")
    for k in ("mnist/scalars", "mnist_norm/scalars", "sprites72/scalars", "sprites72_se/scalars", "ball/J"):
        print(k, out[k])


if __name__ == "__main__":
    main()
