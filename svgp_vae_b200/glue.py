"""PyTorch glue either side of the SVGP hot path: the reference's ``forward_pass_SVGPVAE`` (SVGPVAE_model.py:823-936)
and ``aux_data_SVGPVAE_sprites`` (:1086-1115) with the L-loop :865-898 replaced by ``svgp.elbo_step``.

Encoder / decoder / representation networks are whatever ``torch.nn.Module`` the caller brings (the reference's
Keras nets stay out of scope): ``vae`` needs ``.encode(images) -> (mu, var)`` and ``.decode(z)``, ``repr_NN`` needs
``.repr_nn(images)``, exactly the attributes the reference functions use.  Same arguments, same 16-tuple back.
"""
import torch


def aux_data_SVGPVAE_sprites(data_batch, repr_nn, segment_ids, repeats):
    """:1086-1115 -- per-character mean of the representation vectors (tf.segment_mean), repeated per frame
    (tf.repeat), with the action id in column 0.  Differentiable w.r.t. the representation network."""
    images, action_IDs = data_batch
    character_vectors = repr_nn.repr_nn(images)                                           # :1104
    segment_ids = torch.as_tensor(segment_ids, device=character_vectors.device, dtype=torch.long)
    n_seg = int(segment_ids.max().item()) + 1
    sums = torch.zeros(n_seg, character_vectors.shape[1], dtype=character_vectors.dtype, device=character_vectors.device)
    sums = sums.index_add(0, segment_ids, character_vectors)
    counts = torch.bincount(segment_ids, minlength=n_seg).clamp_min(1).to(character_vectors.dtype)
    character_vectors = sums / counts[:, None]                                            # :1108 segment_mean
    repeats = torch.as_tensor(repeats, device=character_vectors.device, dtype=torch.long)
    character_vectors = torch.repeat_interleave(character_vectors, repeats, dim=0)        # :1111
    ids = torch.as_tensor(action_IDs, device=character_vectors.device).to(character_vectors.dtype)
    return torch.cat([ids[:, None], character_vectors], dim=1)                            # :1114


def forward_pass_SVGPVAE(data_batch, beta, vae, svgp, C_ma, lagrange_mult, alpha, kappa, clipping_qs=False, GECO=False,
                         repr_NN=None, segment_ids=None, repeats=None, bias_analysis=False, epsilon=None, group=None):
    """SVGPVAE_model.py:823-936.  ``epsilon`` (optional, (b, L)) fixes the reparametrisation noise of :901.

    ``group`` shards the batch over a torch.distributed group (see mainSVGP.elbo_step).  Convention when sharded: the
    scalars of the SVGP step (KL_term, inside_elbo, ce_term ...) are GLOBAL and identical on every rank; the returned
    ``elbo`` is this rank's SHARE of the global objective -- local reconstruction term + global terms / world size -- so
    that the sum of the ranks' elbo values is the reference's elbo of the whole batch and back-propagating each rank's
    share gives per-rank partial gradients of the replicated parameters (all-reduce-sum them).  ``recon_loss`` is the
    local share as well; ``C_ma`` / ``lagrange_mult`` (GECO) are computed from the all-reduced reconstruction loss and
    are identical on all ranks."""
    images, aux_data = data_batch
    world = 1 if group is None else torch.distributed.get_world_size(group)
    _, w, h, c = images.shape                                                             # :850 (NHWC like the reference)
    K = float(w * h * c)
    b = float(images.shape[0])

    qnet_mu, qnet_var = vae.encode(images)                                                # :854
    L = float(qnet_mu.shape[1])
    if clipping_qs:
        qnet_var = torch.clamp(qnet_var, 1e-3, 10)                                        # :858-859
    if repr_NN is not None:
        aux_data = aux_data_SVGPVAE_sprites(data_batch=data_batch, repr_nn=repr_NN, segment_ids=segment_ids, repeats=repeats)

    # :865-898 -- every latent channel at once on the B200 path
    res = svgp.elbo_step(aux_data, qnet_mu, qnet_var, clip_pv=bool(repr_NN), group=group)
    p_m, p_v = res["p_m"], res["p_v"]
    inside_elbo_recon, inside_elbo_kl = res["inside_elbo_recon"], res["inside_elbo_kl"]
    inside_elbo, ce_term, KL_term = res["inside_elbo"], res["ce_term"], res["KL_term"]

    if epsilon is None:
        epsilon = torch.randn_like(p_m)                                                   # :901
    latent_samples = p_m + epsilon * torch.sqrt(p_v)                                      # :902
    recon_images_logits = vae.decode(latent_samples)                                      # :905
    recon_images = recon_images_logits

    if GECO:                                                                              # :908-915
        recon_loss = ((images - recon_images_logits) ** 2).mean(dim=(1, 2, 3))
        recon_loss = (recon_loss - kappa ** 2).sum()
        recon_all, b_all = recon_loss.detach().clone(), b
        if group is not None:                              # the moving average and the multiplier see the WHOLE batch
            count = torch.tensor([b], dtype=recon_all.dtype, device=recon_all.device)
            torch.distributed.all_reduce(recon_all, group=group)
            torch.distributed.all_reduce(count, group=group)
            b_all = float(count.item())
        C_ma = alpha * C_ma + (1 - alpha) * recon_all / b_all
        elbo = -KL_term / world + lagrange_mult * (recon_loss / b_all + (C_ma - recon_all / b_all).detach() / world)
        lagrange_mult = lagrange_mult * torch.exp(torch.as_tensor(C_ma))
    else:                                                                                 # :917-925
        recon_loss = ((images - recon_images_logits) ** 2).sum()
        recon_loss = recon_loss / K
        elbo = -recon_loss + (beta / L) * KL_term / world

    if bias_analysis:                                                                     # :928-931
        mean_vectors = [svgp.mean_vector_bias_analysis(aux_data, qnet_mu[:, l], qnet_var[:, l]) for l in range(qnet_mu.shape[1])]
    else:
        mean_vectors = torch.tensor(1.0)
    return (elbo, recon_loss, KL_term, inside_elbo, ce_term, p_m, p_v, qnet_mu, qnet_var, recon_images, inside_elbo_recon,
            inside_elbo_kl, latent_samples, C_ma, lagrange_mult, mean_vectors)


def ball_svgp_terms(svgp_x, svgp_y, qnet_mu, qnet_var):
    """The SVGP part of ``build_SVGPVAE_elbo_graph`` (SVGPVAE_model.py:663-664, 674-697): both latent objects of the
    moving-ball model on the time grid 1..tmax.  qnet_mu / qnet_var (batch, tmax, 2) ->
    dict(full_p_mu, full_p_var (batch, tmax, 2), inside_elbo_recon, inside_elbo_kl, inside_elbo, ce_term, KL_term (batch,),
         p_v_x, p_v_y (batch, tmax, tmax) full posterior covariances, mu_hat_*, A_hat_*)."""
    from .svgp import gauss_cross_entropy
    batch, tmax, _ = qnet_mu.shape
    T = torch.arange(tmax, dtype=qnet_mu.dtype, device=qnet_mu.device) + 1.0            # :663  (1..tmax, not 0..tmax-1)
    batch_T = T.reshape(1, tmax).repeat(batch, 1)                                       # :664
    p_m_x, p_v_x, mu_hat_x, A_hat_x = svgp_x.approximate_posterior_params(batch_T, y=qnet_mu[:, :, 0], noise=qnet_var[:, :, 0])   # :674
    p_m_y, p_v_y, mu_hat_y, A_hat_y = svgp_y.approximate_posterior_params(batch_T, y=qnet_mu[:, :, 1], noise=qnet_var[:, :, 1])   # :676
    rec_x, kl_x = svgp_x.variational_loss(batch_T, qnet_mu[:, :, 0], qnet_var[:, :, 0], mu_hat=mu_hat_x, A_hat=A_hat_x)           # :680
    rec_y, kl_y = svgp_y.variational_loss(batch_T, qnet_mu[:, :, 1], qnet_var[:, :, 1], mu_hat=mu_hat_y, A_hat=A_hat_y)           # :682
    inside_elbo_recon = rec_x + rec_y                                                   # :684
    inside_elbo_kl = kl_x + kl_y
    inside_elbo = inside_elbo_recon - inside_elbo_kl                                    # :686 (no b / N_train factor: per-video GPs)
    full_p_mu = torch.stack([p_m_x, p_m_y], dim=2)                                      # :692
    full_p_var = torch.stack([torch.diagonal(p_v_x, dim1=-2, dim2=-1), torch.diagonal(p_v_y, dim1=-2, dim2=-1)], dim=2)       # :693
    ce_term = -gauss_cross_entropy(full_p_mu, full_p_var, qnet_mu, qnet_var).sum((1, 2))   # :696-697 (note the sign)
    KL_term = ce_term + inside_elbo                                                     # :709
    return dict(full_p_mu=full_p_mu, full_p_var=full_p_var, inside_elbo_recon=inside_elbo_recon, inside_elbo_kl=inside_elbo_kl,
                inside_elbo=inside_elbo, ce_term=ce_term, KL_term=KL_term, p_v_x=p_v_x, p_v_y=p_v_y, mu_hat_x=mu_hat_x,
                A_hat_x=A_hat_x, mu_hat_y=mu_hat_y, A_hat_y=A_hat_y)


def build_SVGPVAE_elbo_graph(vid_batch, beta, svgp_x, svgp_y, clipping_qs=False, encoder=None, decoder=None, epsilon=None):
    """SVGPVAE_model.py:638-716 for the moving-ball model.  The reference builds its MLP encoder / decoder inside
    (VAE_utils.build_MLP_inference_graph / build_MLP_decoder_graph, out of the graft); here they are the caller's modules:
    ``encoder(vid_batch) -> (qnet_mu, qnet_var)`` (batch, tmax, 2) each, ``decoder(latent_samples) -> logits``
    (batch, tmax, px, py).  ``epsilon`` (batch, tmax, 2) fixes the reparametrisation noise of :700.
    Returns the reference's tuple (the trailing ``globals()`` entry is None)."""
    qnet_mu, qnet_var = encoder(vid_batch)                                              # :667
    if clipping_qs:
        qnet_var = torch.clamp(qnet_var, 1e-6, 1e3)                                     # :671
    t = ball_svgp_terms(svgp_x, svgp_y, qnet_mu, qnet_var)
    full_p_mu, full_p_var = t["full_p_mu"], t["full_p_var"]
    if epsilon is None:
        epsilon = torch.randn_like(full_p_mu)                                           # :700
    latent_samples = full_p_mu + epsilon * torch.sqrt(torch.clamp(full_p_var, 1e-4, 1000))   # :701
    logits = decoder(latent_samples)                                                    # :704
    pred_vid = torch.sigmoid(logits)
    recon_term = -torch.nn.functional.binary_cross_entropy_with_logits(logits, vid_batch, reduction="none").sum((1, 2, 3))   # :706-707
    CPH_elbo = recon_term + beta * t["KL_term"]                                         # :710
    return (CPH_elbo, recon_term, t["KL_term"], t["inside_elbo"], t["ce_term"], full_p_mu, full_p_var, qnet_mu, qnet_var, pred_vid,
            svgp_x.l_GP, svgp_y.l_GP, t["inside_elbo_recon"], t["inside_elbo_kl"], svgp_x.inducing_index_points,
            svgp_y.inducing_index_points, t["p_v_x"].mean(0), t["p_v_y"].mean(0), None)


class GraphedBallStep:
    """``ball_svgp_terms`` captured as CUDA graphs (forward and backward; the encoder / decoder run between them as ordinary
    PyTorch).  The ball configuration is ~500 launches of a few microseconds: launch-bound, so graphs are the lever.
    Call with (qnet_mu, qnet_var) of the captured shape -> (full_p_mu, full_p_var, KL_term)."""

    def __init__(self, svgp_x, svgp_y, batch, tmax, device="cuda", dtype=torch.float32):
        self.svgp_x, self.svgp_y = svgp_x, svgp_y
        mu = torch.zeros(batch, tmax, 2, device=device, dtype=dtype, requires_grad=True)
        var = torch.ones(batch, tmax, 2, device=device, dtype=dtype, requires_grad=True)

        class _M(torch.nn.Module):
            def __init__(self, sx, sy):
                super().__init__()
                self.sx, self.sy = sx, sy

            def forward(self, m, v):
                t = ball_svgp_terms(self.sx, self.sy, m, v)
                return t["full_p_mu"], t["full_p_var"], t["KL_term"]
        from . import ops
        ops.pd_flag(mu.device)                      # static device word the captured Cholesky calls report a bad pivot into
        self._fn = torch.cuda.make_graphed_callables(_M(svgp_x, svgp_y), (mu, var), num_warmup_iters=3, allow_unused_input=True)
        self._device = mu.device

    def __call__(self, qnet_mu, qnet_var):
        return self._fn(qnet_mu, qnet_var)

    def check(self):
        """Raise ops.NotPositiveDefinite if any Cholesky factorisation of a replayed step met a non-positive pivot."""
        from . import ops
        ops.pd_flag_check(self._device)
