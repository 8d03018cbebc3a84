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
