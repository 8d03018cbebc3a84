"""svgp_vae_b200 -- B200-native (sm_100a) implementation of the SVGP hot path of ratschlab/SVGP-VAE.

Public surface (mirrors SVGPVAE_model.py): SVGP, mainSVGP, mnistSVGP, spritesSVGP, gauss_cross_entropy,
_add_diagonal_jitter; plus the batched entry ``mainSVGP.elbo_step`` / ``step.svgp_step`` and the
synthetic ``productSVGP``; prediction-time entries ``precompute_GP_params_SVGPVAE`` (:989-1023),
``predict_from_precomputed`` (:610-635 over all channels) and ``posterior_predict`` (:1048-1050); ``SVIGP_Hensman``
(SVIGP_Hensman_model.py:14-227, the free-form q(u) variant on the same kernels).  All compute runs in libsvgp_b200.so (include/svgp_b200.h); there is no
CPU fallback -- importing is cheap, the first kernel call raises if the library or a B200 is missing.
"""
from .svgp import (SVGP, _add_diagonal_jitter, gauss_cross_entropy, mainSVGP, mnistSVGP, productSVGP,  # noqa: F401
                   reciprocal_no_nan, spritesSVGP)
from .step import elbo_terms, svgp_step  # noqa: F401
from .glue import (GraphedBallStep, aux_data_SVGPVAE_sprites, ball_svgp_terms, build_SVGPVAE_elbo_graph,  # noqa: F401
                   forward_pass_SVGPVAE)
from .svigp import SVIGP_Hensman, forward_pass_deep_SVIGP_Hensman  # noqa: F401
from .graphed import GraphedElboStep  # noqa: F401
from .predict import posterior_predict, precompute_GP_params_SVGPVAE, predict_from_precomputed  # noqa: F401

__version__ = "0.1.0"
