"""CPU oracle for the SVGP hot path of ratschlab/SVGP-VAE -- TEST INFRASTRUCTURE ONLY.

This package is a float64 CPU restatement (torch-CPU, so reverse-mode gradients come
from torch autograd) of the reference's Hensman-style sparse variational GP:

  * ``tfp_kernels``       -- tensorflow-probability 0.8.0 ``psd_kernels`` (third-party,
                             un-vendored; requirements.txt:9) restated from its public
                             definition: ExponentiatedQuadratic, ExpSinSquared, Linear.
  * ``svgp_literal``      -- op-for-op restatement of SVGPVAE_model.py:13-635 (SVGP,
                             mainSVGP, mnistSVGP, spritesSVGP), utils.py:483-504
                             (gauss_cross_entropy) and the two call sites
                             SVGPVAE_model.py:865-898 / :674-697.  Keeps the (b,m,m)
                             lambda_mat tensor, the explicit inverses and every quirk.
  * ``svgp_streamlined``  -- the same mathematics in the jitter-exact collapsed form the
                             CUDA path computes; must agree with the
                             literal form to ~1e-12 (tests/test_oracle.py).

PINNING.  The reference ships no tests, golden vectors or known-answer fixtures for this path
(SURVEY section 4 / 8c), and TensorFlow 1.15 + TFP 0.8 cannot be installed in this image
(python 3.12, no wheel, no network).  The oracle is therefore pinned to the reference SOURCE
rather than to a TensorFlow run: the unmodified /root/reference/SVGPVAE_model.py, utils.py and
SVIGP_Hensman_model.py are executed under tests/golden/tf_shim.py (a ~40-symbol TF-1.15 API shim
on torch float64) by tests/golden/make_reference_golden.py / make_svigp_golden.py, and
tests/test_oracle.py::test_oracle_matches_reference_source checks this package against those
outputs (values 1e-10, gradients 1e-8).  PARITY UNPINNED with respect to TensorFlow / TFP
themselves: the arithmetic inside tf.linalg.* and tfp.math.psd_kernels is restated from the
documented definitions, no number here was produced by TensorFlow.  Further anchors: (i) literal
== streamlined, (ii) an independent implementation of the kernel formulas (scikit-learn),
(iii) the exact-GP limit (m = b, Z = X) against GPVAE_Pearce_model.py:49-84, (iv) the
survey-session numbers of SURVEY App. B.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this package.  The product package
``svgp_vae_b200`` never does: it fails loudly when its CUDA library is missing.
"""
