"""End-to-end parity on the B200: product (CUDA path, through the C ABI) vs the float64 literal oracle.

Tolerance: 1e-4 relative to max|oracle tensor| (BASELINE.json north_star), for posterior means /
variances, the ELBO scalars -- including the cancelling combination KL_term -- and all gradients.
"""
import os

import pytest
import torch

import refs
from conftest import GOLDEN, MNIST_FIXTURE, rel_err
from oracle import svgp_literal as lit
from oracle import svgp_streamlined as st
import svgp_vae_b200 as pkg
from svgp_vae_b200 import configs

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _cmp_scalars(r1, r0):
    for k in ("inside_elbo_recon", "inside_elbo_kl", "ce_term", "KL_term"):
        a, b = float(r1[k]), float(r0[k])
        assert abs(a - b) <= TOL * abs(b), (k, a, b)


def _check(kind, cfg, clip=False, skip_grads=(), streamlined=False, **kw):
    dev = "cuda"
    o, s, op, sp = refs.make_pair(kind, cfg, dev)
    oracle = refs.streamlined_objective if streamlined else refs.oracle_objective
    r0, J0, g0 = oracle(o, op, cfg["aux"], cfg["y"], cfg["noise"], clip_pv=clip)
    r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].to(dev), cfg["y"].to(dev), cfg["noise"].to(dev), clip_pv=clip, **kw)
    assert rel_err(r1["p_m"], r0["p_m"]) < TOL and rel_err(r1["p_v"], r0["p_v"]) < TOL
    _cmp_scalars(r1, r0)
    assert abs(float(J1) - float(J0)) <= TOL * abs(float(J0))
    for i, (a, b) in enumerate(zip(g0, g1)):
        if a is not None and a.abs().max() > 0 and i not in skip_grads:
            assert rel_err(b, a) < TOL, i
    assert rel_err(r1["mu_hat"], r0["mu_hat"]) < TOL and rel_err(r1["A_hat"], r0["A_hat"]) < TOL


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("rows,b", [("eval", 256), ("train", 210)])
def test_mnist(cuda_backend, normalize, rows, b):
    _check("mnist", configs.mnist_inputs(MNIST_FIXTURE, L=16, b=b, rows=rows, normalize=normalize))


def test_sprites_linear_normalised(cuda_backend):
    _check("sprites", configs.sprites_inputs(M=72, L=64), clip=True)


def test_sprites_M500_rank_deficient(cuda_backend):
    """BASELINE's M = 500 with the reference's normalised linear x linear kernel: that kernel has rank <= 8 * 16 = 128,
    so K_mm + jI is jitter-dominated and d/dZ is conditioned like 1/jitter: fp32 storage of K (what the reference's own
    float32 SPRITES path has too) moves the inducing-point gradient by O(1) relative to float64 (same figure with the
    float64 stand-in backend on the CPU).  Everything else is held to 1e-4; dZ (index 2) is only required finite."""
    _check("sprites", configs.sprites_inputs(M=500, L=8), clip=True, skip_grads=(2,))


def test_sprites_M500_at_L64(cuda_backend):
    """configs[2] at its full channel count (b = 500, M = 500, L = 64) against the collapsed float64 oracle (the literal
    per-channel loop with its (b, m, m) tensor takes minutes at this size; the two restatements agree to 1e-12,
    refs.streamlined_objective).  Same bar as above: everything at 1e-4, dZ only finite (ill-conditioned in the reference's
    own fp32 formulation: tests/test_host_logic.py::test_sprites_M500_inducing_gradient_is_ill_conditioned_in_the_reference_itself)."""
    _check("sprites", configs.sprites_inputs(M=500, L=64), clip=True, skip_grads=(2,), streamlined=True)


def test_sprites_unnormalised_clip_active(cuda_backend):
    _check("sprites", configs.sprites_inputs(M=72, L=8, normalize=False), clip=True)


def test_sprites_se(cuda_backend):
    cfg = configs.sprites_inputs(M=72, L=8, K_SE=True)
    # the reference's sigma=0.1 / N(0,1.5^2) inputs make K ~ 1e-4 I ("trivial"); shrink the inputs
    # so that the SE factors are O(1) and the test exercises the arithmetic
    cfg["aux"][:, 1:] *= 0.3
    cfg["ctor"]["initial_inducing_points"] = cfg["ctor"]["initial_inducing_points"] * 0.2
    cfg["ctor"]["initial_GPLVM_action"] = cfg["ctor"]["initial_GPLVM_action"] * 0.2
    _check("sprites", cfg)


@pytest.mark.parametrize("tc", [False, True])
def test_sweep_small(cuda_backend, tc):
    _check("sweep", configs.sweep_inputs(2304, 200, 3), tc=tc)


@pytest.mark.parametrize("tc", [False, True])
def test_chunked_mm_stage_matches_one_chunk(cuda_backend, tc):
    """Channel-chunked float64 M x M stage (memory plan of configs[4]: M = 4096, L = 128) against the one-chunk
    stage on the same kernels: identical factorisations per channel, so only the order of a few float64 sums differs."""
    cfg = configs.sweep_inputs(4096, 256, 6)
    o, s, op, sp = refs.make_pair("sweep", cfg, "cuda")
    args = (cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda())
    r0, J0, g0 = refs.product_objective(s, sp, *args, tc=tc)
    r1, J1, g1 = refs.product_objective(s, sp, *args, tc=tc, mm_chunk=4)          # chunks of 4 + 2 channels
    for k in ("p_m", "p_v", "recon_l", "kl_l", "ce_l", "mu_hat", "A_hat"):
        assert rel_err(r1[k], r0[k]) < 1e-6, k
    for a, b in zip(g0, g1):
        assert rel_err(b, a) < 1e-6
    r2 = s.elbo_step(*args, tc=tc, mm_chunk=4, return_A_hat=False)
    assert r2["A_hat"] is None and rel_err(r2["p_v"], r0["p_v"]) < 1e-6


def test_sweep_subsample_tc_vs_streamlined(cuda_backend):
    """N = 16384, M = 256, L = 4 on the tcgen05 path against the streamlined float64 oracle (the literal form
    would need a 16384 x 256 x 256 tensor per channel)."""
    cfg = configs.sweep_inputs(16384, 256, 4)
    o, s, op, sp = refs.make_pair("sweep", cfg, "cuda")
    X, y, nz = cfg["aux"].double(), cfg["y"].double().requires_grad_(True), cfg["noise"].double().requires_grad_(True)
    for t in op:
        t.requires_grad_(True)
    Z = o.inducing_index_points
    t0 = st.streamlined_terms(o.kernel_matrix(X, Z), o.kernel_matrix(Z, Z), o.kernel_matrix(X, X, diag_only=True), y, nz,
                              cfg["ctor"]["N_train"], cfg["ctor"]["jitter"])
    g0 = st.glue_from_terms(t0, float(X.shape[0]), cfg["ctor"]["N_train"])
    gm, gv = refs.upstream(tuple(y.shape))
    J0 = g0["KL_term"] + (gm * t0["p_m"]).sum() + (gv * t0["p_v"]).sum()
    gr0 = torch.autograd.grad(J0, [y, nz] + op)
    r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), tc=True)
    assert rel_err(r1["p_m"], t0["p_m"]) < TOL and rel_err(r1["p_v"], t0["p_v"]) < TOL
    _cmp_scalars(r1, g0)
    for a, b in zip(gr0, g1):
        assert rel_err(b, a) < TOL


def test_per_channel_api_mnist(cuda_backend):
    """The reference's own calling pattern (SVGPVAE_model.py:868-873) through the drop-in methods."""
    cfg = configs.mnist_inputs(MNIST_FIXTURE, L=3)
    o, s, op, sp = refs.make_pair("mnist", cfg, "cuda")
    aux, y, nz = cfg["aux"], cfg["y"], cfg["noise"]
    r0 = lit.minibatch_glue(o, aux, y, nz)
    auxc, yc, nzc = aux.cuda(), y.cuda().requires_grad_(True), nz.cuda().requires_grad_(True)
    rec = kl = 0.0
    pms, pvs = [], []
    for l in range(3):
        pm, pv, mu, Ah = s.approximate_posterior_params(auxc, auxc, yc[:, l], nzc[:, l])
        a, b = s.variational_loss(auxc, yc[:, l], mu, Ah, nzc[:, l])
        assert rel_err(mu, r0["mu_hat"][l]) < TOL and rel_err(Ah, r0["A_hat"][l]) < TOL
        rec, kl = rec + a, kl + b
        pms.append(pm); pvs.append(pv)
    pm, pv = torch.stack(pms, 1), torch.stack(pvs, 1)
    assert rel_err(pm, r0["p_m"]) < TOL and rel_err(pv, r0["p_v"]) < TOL
    assert abs(float(rec) - float(r0["inside_elbo_recon"])) < TOL * abs(float(r0["inside_elbo_recon"]))
    assert abs(float(kl) - float(r0["inside_elbo_kl"])) < TOL * abs(float(r0["inside_elbo_kl"]))
    ce = pkg.gauss_cross_entropy(pm, pv, yc, nzc).sum()
    assert abs(float(ce) - float(r0["ce_term"])) < TOL * abs(float(r0["ce_term"]))
    (rec + kl).backward()                                   # the compat path is differentiable end to end
    assert torch.isfinite(yc.grad).all() and s.inducing_index_points.grad is not None
    K = s.kernel_matrix(auxc, s.inducing_index_points, x_inducing=False)
    assert rel_err(K, o.kernel_matrix(aux, o.inducing_index_points, x_inducing=False)) < 1e-6


def test_ball(cuda_backend):
    cfg = configs.ball_inputs()
    x, y, nz = cfg["x"], cfg["y"], cfg["noise"]
    ox, oy = lit.BallSVGP(name="x", **cfg["ctor"]), lit.BallSVGP(name="y", **cfg["ctor"])
    y64, n64 = y.double().requires_grad_(True), nz.double().requires_grad_(True)
    r0 = lit.ball_glue(ox, oy, y64, n64)
    sx, sy = pkg.SVGP(name="x", **cfg["ctor"]).cuda(), pkg.SVGP(name="y", **cfg["ctor"]).cuda()
    xc, yc, nc = x.cuda(), y.cuda().requires_grad_(True), nz.cuda().requires_grad_(True)
    rec = kl = 0.0
    pms, pvs = [], []
    for ch, s in enumerate((sx, sy)):
        pm, B, mu, Ah = s.approximate_posterior_params(xc, y=yc[:, :, ch], noise=nc[:, :, ch])
        a, b = s.variational_loss(xc, yc[:, :, ch], nc[:, :, ch], mu_hat=mu, A_hat=Ah)
        assert B.shape == (35, 30, 30) and mu.shape == (35, 15) and Ah.shape == (35, 15, 15)
        assert rel_err(B, r0["B_%d" % ch]) < TOL and rel_err(mu, r0["mu_hat_%d" % ch]) < TOL
        assert rel_err(Ah, r0["A_hat_%d" % ch]) < TOL
        rec, kl = rec + a, kl + b
        pms.append(pm); pvs.append(torch.diagonal(B, dim1=-2, dim2=-1))
    pm, pv = torch.stack(pms, 2), torch.stack(pvs, 2)
    assert rel_err(pm, r0["p_m"]) < TOL and rel_err(pv, r0["p_v"]) < TOL
    assert rel_err(rec, r0["inside_elbo_recon"]) < TOL and rel_err(kl, r0["inside_elbo_kl"]) < TOL
    ce = -pkg.gauss_cross_entropy(pm, pv, yc, nc).sum((1, 2))
    KL_term = ce + rec - kl
    assert rel_err(KL_term, r0["KL_term"]) < TOL
    # gradients of the decoder-facing objective (d KL_term / dy alone is ~1e-8: L3 and CE cancel)
    gm, gv = refs.upstream((35, 30, 2))
    g0 = torch.autograd.grad(r0["KL_term"].sum() + (gm * r0["p_m"]).sum() + (gv * r0["p_v"]).sum(), [y64, n64])
    (KL_term.sum() + (gm.cuda().float() * pm).sum() + (gv.cuda().float() * pv).sum()).backward()
    assert rel_err(yc.grad, g0[0]) < TOL and rel_err(nc.grad, g0[1]) < TOL


# ----------------------------------------------------------------------------------------------------------
# CUDA path vs tests/golden/reference_golden.npz: outputs of the unmodified reference source executed under the
# TensorFlow-API shim (tests/golden/make_reference_golden.py) -- no oracle in between.
# ----------------------------------------------------------------------------------------------------------
REF_CASES = [
    ("mnist", "mnist", lambda: configs.mnist_inputs(MNIST_FIXTURE, L=4), False),
    ("mnist_norm", "mnist", lambda: configs.mnist_inputs(MNIST_FIXTURE, L=4, normalize=True), False),
    ("mnist_train_last", "mnist", lambda: configs.mnist_inputs(MNIST_FIXTURE, L=2, b=210, rows="train", batch_index=15), False),
    ("sprites72", "sprites", lambda: configs.sprites_inputs(M=72, L=4), True),
    ("sprites72_raw", "sprites", lambda: configs.sprites_inputs(M=72, L=4, normalize=False), True),
]


@pytest.mark.parametrize("name,kind,maker,clip", REF_CASES, ids=[c[0] for c in REF_CASES])
def test_against_reference_source_golden(cuda_backend, name, kind, maker, clip):
    import os
    import numpy as np
    from conftest import GOLDEN
    gold = np.load(os.path.join(GOLDEN, "reference_golden.npz"))
    cfg = maker()
    _, s, _, sp = refs.make_pair(kind, cfg, "cuda")
    r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), clip_pv=clip)
    T = lambda k: torch.from_numpy(gold[name + "/" + k])
    assert rel_err(r1["p_m"], T("p_m")) < TOL and rel_err(r1["p_v"], T("p_v")) < TOL
    sc = gold[name + "/scalars"]
    for k, ref_v in zip(("inside_elbo_recon", "inside_elbo_kl", "ce_term", "KL_term"), sc[:4]):
        assert abs(float(r1[k]) - ref_v) <= TOL * abs(ref_v), (k, float(r1[k]), ref_v)
    assert abs(float(J1) - sc[4]) <= TOL * abs(sc[4])
    names = ["y", "noise", "Z", "table"] + (["amplitude", "length"] if kind == "mnist" else [])
    for gname, g in zip(names, g1):
        ref_g = T("grad_" + gname)
        if ref_g.abs().max() > 0:
            # every gradient at 1e-4, the two kernel hyper-parameter scalars (left over from a ~7000-fold cancellation between
            # the K_nm, K_mm and kappa contributions) included: measured 7e-7 .. 3.3e-5 (tests/probes/small_grad_errors.py)
            assert rel_err(g, ref_g) < TOL, gname
    auxc = cfg["aux"].cuda()
    K = s.kernel_matrix(auxc, s.inducing_index_points, x_inducing=False)
    assert rel_err(K, T("K_nm")) < 1e-6
    assert rel_err(s.kernel_matrix(auxc, auxc, x_inducing=False, y_inducing=False, diag_only=True), T("K_nn_diag")) < 1e-6
    if kind == "sprites":                                       # prediction-time entry, SVGPVAE_model.py:610-635
        mean, B = s.approximate_posterior_params_precomputed_GP_posterior_params(
            auxc, T("pred_mean_term").cuda(), T("pred_sigma_term").cuda())
        assert rel_err(mean, T("pred_mean")) < TOL and rel_err(B, T("pred_B")) < TOL


def test_ball_against_reference_source_golden(cuda_backend):
    import os
    import numpy as np
    from conftest import GOLDEN
    gold = np.load(os.path.join(GOLDEN, "reference_golden.npz"))
    cfg = configs.ball_inputs()
    sx, sy = pkg.SVGP(name="x", **cfg["ctor"]).cuda(), pkg.SVGP(name="y", **cfg["ctor"]).cuda()
    xc, yc, nc = cfg["x"].cuda(), cfg["y"].cuda().requires_grad_(True), cfg["noise"].cuda().requires_grad_(True)
    rec = kl = 0.0
    pms, pvs = [], []
    for ch, s in enumerate((sx, sy)):
        pm, B, mu, Ah = s.approximate_posterior_params(xc, y=yc[:, :, ch], noise=nc[:, :, ch])
        a, b = s.variational_loss(xc, yc[:, :, ch], nc[:, :, ch], mu_hat=mu, A_hat=Ah)
        if ch == 0:
            assert rel_err(B, torch.from_numpy(gold["ball/B_x"])) < TOL and rel_err(mu, torch.from_numpy(gold["ball/mu_hat_x"])) < TOL
            assert rel_err(Ah, torch.from_numpy(gold["ball/A_hat_x"])) < TOL
        rec, kl = rec + a, kl + b
        pms.append(pm); pvs.append(torch.diagonal(B, dim1=-2, dim2=-1))
    pm, pv = torch.stack(pms, 2), torch.stack(pvs, 2)
    assert rel_err(pm, torch.from_numpy(gold["ball/p_m"])) < TOL and rel_err(pv, torch.from_numpy(gold["ball/p_v"])) < TOL
    assert rel_err(rec, torch.from_numpy(gold["ball/recon"])) < TOL and rel_err(kl, torch.from_numpy(gold["ball/kl"])) < TOL
    KL_term = -pkg.gauss_cross_entropy(pm, pv, yc, nc).sum((1, 2)) + rec - kl
    assert rel_err(KL_term, torch.from_numpy(gold["ball/KL_term"])) < TOL
    gm, gv = refs.upstream((35, 30, 2))
    (KL_term.sum() + (gm.cuda().float() * pm).sum() + (gv.cuda().float() * pv).sum()).backward()
    assert rel_err(yc.grad, torch.from_numpy(gold["ball/grad_y"])) < TOL and rel_err(nc.grad, torch.from_numpy(gold["ball/grad_noise"])) < TOL


def test_prediction_path_against_reference_source(cuda_backend):
    """Prediction-time entries (SVGPVAE_model.py:989-1023, :610-635, :1048-1050) on the device vs the reference source."""
    import os
    import numpy as np
    from conftest import GOLDEN
    gold = np.load(os.path.join(GOLDEN, "reference_golden.npz"))
    T = lambda k: torch.from_numpy(gold[k])
    cfg = configs.sprites_inputs(M=72, L=4)
    _, s, _, _ = refs.make_pair("sprites", cfg, "cuda")
    aux, y, nz = cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda()
    mt, si = pkg.precompute_GP_params_SVGPVAE(y, nz, aux, s)
    assert rel_err(mt, T("sprites72/precomp_mean_terms")) < TOL and rel_err(si, T("sprites72/precomp_inv_sigma")) < TOL
    pm, pv = pkg.predict_from_precomputed(s, aux[:100], T("sprites72/precomp_mean_terms").cuda(), T("sprites72/precomp_inv_sigma").cuda())
    assert rel_err(pm, T("sprites72/precomp_p_m")) < TOL and rel_err(pv, T("sprites72/precomp_p_v")) < TOL
    cfg = configs.mnist_inputs(MNIST_FIXTURE, L=4)
    _, s, _, _ = refs.make_pair("mnist", cfg, "cuda")
    test_aux = cfg["aux"][:48].clone()
    test_aux[:, 1] += 0.3
    pm, pv = pkg.posterior_predict(s, test_aux.cuda(), cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda())
    assert rel_err(pm, T("mnist/cgen_p_m")) < TOL and rel_err(pv, T("mnist/cgen_p_v")) < TOL


def test_titsias_branch_against_reference_source(cuda_backend):
    import os
    import numpy as np
    from conftest import GOLDEN
    gold = np.load(os.path.join(GOLDEN, "reference_golden.npz"))
    T = lambda k: torch.from_numpy(gold[k])
    cfg = configs.mnist_inputs(MNIST_FIXTURE, L=2, b=64)
    cfg["ctor"]["titsias"] = True
    _, s, _, sp = refs.make_pair("mnist", cfg, "cuda")
    r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda())
    sc = gold["mnist_titsias/scalars"]
    assert abs(float(r1["inside_elbo_recon"]) - sc[0]) < TOL * abs(sc[0]) and abs(float(r1["KL_term"]) - sc[3]) < TOL * abs(sc[3])
    assert rel_err(r1["p_m"], T("mnist_titsias/p_m")) < TOL and rel_err(r1["p_v"], T("mnist_titsias/p_v")) < TOL
    for g, n in zip(g1, ["y", "noise", "Z", "table"]):
        # all four at 1e-4: measured 3.2e-5, 2.6e-5, 6.1e-6, 5.4e-6 (tests/probes/small_grad_errors.py).  dy = -cov^-1 y of the (b x b)
        # system diag(noise) + K_nm Kinv K_mn amplifies errors of K_nm by cond(cov) ~ 1e3: 1.2e-4 while the small K_nm was evaluated
        # in fp32 (5e-7 relative), 3.2e-5 with the float64 evaluation rounded once (kernel_matrix.cu, kernel_fwd_f64_kernel<float>)
        assert rel_err(g, T("mnist_titsias/grad_" + n)) < TOL, n
    cfgb = configs.ball_inputs()
    sb = pkg.SVGP(name="x", **dict(cfgb["ctor"], titsias=True)).cuda()
    xc = cfgb["x"].cuda()
    y, nz = cfgb["y"][:, :, 0].cuda().requires_grad_(True), cfgb["noise"][:, :, 0].cuda().requires_grad_(True)
    _, _, mu_hat, A_hat = sb.approximate_posterior_params(xc, y=y, noise=nz)
    L2, zero = sb.variational_loss(xc, y, nz, mu_hat=mu_hat, A_hat=A_hat)
    assert rel_err(L2, T("ball_titsias/L2")) < TOL


@pytest.mark.parametrize("geco", [False, True])
def test_forward_pass_glue_against_reference_source(cuda_backend, geco):
    """glue.forward_pass_SVGPVAE on the device vs the reference's forward_pass_SVGPVAE (:823-936) under the TF shim."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from test_host_logic import _GlueVAE, _glue_images
    gold = np.load(os.path.join(GOLDEN, "reference_golden.npz"))
    tag = "glue_geco" if geco else "glue_beta"
    cfg = configs.mnist_inputs(MNIST_FIXTURE, L=4)
    _, s, _, _ = refs.make_pair("mnist", cfg, "cuda")
    mu, var = cfg["y"].cuda().requires_grad_(True), cfg["noise"].cuda().requires_grad_(True)
    r = pkg.forward_pass_SVGPVAE((_glue_images(256).cuda(), cfg["aux"].cuda()), beta=0.7, vae=_GlueVAE(mu, var), svgp=s, C_ma=0.3,
                                 lagrange_mult=1.5, alpha=0.99, kappa=0.02, clipping_qs=True, GECO=geco,
                                 epsilon=torch.from_numpy(gold["glue/epsilon"]).cuda())
    for idx, key in ((0, "elbo"), (1, "recon_loss"), (13, "C_ma"), (14, "lagrange_mult")):
        assert abs(float(r[idx]) - float(gold[tag + "/" + key][0])) < TOL * abs(float(gold[tag + "/" + key][0])), key
    assert rel_err(r[12], torch.from_numpy(gold[tag + "/latent_samples"])) < TOL
    g = torch.autograd.grad(r[0], [mu, var, s.inducing_index_points])
    for t, n in zip(g, ("y", "noise", "Z")):
        assert rel_err(t, torch.from_numpy(gold[tag + "/grad_" + n])) < TOL, n


@pytest.mark.parametrize("normalize", [False, True])
def test_svigp_hensman_against_reference_source(cuda_backend, normalize):
    """SVIGP_Hensman on the CUDA library against the outputs of the unmodified reference source."""
    import numpy as np
    gold = np.load(os.path.join(GOLDEN, "svigp_golden.npz"))
    s, aux, gm = refs.svigp_case(normalize, "cuda", MNIST_FIXTURE)
    refs.svigp_check(s, aux, gm, gold, "svigp_norm" if normalize else "svigp", TOL, rel_err, batched=True)


@pytest.mark.parametrize("kind", ["mnist", "sprites"])
def test_cuda_graph_step_matches_eager(cuda_backend, kind):
    """GraphedElboStep (forward and backward of elbo_step captured as CUDA graphs) replays the same numbers as the
    eager step, also for new input values of the captured shapes."""
    if kind == "mnist":
        cfg = configs.mnist_inputs(MNIST_FIXTURE, L=16)
        s = pkg.mnistSVGP(name="g", **cfg["ctor"]).cuda()
        clip = False
    else:
        cfg = configs.sprites_inputs(M=72, L=8)
        s = pkg.spritesSVGP(name="g", **cfg["ctor"]).cuda()
        clip = True
    aux, y, nz = cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda()
    step = pkg.GraphedElboStep(s, aux, y, nz, clip_pv=clip)
    params = [p for p in s.parameters()]
    for scale in (1.0, 0.7):                      # second pass: different values through the same graphs
        outs = []
        for fn in (lambda a, b, c: s.elbo_step(a, b, c, clip_pv=clip), step):
            yy, nn = (y * scale).clone().requires_grad_(True), (nz * (2.0 - scale)).clone().requires_grad_(True)
            for p in params:
                p.grad = None
            r = fn(aux, yy, nn)
            gm, gv = refs.upstream(tuple(yy.shape), "cuda")
            J = r["KL_term"] + (gm.to(r["p_m"].dtype) * r["p_m"]).sum().double() + (gv.to(r["p_v"].dtype) * r["p_v"]).sum().double()
            J.backward()
            outs.append((r["p_m"].detach().clone(), r["p_v"].detach().clone(), J.detach().clone(), yy.grad.clone(), nn.grad.clone(),
                         [None if p.grad is None else p.grad.clone() for p in params]))
        e, g = outs
        # same kernels on the same inputs; the float64 atomics of the reductions make the last digits order-dependent
        for a, b in zip(e[:5], g[:5]):
            assert rel_err(b, a) < 1e-6
        for a, b in zip(e[5], g[5]):
            assert (a is None) == (b is None)
            if a is not None and a.abs().max() > 0:
                assert rel_err(b, a) < 1e-6


def test_sweep_ragged_shapes_on_the_tensor_core_path(cuda_backend):
    """N, M not multiples of the tile sizes (30000 rows, 1000 inducing points -> padded planes, ragged last row / column
    tiles, ragged SYRK window) and an odd channel count on the tensor-core path against the streamlined float64 oracle:
    every quantity at 1e-4 (round 1 needed 2e-4 .. 1e-3 here; the integer path accumulates exactly)."""
    cfg = configs.sweep_inputs(30000, 1000, 3)
    o, s, op, sp = refs.make_pair("sweep", cfg, "cuda")
    X, y, nz = cfg["aux"].double(), cfg["y"].double().requires_grad_(True), cfg["noise"].double().requires_grad_(True)
    for t in op:
        t.requires_grad_(True)
    Z = o.inducing_index_points
    t0 = st.streamlined_terms(o.kernel_matrix(X, Z), o.kernel_matrix(Z, Z), o.kernel_matrix(X, X, diag_only=True), y, nz,
                              cfg["ctor"]["N_train"], cfg["ctor"]["jitter"])
    g0 = st.glue_from_terms(t0, float(X.shape[0]), cfg["ctor"]["N_train"])
    gm, gv = refs.upstream(tuple(y.shape))
    J0 = g0["KL_term"] + (gm * t0["p_m"]).sum() + (gv * t0["p_v"]).sum()
    gr0 = torch.autograd.grad(J0, [y, nz] + op)
    r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), tc=True)
    assert rel_err(r1["p_m"], t0["p_m"]) < TOL and rel_err(r1["p_v"], t0["p_v"]) < TOL
    _cmp_scalars(r1, g0)
    for name, a, b in zip(["dy", "dnoise", "dZ", "dhyp"], gr0, g1):
        assert rel_err(b, a) < TOL, name


def test_ball_glue_graphed_matches_eager(cuda_backend):
    """glue.GraphedBallStep (forward and backward of ball_svgp_terms as CUDA graphs) == the eager product code; the
    captured Cholesky calls report into the static device word instead of being skipped."""
    cfg = configs.ball_inputs()
    sx, sy = pkg.SVGP(name="x", **cfg["ctor"]).cuda(), pkg.SVGP(name="y", **cfg["ctor"]).cuda()
    y, nz = cfg["y"].cuda().requires_grad_(True), cfg["noise"].cuda().requires_grad_(True)
    t = pkg.ball_svgp_terms(sx, sy, y, nz)
    gm, gv = refs.upstream((35, 30, 2), "cuda")
    J = t["KL_term"].sum() + (gm.float() * t["full_p_mu"]).sum() + (gv.float() * t["full_p_var"]).sum()
    g0 = torch.autograd.grad(J, [y, nz])
    step = pkg.GraphedBallStep(sx, sy, 35, 30)
    for _ in range(2):                                   # replay twice: static buffers are reused correctly
        y2, n2 = cfg["y"].cuda().requires_grad_(True), cfg["noise"].cuda().requires_grad_(True)
        pm, pv, kl = step(y2, n2)
        J2 = kl.sum() + (gm.float() * pm).sum() + (gv.float() * pv).sum()
        g1 = torch.autograd.grad(J2, [y2, n2])
        assert rel_err(pm, t["full_p_mu"]) < 1e-6 and rel_err(pv, t["full_p_var"]) < 1e-6 and rel_err(kl, t["KL_term"]) < 1e-6
        assert rel_err(g1[0], g0[0]) < 1e-5 and rel_err(g1[1], g0[1]) < 1e-5
    step.check()                                         # no bad pivot recorded
    # a non-positive-definite replay is reported, not silently NaN
    bad = torch.full_like(n2, -1.0)
    step(y2.detach(), bad)
    with pytest.raises(pkg.ops.NotPositiveDefinite):
        step.check()
