"""GPU parity of every C-ABI entry point against the float64 oracle backend (same seeded inputs)."""
import pytest
import torch

from conftest import rel_err
from oracle_backend import OracleBackend, kernel_value
from svgp_vae_b200.backend import Kop

pytestmark = pytest.mark.gpu
ORA = OracleBackend()
SPECS = {
    "ball_rbf": (1, 1, 0, 0),
    "mnist_lin": (2, 1, 3, 8),
    "mnist_cos": (2, 1, 4, 8),
    "sprites_lin": (3, 8, 3, 16),
    "sprites_cos": (4, 8, 4, 16),
    "sprites_se": (1, 8, 1, 16),
    "sweep_se": (1, 4, 1, 4),
}


def _feat(spec, N, M, seed=0):
    g = torch.Generator().manual_seed(seed)
    d = spec[1] + spec[3]
    Fx, Fz = torch.randn(N, d, generator=g), torch.randn(M, d, generator=g)
    hyp = torch.tensor([0.9, 1.3, 1.1, 0.8])
    return Fx, Fz, hyp


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("shape", [(1, 1), (37, 15), (300, 72), (1000, 130)])
def test_kernel_fwd_bwd(cuda_backend, name, shape):
    be, spec = cuda_backend, SPECS[name]
    N, M = shape
    Fx, Fz, hyp = _feat(spec, N, M)
    ref = kernel_value(spec, Fx, Fz, hyp)
    kop = be.kernel_fwd(spec, Fx.cuda(), Fz.cuda(), hyp.cuda())
    assert rel_err(kop.K, ref) < 5e-6                     # element-wise op: far tighter than the 1e-4 bar
    kopt = be.kernel_fwd(spec, Fx.cuda(), Fz.cuda(), hyp.cuda(), tc=True, i8=False)
    # fp16 pair planes: 22 significand bits down to the subnormal floor of the lo plane (2^-24 in scaled units,
    # i.e. 2^-38 of the plane bound 2^14 / scale) -- entries below that floor vanish, by design
    floor = 2.0 ** -24 * float(kopt.kscale[1])
    refc = ref.cuda()
    assert float((kopt.value().double() - refc).abs().max()) < 5e-6 * float(refc.abs().max()) + floor
    Kt = kopt.value_t()
    assert float((Kt.double() - refc.t()).abs().max()) < 5e-6 * float(refc.abs().max()) + floor
    assert float(kopt.value().abs().max() * kopt.kscale[0]) < 2 ** 14 + 1
    G = torch.randn(N, M, generator=torch.Generator().manual_seed(1))
    rx, rz, rh = ORA.kernel_bwd(spec, Fx, Fz, hyp, G)
    dx, dz, dh = be.kernel_bwd(spec, Fx.cuda(), Fz.cuda(), hyp.cuda(), G.cuda())
    assert rel_err(dx, rx) < 2e-5 and rel_err(dz, rz) < 2e-5
    if rh.abs().max() > 0:
        assert rel_err(dh, rh) < 2e-5
    # diag
    Fy = torch.randn(N, spec[1] + spec[3], generator=torch.Generator().manual_seed(2))
    kd = be.kernel_diag_fwd(spec, Fx.cuda(), Fy.cuda(), hyp.cuda())
    assert rel_err(kd, kernel_value(spec, Fx, Fy, hyp, pairwise=False)) < 5e-6
    g = torch.randn(N, generator=torch.Generator().manual_seed(3))
    ex, ey, eh = ORA.kernel_diag_bwd(spec, Fx, Fy, hyp, g)
    fx, fy, fh = be.kernel_diag_bwd(spec, Fx.cuda(), Fy.cuda(), hyp.cuda(), g.cuda())
    assert rel_err(fx, ex) < 2e-5 and rel_err(fy, ey) < 2e-5
    if eh.abs().max() > 0:
        assert rel_err(fh, eh) < 2e-5


@pytest.mark.parametrize("name", ["mnist_cos", "sweep_se"])
def test_kernel_fwd_large_plain_uses_the_fp32_tile_builder(cuda_backend, name):
    """Plain fp32 K_nm above 2^26 entries comes from the tiled fp32 builder (below that from the float64-evaluating kernel,
    test_kernel_fwd_bwd): both against the float64 oracle; the small one is held to ONE fp32 rounding."""
    be, spec = cuda_backend, SPECS[name]
    Fx, Fz, hyp = _feat(spec, 66000, 1024)                    # 67.6 M entries
    ref = kernel_value(spec, Fx.cuda(), Fz.cuda(), hyp.cuda())
    kop = be.kernel_fwd(spec, Fx.cuda(), Fz.cuda(), hyp.cuda())
    assert rel_err(kop.K, ref) < 5e-6
    small = be.kernel_fwd(spec, Fx[:4000].cuda(), Fz.cuda(), hyp.cuda())
    r = ref[:4000]
    # one rounding to fp32 of the float64 value (+ the float64 evaluation's own last bits on entries that cancel to ~0)
    assert ((small.K.double() - r).abs() <= 2.0 ** -24 * 1.02 * r.abs() + 1e-13 * float(r.abs().max())).all()


def test_gather_scatter(cuda_backend):
    be = cuda_backend
    g = torch.Generator().manual_seed(0)
    table = torch.randn(72, 8, generator=g)
    ids = torch.sort(torch.randint(0, 72, (500,), generator=g)).values
    ids = torch.cat([ids, torch.randint(0, 72, (77,), generator=g)])           # grouped runs + a ragged random tail
    out = be.gather_rows(table.cuda(), ids.cuda())
    assert torch.equal(out.cpu(), table[ids])
    G = torch.randn(ids.shape[0], 8, generator=g)
    dt = be.scatter_add_rows(G.cuda(), ids.cuda(), 72)
    assert rel_err(dt, ORA.scatter_add_rows(G, ids, 72)) < 1e-6


@pytest.mark.parametrize("shape", [(1, 1, 1), (30, 15, 35), (256, 32, 16), (500, 72, 64), (700, 500, 3), (5000, 130, 2)])
def test_simt_gemm_family(cuda_backend, shape):
    be = cuda_backend
    N, M, L = shape
    g = torch.Generator().manual_seed(N + M + L)
    K = torch.randn(N, M, generator=g)
    W = torch.randn(N, L, generator=g)
    S = torch.randn(L, M, M, generator=g, dtype=torch.float64)
    S = S + S.transpose(1, 2)
    Lt = torch.tril(torch.randn(L, M, M, generator=g, dtype=torch.float64))
    kop_c, kop = Kop(K), Kop(K.cuda())
    assert rel_err(be.syrk(kop, W.cuda()), ORA.syrk(kop_c, W)) < 1e-5
    assert rel_err(be.gemm_tn(kop, W.cuda()), ORA.gemm_tn(kop_c, W)) < 1e-5
    Wm = torch.randn(L, M, generator=g)
    assert rel_err(be.gemm_nn(kop, Wm.cuda()), ORA.gemm_nn(kop_c, Wm)) < 1e-5
    assert rel_err(be.rowquad(kop, S.cuda()), ORA.rowquad(kop_c, S)) < 1e-5
    assert rel_err(be.rowquad(kop, Lt.cuda(), tri=True), ORA.rowquad(kop_c, Lt, tri=True)) < 1e-5
    assert rel_err(be.scaled_gemm(kop, W.cuda(), S.cuda()), ORA.scaled_gemm(kop_c, W, S)) < 1e-5
    out = torch.ones(N, M).cuda()
    _, dots = be.scaled_gemm(kop, W.cuda(), S.cuda(), out=out, ndot=L)
    assert rel_err(out - 1, ORA.scaled_gemm(kop_c, W, S)) < 1e-5
    assert rel_err(dots, ORA.rowquad(kop_c, S)) < 1e-5
    B = torch.randn(L, M, generator=g)
    assert rel_err(be.gemm_f32(W.cuda(), B.cuda()), ORA.gemm_f32(W, B)) < 1e-5


@pytest.mark.parametrize("M,B", [(1, 1), (15, 35), (32, 16), (33, 2), (72, 64), (500, 3), (1024, 2)])
def test_linalg_f64(cuda_backend, M, B):
    be = cuda_backend
    g = torch.Generator().manual_seed(M)
    X = torch.randn(B, M, M + 3, generator=g, dtype=torch.float64)
    X = X @ X.transpose(1, 2) / M + 0.1 * torch.eye(M, dtype=torch.float64)
    Lf, status = be.chol(X.cuda())
    assert int(status.abs().sum()) == 0
    assert rel_err(Lf, torch.linalg.cholesky(X)) < 1e-11
    Linv = be.trinv(Lf)
    assert rel_err(Linv, ORA.trinv(torch.linalg.cholesky(X))) < 1e-10
    assert rel_err(be.ltl(Linv), torch.linalg.inv(X)) < 1e-9                  # S = Linv^T Linv (triangular-aware)
    Y = torch.randn(B, M, 7, generator=g, dtype=torch.float64)
    for tA in (False, True):
        for tB in (False, True):
            A_ = X.transpose(1, 2).contiguous() if tA else X
            B_ = Y.transpose(1, 2).contiguous() if tB else Y
            assert rel_err(be.bmm64(A_.cuda(), B_.cuda(), tA, tB), X @ Y) < 1e-12
    assert rel_err(be.bmm64(X[:1].cuda(), Y.cuda()), X[:1] @ Y) < 1e-12       # broadcast batch of one
    if M >= 96:                                                                # 128 x 128 register-tiled kernel, ragged edges
        Z = torch.randn(B, M, M - 5, generator=g, dtype=torch.float64)
        for tA in (False, True):
            for tB in (False, True):
                A_ = X.transpose(1, 2).contiguous() if tA else X
                B_ = Z.transpose(1, 2).contiguous() if tB else Z
                assert rel_err(be.bmm64(A_.cuda(), B_.cuda(), tA, tB), X @ Z) < 1e-12
    # a non-PD matrix is reported, not silently NaN-propagated
    bad = X.clone()
    bad[0, M - 1, M - 1] = -1.0
    _, status = be.chol(bad.cuda())
    assert int(status[0]) == M and int(status[1:].abs().sum()) == 0


def test_row_terms(cuda_backend):
    be = cuda_backend
    g = torch.Generator().manual_seed(0)
    N, L = 1037, 70
    y, noise, kappa = torch.randn(N, L, generator=g), torch.rand(N, L, generator=g) + 0.01, torch.rand(N, generator=g)
    noise[3, 5] = 0.0
    noise_ref = noise.clone()
    p, py, sums = be.rowstats(y.cuda(), noise.cuda(), kappa.cuda())
    noise_ref[3, 5] = 1.0                                                    # log(0) only appears in the reference as -inf
    rp, rpy, rs = ORA.rowstats(y, noise, kappa)
    assert rel_err(p, rp) < 1e-6 and rel_err(py, rpy) < 1e-6 and float(p[3, 5]) == 0.0
    assert rel_err(sums[:2], rs[:2]) < 1e-6
    h, q1 = torch.rand(N, generator=g), torch.rand(N, L, generator=g) * 2 - 0.5
    pv, cs, mask = be.predictive(kappa.cuda(), h.cuda(), q1.cuda(), p, clip=(0.3, 1.2))
    rpv, rcs, rmask = ORA.predictive(kappa, h, q1, rp, clip=(0.3, 1.2))
    assert rel_err(pv, rpv) < 1e-6 and rel_err(cs, rcs) < 1e-5 and torch.equal(mask.cpu(), rmask)
    pv2, _, _ = be.predictive(kappa.cuda(), h.cuda(), q1.cuda(), p)
    assert rel_err(pv2, kappa[:, None] - h[:, None] + q1) < 1e-6
