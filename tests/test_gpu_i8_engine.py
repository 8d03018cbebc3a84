"""Integer tensor-core path (tcgen05.mma.kind::i8, csrc/tc_i8_engine.cu + i8_planes.cu) against a digit-exact emulation.

The integer MMAs do not round, so the checks are much sharper than 1e-4: the digit planes must reassemble to the
operand within half a unit of their fixed-point grid, the SYRK must equal the sum of the ten kept digit-plane pairs
computed in float64 from THE SAME planes to ~1e-13, and the scaled GEMM (one fp32 rounding per accumulator tile) to 1e-6.
"""
import pytest
import torch

from conftest import rel_err
from oracle_backend import kernel_value
from svgp_vae_b200 import configs
from svgp_vae_b200._lib import IMPL_TC_I8, IMPL_TC_I8_D3, IMPL_TC_I8_O4

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _kop(be, N, M, L=2):
    cfg = configs.sweep_inputs(N, M, L, device="cuda")
    Z = torch.from_numpy(cfg["ctor"]["initial_inducing_points"]).float().cuda()
    hyp = torch.ones(4, device="cuda")
    kop = be.kernel_fwd((1, 4, 1, 4), cfg["aux"].float().contiguous(), Z.contiguous(), hyp, tc=True, i8=True)
    kop.K64 = kernel_value((1, 4, 1, 4), cfg["aux"], Z, hyp)                # float64 kernel values (oracle formulas)
    return cfg, kop


def _ints(planes):
    d = _digits(planes)
    return ((d[0] * 256 + d[1]) * 256 + d[2]) * 256 + d[3]


def _digits(planes):
    """(S, ...) int8 planes -> list of float64 digit tensors (most significant first)."""
    return [planes[s].double() for s in range(planes.shape[0])]


def _balanced_digits4(V):
    """Balanced base-256 digits of an int64 tensor |V| < 2^31 (most significant first), as float64."""
    out = []
    for _ in range(3):
        d = ((V + 128) % 256) - 128
        out.append(d.double())
        V = (V - d) // 256
    out.append(V.double())
    return out[::-1]


@pytest.mark.parametrize("N,M", [(4096, 256), (5000, 320)])
def test_kplanes_reassemble(cuda_backend, N, M):
    be = cuda_backend
    _, kop = _kop(be, N, M)
    K = kop.K64
    Kr, Kc = kop.value_i8("r"), kop.value_i8("c")
    # fp32 kernel arithmetic (a few ulp of the entry) + one unit of the 32-bit grid relative to the row / column maximum
    # power-of-two grids of 2^-30 of the (rounded-up) row / column maximum; the evaluation is float64 rounded once to fp32
    unit_r, unit_c = K.abs().amax(1, keepdim=True) / 2.0 ** 29, K.abs().amax(0, keepdim=True) / 2.0 ** 29
    assert ((Kr - K).abs() <= 1e-7 * K.abs() + 0.51 * unit_r).all()
    assert ((Kc - K).abs() <= 1e-7 * K.abs() + 0.51 * unit_c).all()
    assert kop.Kr[0].int().abs().max() <= 127 and kop.Kc[0].int().abs().max() <= 127
    # the two sets of planes hold the same fp32 number wherever it fits both grids (entries within 2^-6 of both maxima)
    big = (K.abs() * 64 >= K.abs().amax(1, keepdim=True)) & (K.abs() * 64 >= K.abs().amax(0, keepdim=True))
    assert big.any() and (Kr[big] == Kc[big]).all()
    # the fp16 hi / lo row planes of the same call
    assert rel_err(kop.value(), K) < 2e-6
    # pad columns / the ragged last block are zeros (TMA boxes read them)
    assert (kop.Kr[:, :, M:] == 0).all() and (kop.Kc.permute(0, 1, 3, 2).reshape(4, -1, M)[:, N:] == 0).all()


@pytest.mark.parametrize("R,C,S", [(300, 256, 4), (64, 1000, 4), (128, 384, 3)])
def test_split_i8(cuda_backend, R, C, S):
    be = cuda_backend
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(2, R, C, generator=g, device="cuda", dtype=F64) * torch.exp(3 * torch.randn(2, R, 1, generator=g, device="cuda", dtype=F64))
    X[1, 3] = 0.0
    P = be.planes_i8(X, nslices=S)
    top = 127.0 * 256.0 ** (S - 1)
    err = (P.value() - X).abs().amax(-1)
    assert (err <= 0.51 * X.abs().amax(-1) / top * (1 + 1e-6) + 1e-300).all()
    assert (P.planes[:, :, C:] == 0).all()


@pytest.mark.parametrize("pair,split", [(True, True), (True, False), (False, False), (True, "full")])
@pytest.mark.parametrize("N,M,L", [(4096, 256, 3), (6000, 384, 2), (20000, 128, 1), (3000, 1024, 1), (2500, 256, 40), (9000, 4096, 1)])
def test_syrk_i8_is_exact(cuda_backend, N, M, L, pair, split, monkeypatch):
    _check_syrk_i8(cuda_backend, N, M, L, pair, split, False, monkeypatch)


@pytest.mark.parametrize("pair,split", [(True, True), (False, False), (True, "full")])
@pytest.mark.parametrize("N,M,L", [(4096, 256, 3), (6000, 384, 2), (9000, 4096, 1)])
def test_syrk_i8_three_leading_digits(cuda_backend, N, M, L, pair, split, monkeypatch):
    """SVGP_IMPL_TC_I8_D3 (the adjoint SYRK of the step): the eight digit-plane pairs with t, u <= 2 -- still exact against the
    emulation of the same eight pairs, and within fp32 operand accuracy of the float64 contraction."""
    _check_syrk_i8(cuda_backend, N, M, L, pair, split, True, monkeypatch)


@pytest.mark.parametrize("N,M,L", [(4096, 256, 3), (6000, 384, 2), (9000, 4096, 1)])
def test_syrk_i8_thirteen_pairs(cuda_backend, N, M, L, monkeypatch):
    """SVGP_IMPL_TC_I8_O4 (the forward SYRK of the step above M = 2048): the three pairs of order 4 as a second set of work items
    on the pair kernel, full form."""
    _check_syrk_i8(cuda_backend, N, M, L, True, "full", "o4", monkeypatch)


def _check_syrk_i8(cuda_backend, N, M, L, pair, split, d3, monkeypatch):
    """pair: the CTA-pair kernel (tcgen05.mma.cta_group::2; the weighted operand is the B side: rows b) / the single-CTA
    kernel (weighted operand = A side: rows a).  Both must equal the digit-exact emulation of their own operand placement.
    split (default): the diagonal blocks (2 t + 1, 2 t + 1), whose pair tile would lie half above the diagonal, run on the
    single-CTA kernel after the pair kernel."""
    monkeypatch.setenv("SVGP_I8_PAIR", "1" if pair else "0")
    monkeypatch.setenv("SVGP_I8_SYRK_SPLIT", "1" if split else "0")
    full = split == "full"          # both triangles of K^T (w o K) on the pair kernel, averaged (default for M > 2048) instead of lower + mirror
    monkeypatch.setenv("SVGP_I8_SYRK_FULL", "1" if full else "0")
    be = cuda_backend
    _, kop = _kop(be, N, M, L)
    g = torch.Generator(device="cuda").manual_seed(2)
    W = torch.randn(N, L, generator=g, device="cuda") * torch.exp(torch.randn(N, L, generator=g, device="cuda"))
    o4, d3 = d3 == "o4", d3 is True
    A = be.syrk(kop, W.contiguous(), impl=IMPL_TC_I8_O4 if o4 else (IMPL_TC_I8_D3 if d3 else IMPL_TC_I8))
    # emulation from the same planes: weighted operand = rn(float(Kint) * (w / wmax) * q_a) as a 32-bit integer, q_a from
    # the largest |float(Kint) * (w / wmax)| of row a
    k = _digits(kop.Kc)                                                   # (nblk, M, 128) each
    Kint = _ints(kop.Kc).permute(0, 2, 1).reshape(-1, M)[:N]              # (N, M)
    kd = [x.permute(0, 2, 1).reshape(-1, M)[:N] for x in k]
    cs = kop.cscale.double()
    ref = torch.empty_like(A)
    for l in range(L):
        wmax = W[:, l].abs().max()
        wt = W[:, l] / wmax
        prod = Kint.to(torch.int64).float() * wt[:, None]
        qa = (torch.tensor(2130706432.0, device="cuda") * torch.tensor(0.99999, device="cuda")) / prod.abs().amax(0)
        V = torch.round(prod * qa[None, :]).to(torch.int64)
        v = _balanced_digits4(V)
        def emulate(weighted_is_b):
            acc = [torch.zeros(M, M, dtype=F64, device="cuda") for _ in range(4)]
            for t in range(4):
                for u in range(4):
                    if t + u <= 3 and not (d3 and max(t, u) == 3):
                        acc[t + u] += (kd[u].t() @ v[t]) if weighted_is_b else (v[t].t() @ kd[u])      # exact: |sum| < 2^53
            i64 = ((acc[0].to(torch.int64) * 256 + acc[1].to(torch.int64)) * 256 + acc[2].to(torch.int64)) * 256 + acc[3].to(torch.int64)
            if o4:                                          # order 4: (1, 3) (2, 2) (3, 1), a 256th of the order-3 unit
                acc4 = sum((kd[4 - t].t() @ v[t]) if weighted_is_b else (v[t].t() @ kd[4 - t]) for t in (1, 2, 3))
                return (i64.double() + acc4 / 256.0) * (16777216.0 * wmax.double()) * cs[:, None] * cs[None, :] / (qa.double()[None, :] if weighted_is_b else qa.double()[:, None])
            return i64.double() * (16777216.0 * wmax.double()) * cs[:, None] * cs[None, :] / (qa.double()[None, :] if weighted_is_b else qa.double()[:, None])
        ref[l] = emulate(pair)
        if full:
            ref[l] = 0.5 * (ref[l] + ref[l].t())
        elif pair and split:
            single = emulate(False)
            for b in range(1, (M + 127) // 128, 2):
                ref[l][128 * b:128 * b + 128, 128 * b:128 * b + 128] = single[128 * b:128 * b + 128, 128 * b:128 * b + 128]
    if not full:
        ref = torch.tril(ref) + torch.tril(ref, -1).transpose(-1, -2)
    assert rel_err(A, ref) < 1e-12 and float((A - A.transpose(-1, -2)).abs().max()) == 0.0
    # and against the plain float64 contraction of the float64 kernel values: fp32 kernel arithmetic is what is left
    full = torch.einsum('il,ia,ib->lab', W.double(), kop.K64, kop.K64)
    assert rel_err(A, full) < (2e-6 if d3 else 1e-6)


@pytest.mark.parametrize("pair,wide", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("N,M,L,ndot", [(4096, 256, 4, 2), (3000, 384, 3, 3), (2048, 1024, 2, 0), (2304, 4096, 2, 1), (2304, 330, 3, 2)])
def test_scaled_gemm_i8(cuda_backend, N, M, L, ndot, pair, wide, monkeypatch):
    _check_scaled_gemm_i8(cuda_backend, N, M, L, ndot, pair, wide, L, monkeypatch, debias=True)


@pytest.mark.parametrize("pair,wide", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("N,M,L,ndot,nfull,debias", [(4096, 256, 4, 2, 2, False), (3000, 384, 3, 3, 1, True), (2304, 4096, 2, 1, 1, True),
                                                     (2304, 330, 3, 0, 0, False)])
def test_scaled_gemm_i8_three_leading_digits(cuda_backend, N, M, L, ndot, nfull, debias, pair, wide, monkeypatch):
    """nfull < L: the matrices s >= nfull (the S_l - Kinv family of pass D) are multiplied with the eight digit-plane pairs
    t, u <= 2 only; the planes that are not needed are not fetched either.  Exact against the emulation of the same pairs."""
    _check_scaled_gemm_i8(cuda_backend, N, M, L, ndot, pair, wide, nfull, monkeypatch, debias=debias)


def _check_scaled_gemm_i8(cuda_backend, N, M, L, ndot, pair, wide, nfull, monkeypatch, debias=True):
    """pair: CTA pairs (tcgen05.mma.cta_group::2, 256 x 128 tiles) / single CTAs; wide: six MMAs per k-step (four of them N = 256
    over two neighbouring digit planes and two neighbouring accumulators) / the ten N = 128 MMAs.  Same ten digit-plane
    products in every variant, so all four must equal the same digit-exact emulation."""
    monkeypatch.setenv("SVGP_I8_PAIR", "1" if pair else "0")
    monkeypatch.setenv("SVGP_I8_WIDE", "1" if wide else "0")
    be = cuda_backend
    _, kop = _kop(be, N, M, L)
    g = torch.Generator(device="cuda").manual_seed(3)
    R = torch.randn(L, M, M, generator=g, device="cuda", dtype=F64)
    d = torch.exp(1.5 * torch.randn(L, M, generator=g, device="cuda", dtype=F64))            # wide dynamic range, symmetric
    G = d[:, :, None] * (R + R.transpose(-1, -2)) * d[:, None, :]
    W = torch.randn(N, L, generator=g, device="cuda")
    P = be.planes_i8(G, nslices=4)
    res = be.scaled_gemm_i8(kop, W.contiguous(), P, ndot=ndot, nfull=nfull, debias=debias)
    out, dots = res if ndot else (res, None)
    # emulation: kept digit-plane pairs, exact
    kd = [x[:, :M] for x in _digits(kop.Kr)]
    gd = [x.reshape(L, M, -1)[:, :, :M] for x in _digits(P.planes)]
    rs, gs = kop.rscale.double(), P.scale.double().reshape(L, M)
    Kint = _ints(kop.Kr)[:, :M]
    kb, gb = be.pair_bias_i8(kop.Kr).double(), be.pair_bias_i8(P.planes).double()
    ref = torch.zeros(N, M, dtype=F64, device="cuda")
    dref = torch.zeros(N, max(ndot, 1), dtype=F64, device="cuda")
    for s in range(L):
        T = torch.zeros(N, M, dtype=F64, device="cuda")
        for t in range(4):
            for u in range(4):
                if t + u <= 3 and not (s >= nfull and max(t, u) == 3):
                    T += (kd[t] @ gd[u][s].t()) * 256.0 ** (3 - t - u)
        # + the expected value of the dropped pairs (svgp_i8_pair_bias of both operands), as the kernel adds it before rounding
        if debias:
            T += kb[int(s >= nfull)][:, None] + gb[int(s >= nfull)][s * M:(s + 1) * M][None, :] - 3.0 * M / 1024.0
        T = T * gs[s][None, :]                                            # plane units of the kernel's `tv`
        ref += W[:, s:s + 1].double() * rs[:, None] * 16777216.0 * T
        if s < ndot:
            dref[:, s] = (T * Kint).sum(1) * rs * rs * 16777216.0
    assert rel_err(out, ref) < 2e-6
    if ndot:
        assert rel_err(dots, dref[:, :ndot]) < 2e-6
    # against the float64 product of the un-quantised operands
    full = torch.einsum('il,ia,lac->ic', W.double(), kop.K64, G)
    assert rel_err(out, full) < 1e-5
    # accumulate flag
    base = torch.ones_like(out)
    out2 = be.scaled_gemm_i8(kop, W.contiguous(), P, out=base.clone(), nfull=nfull, debias=debias)
    assert rel_err(out2 - 1.0, ref) < 5e-6


def test_gemm_nn_i8(cuda_backend):
    be = cuda_backend
    N, M, L = 4096, 384, 5
    _, kop = _kop(be, N, M, L)
    g = torch.Generator(device="cuda").manual_seed(4)
    Wm = torch.randn(L, M, generator=g, device="cuda", dtype=F64) * torch.exp(2 * torch.randn(L, M, generator=g, device="cuda", dtype=F64))
    out = be.gemm_nn(kop, Wm)
    assert out.shape == (N, L)
    assert rel_err(out, kop.K64 @ Wm.t()) < 2e-6


def test_pair_bias_i8(cuda_backend):
    be = cuda_backend
    _, kop = _kop(be, 3000, 330, 2)
    b = be.pair_bias_i8(kop.Kr).double()
    d = _digits(kop.Kr)
    s123 = (d[1] + d[2] + d[3]).sum(1)
    assert torch.equal(b[0], (-s123 / 512.0).float().double())
    assert rel_err(b[1], -s123 / 512.0 - 0.5 * d[0].sum(1)) < 1e-6


@pytest.mark.parametrize("nfull", [1, 0])
def test_scaled_gemm_i8_dropped_pairs_are_unbiased(cuda_backend, nfull):
    """The digits of the format have mean -1/2, so the digit-plane pairs the product drops do NOT average to zero: without the
    correction every entry is off by the same few units of the order-3 accumulator (1e-9 of a typical entry -- and the reason
    the kernel hyper-parameter gradients missed 1e-4 at M = 4096).  Made visible with a G whose two leading digit planes are
    empty (small products: fp32 resolves single units): mean deviation from the exact 16-pair product, in accumulator units."""
    from svgp_vae_b200.backend import PlanesI8
    be = cuda_backend
    N, M, Mc = 2304, 4096, 1024
    _, kop = _kop(be, N, M, 1)
    g = torch.Generator(device="cuda").manual_seed(5)
    planes = torch.zeros(4, Mc, M, dtype=torch.int8, device="cuda")
    planes[2:] = torch.randint(-128, 128, (2, Mc, M), generator=g, device="cuda", dtype=torch.int8)
    P = PlanesI8(planes, torch.ones(Mc, device="cuda"), 1, Mc, M)
    Kint = _ints(kop.Kr)[:, :M]
    exact = Kint @ _ints(planes).t() / 16777216.0                                               # accumulator units
    unit = (kop.rscale.double() * 16777216.0)[:, None]
    dev = {}
    for debias in (False, True):
        out = be.scaled_gemm_i8(kop, None, P, nfull=nfull, debias=debias)
        dev[debias] = float((out.double() / unit - exact).mean())
    # (1) without the correction the kernel is off by exactly the pairs it drops (all ten kept: (1, 3) (2, 2) ... of the non-empty
    # planes -- a digit of K against a digit of mean -1/2, i.e. about -(S_1 + S_2)[i] / 512 where S are the digit sums of K's row:
    # NOT -M / 2 each, the leading digit of a small entry is positive; eight pairs: also (0, 3), -S_0[i] / 2), (2) the correction
    # shifts the mean by exactly bias_K + bias_G - 3 M / 1024 although it is far below one fp32 rounding of an entry, (3) what is
    # left is the zero-mean part of the dropped pairs for THIS G (one realisation, shared by all rows)
    kd, gd = _digits(kop.Kr), _digits(planes)
    lvl = 0 if nfull else 1
    kept = lambda t, u: t + u <= 3 and not (lvl == 1 and max(t, u) == 3)
    dropped = sum((kd[t][:, :M] @ gd[u].t()) * 256.0 ** (3 - t - u) for t in range(4) for u in (2, 3) if not kept(t, u))
    corr = be.pair_bias_i8(kop.Kr).double()[lvl][:, None] + be.pair_bias_i8(planes).double()[lvl][None, :] - 3.0 * M / 1024.0
    scale = max(1.0, abs(float(dropped.mean())))
    assert abs(dev[False] + float(dropped.mean())) < 1.0 + 1e-3 * scale
    assert abs((dev[True] - dev[False]) - float(corr.mean())) < 1.0 + 1e-3 * scale
    assert abs(dev[True]) < 0.3 * abs(dev[False])
