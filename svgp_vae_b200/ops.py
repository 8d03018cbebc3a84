"""Differentiable primitives over the C ABI (torch.autograd.Function with hand-written backward).

Two families:

* float64 M x M algebra for the per-channel stage K3 (``bmm64``, ``spd_inverse_logdet``,
  ``spd_logdet``): torch autograd only chains them, every product / factorisation runs in
  svgp_gemm_f64 / svgp_chol_f64 / svgp_trinv_f64.
* datapoint-sized primitives on a plain fp32 K_nm (``kernel_matrix``, ``kernel_diag``,
  ``gather_rows``, ``syrk``, ``rowquad``, ``kt_matmul``, ``k_matmul``): these back the reference's
  per-channel call signature (svgp.py).  The batched hot path (step.py) calls the same backend
  entry points directly and never materialises K_nm as an autograd tensor.

Adjoint identities used (A.5 of SURVEY.md, re-derived for this operator set):
  A_l = sum_i w_il k_i k_i^T      =>  dK_i += sum_l w_il (G_l + G_l^T) k_i,   dw_il = k_i^T G_l k_i
  q_il = k_i^T S_l k_i            =>  dS_l  = sum_i g_il k_i k_i^T,           dK_i += sum_l 2 g_il S_l k_i
"""
import torch

from .backend import Kop, get_backend


def _sym(G):
    return 0.5 * (G + G.transpose(-1, -2))


# ------------------------------------------------------------------------------------------
# float64 M x M algebra
# ------------------------------------------------------------------------------------------
class _BMM64(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, transA, transB):
        ctx.save_for_backward(A, B)
        ctx.flags = (transA, transB)
        return get_backend().bmm64(A.contiguous(), B.contiguous(), transA, transB)

    @staticmethod
    def backward(ctx, G):
        A, B = ctx.saved_tensors
        tA, tB = ctx.flags
        be = get_backend()
        G = G.contiguous()
        gA = gB = None
        if ctx.needs_input_grad[0]:
            if not tA:
                gA = be.bmm64(G, B, False, not tB)          # G op(B)^T
            else:
                gA = be.bmm64(B, G, tB, True)               # op(B) G^T
            if A.shape[0] == 1 and gA.shape[0] > 1:
                gA = gA.sum(0, keepdim=True)
        if ctx.needs_input_grad[1]:
            if not tB:
                gB = be.bmm64(A, G, not tA, False)          # op(A)^T G
            else:
                gB = be.bmm64(G, A, True, tA)               # G^T op(A)
            if B.shape[0] == 1 and gB.shape[0] > 1:
                gB = gB.sum(0, keepdim=True)
        return gA, gB, None, None


def bmm64(A, B, transA=False, transB=False):
    """Batched float64 product op(A) op(B) on (B, m, k) tensors; a batch of 1 broadcasts."""
    return _BMM64.apply(A, B, transA, transB)


def bmv64(A, x):
    """Batched mat-vec: (B, m, k) x (B, k) -> (B, m)."""
    return bmm64(A, x.unsqueeze(-1)).squeeze(-1)


class NotPositiveDefinite(RuntimeError):
    """Raised when a Cholesky pivot is <= 0 (the reference raises InvalidArgumentError at sess.run)."""


_PD_FLAGS = {}


def pd_flag(device):
    """Static device word (int32[1]) into which Cholesky calls captured in a CUDA graph write their worst status: no host
    synchronisation is possible under capture, so the bad pivot is recorded on the device and read later (pd_flag_check).
    Create it BEFORE capturing."""
    key = torch.device(device)
    if key not in _PD_FLAGS:
        _PD_FLAGS[key] = torch.zeros(1, dtype=torch.int32, device=key)
    return _PD_FLAGS[key]


def pd_flag_check(device):
    """Read (and clear) the device word: raises NotPositiveDefinite if a replayed graph met a non-positive pivot."""
    f = _PD_FLAGS.get(torch.device(device))
    if f is not None and int(f.item()) != 0:
        piv = int(f.item()) - 1
        f.zero_()
        raise NotPositiveDefinite("a CUDA-graph replay met a non-positive Cholesky pivot (column %d)" % piv)


def _check_status(status, what):
    if status.is_cuda and torch.cuda.is_current_stream_capturing():
        # no host synchronisation inside a CUDA graph (graphed.py): record the worst status in the static device word
        f = _PD_FLAGS.get(status.device)
        if f is not None:
            f.copy_(torch.maximum(f, status.max().reshape(1).to(torch.int32)))
        return
    bad = torch.nonzero(status)
    if bad.numel():
        b = int(bad[0, 0])
        raise NotPositiveDefinite("%s: matrix %d is not positive definite (pivot %d)" % (what, b, int(status[b]) - 1))


class _SPDInverseLogdet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, check):
        be = get_backend()
        Lf, status = be.chol(X.contiguous())
        if check:
            _check_status(status, "spd_inverse_logdet")
        Linv = be.trinv(Lf)
        Xinv = be.ltl(Linv)
        logdet = 2.0 * torch.log(torch.diagonal(Lf, dim1=-2, dim2=-1)).sum(-1)
        ctx.save_for_backward(Xinv)
        ctx.mark_non_differentiable(Linv)
        return Xinv, logdet, Linv

    @staticmethod
    def backward(ctx, G_inv, g_ld, _g_linv):
        (Xinv,) = ctx.saved_tensors
        be = get_backend()
        gX = None
        if G_inv is not None:
            T = be.bmm64(Xinv, _sym(G_inv).contiguous())
            gX = -be.bmm64(T, Xinv)
        if g_ld is not None:
            t = g_ld[:, None, None] * Xinv
            gX = t if gX is None else gX + t
        return gX, None


def spd_inverse_logdet(X, check=True):
    """(X^-1, logdet X, L^-1) of a batch of SPD float64 matrices via blocked Cholesky (svgp_chol_f64)."""
    return _SPDInverseLogdet.apply(X, check)


class _SPDLogdet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, check):
        be = get_backend()
        Lf, status = be.chol(X.contiguous())
        if check:
            _check_status(status, "spd_logdet")
        ctx.save_for_backward(Lf)
        return 2.0 * torch.log(torch.diagonal(Lf, dim1=-2, dim2=-1)).sum(-1)

    @staticmethod
    def backward(ctx, g):
        (Lf,) = ctx.saved_tensors
        be = get_backend()
        Linv = be.trinv(Lf)
        Xinv = be.ltl(Linv)
        return g[:, None, None] * Xinv, None


def spd_logdet(X, check=True):
    return _SPDLogdet.apply(X, check)


# ------------------------------------------------------------------------------------------
# K1 as autograd functions (plain fp32 K)
# ------------------------------------------------------------------------------------------
class _KernelMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Fx, Fz, hyp, spec):
        be = get_backend()
        Fx32, Fz32, hyp32 = Fx.float().contiguous(), Fz.float().contiguous(), hyp.float().contiguous()
        ctx.save_for_backward(Fx32, Fz32, hyp32)
        ctx.spec = spec
        ctx.dtypes = (Fx.dtype, Fz.dtype, hyp.dtype)
        return be.kernel_fwd(spec, Fx32, Fz32, hyp32, tc=False).K

    @staticmethod
    def backward(ctx, G):
        Fx, Fz, hyp = ctx.saved_tensors
        be = get_backend()
        dFx, dFz, dhyp = be.kernel_bwd(ctx.spec, Fx, Fz, hyp, G.float().contiguous(),
                                       need_x=ctx.needs_input_grad[0], need_z=True)
        dx, dz, dh = ctx.dtypes
        return (dFx.to(dx) if dFx is not None else None, dFz.to(dz), dhyp.to(dh), None)


def kernel_matrix(Fx, Fz, hyp, spec):
    """K(Fx, Fz) (N x M fp32) for the two-block product kernel ``spec`` = (type_a, dim_a, type_b, dim_b)."""
    return _KernelMatrix.apply(Fx, Fz, hyp, spec)


class _KernelDiag(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Fx, Fy, hyp, spec):
        be = get_backend()
        Fx32, Fy32, hyp32 = Fx.float().contiguous(), Fy.float().contiguous(), hyp.float().contiguous()
        ctx.save_for_backward(Fx32, Fy32, hyp32)
        ctx.spec = spec
        ctx.dtypes = (Fx.dtype, Fy.dtype, hyp.dtype)
        return be.kernel_diag_fwd(spec, Fx32, Fy32, hyp32)

    @staticmethod
    def backward(ctx, g):
        Fx, Fy, hyp = ctx.saved_tensors
        dFx, dFy, dhyp = get_backend().kernel_diag_bwd(ctx.spec, Fx, Fy, hyp, g.float().contiguous())
        dx, dy, dh = ctx.dtypes
        return dFx.to(dx), dFy.to(dy), dhyp.to(dh), None


def kernel_diag(Fx, Fy, hyp, spec):
    return _KernelDiag.apply(Fx, Fy, hyp, spec)


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids):
        ctx.save_for_backward(ids)
        ctx.rows, ctx.dtype = table.shape[0], table.dtype
        return get_backend().gather_rows(table.float().contiguous(), ids)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        return get_backend().scatter_add_rows(g.float().contiguous(), ids, ctx.rows).to(ctx.dtype), None


def gather_rows(table, ids):
    """tf.gather(table, ids) with a duplicate-folding scatter-add backward."""
    return _GatherRows.apply(table, ids.contiguous())


# ------------------------------------------------------------------------------------------
# datapoint-sized contractions on a plain fp32 K (N x M)
# ------------------------------------------------------------------------------------------
class _Syrk(torch.autograd.Function):
    @staticmethod
    def forward(ctx, K, W):
        K32, W32 = K.float().contiguous(), W.float().contiguous()
        ctx.save_for_backward(K32, W32)
        ctx.dtypes = (K.dtype, W.dtype)
        return get_backend().syrk(Kop(K32), W32)

    @staticmethod
    def backward(ctx, G):
        K, W = ctx.saved_tensors
        be = get_backend()
        kop = Kop(K)
        Gs = (G + G.transpose(-1, -2)).contiguous()
        gK = be.scaled_gemm(kop, W, Gs) if ctx.needs_input_grad[0] else None
        gW = be.rowquad(kop, (0.5 * Gs).contiguous()) if ctx.needs_input_grad[1] else None
        dk, dw = ctx.dtypes
        return (gK.to(dk) if gK is not None else None, gW.to(dw) if gW is not None else None)


def syrk(K, W):
    """A[l] = sum_i W[i,l] k_i k_i^T  -> (L, M, M) float64."""
    return _Syrk.apply(K, W)


class _RowQuad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, K, S):
        K32 = K.float().contiguous()
        S64 = S.double().contiguous()
        ctx.save_for_backward(K32, S64)
        ctx.dtypes = (K.dtype, S.dtype)
        return get_backend().rowquad(Kop(K32), S64)

    @staticmethod
    def backward(ctx, g):
        K, S = ctx.saved_tensors
        be = get_backend()
        kop = Kop(K)
        g32 = g.float().contiguous()
        gK = be.scaled_gemm(kop, (2.0 * g32).contiguous(), _sym(S).contiguous()) if ctx.needs_input_grad[0] else None
        gS = be.syrk(kop, g32) if ctx.needs_input_grad[1] else None
        dk, ds = ctx.dtypes
        return (gK.to(dk) if gK is not None else None, gS.to(ds) if gS is not None else None)


def rowquad(K, S):
    """q[i,l] = k_i^T S_l k_i  -> (N, L) fp32;  S (L, M, M)."""
    return _RowQuad.apply(K, S)


class _KtMatmul(torch.autograd.Function):
    """V = X^T K  (L, M) float64 from K (N, M), X (N, L)."""

    @staticmethod
    def forward(ctx, K, X):
        K32, X32 = K.float().contiguous(), X.float().contiguous()
        ctx.save_for_backward(K32, X32)
        ctx.dtypes = (K.dtype, X.dtype)
        return get_backend().gemm_tn(Kop(K32), X32)

    @staticmethod
    def backward(ctx, G):
        K, X = ctx.saved_tensors
        be = get_backend()
        G32 = G.float().contiguous()
        gK = be.gemm_f32(X, G32) if ctx.needs_input_grad[0] else None              # (N,L)(L,M)
        gX = be.gemm_nn(Kop(K), G32) if ctx.needs_input_grad[1] else None           # K G^T
        dk, dx = ctx.dtypes
        return (gK.to(dk) if gK is not None else None, gX.to(dx) if gX is not None else None)


def kt_matmul(K, X):
    return _KtMatmul.apply(K, X)


class _KMatmul(torch.autograd.Function):
    """out = K W^T  (N, L) fp32 from K (N, M), W (L, M)."""

    @staticmethod
    def forward(ctx, K, Wm):
        K32, W32 = K.float().contiguous(), Wm.float().contiguous()
        ctx.save_for_backward(K32, W32)
        ctx.dtypes = (K.dtype, Wm.dtype)
        return get_backend().gemm_nn(Kop(K32), W32)

    @staticmethod
    def backward(ctx, G):
        K, Wm = ctx.saved_tensors
        be = get_backend()
        G32 = G.float().contiguous()
        gK = be.gemm_f32(G32, Wm) if ctx.needs_input_grad[0] else None              # (N,L)(L,M)
        gW = be.gemm_tn(Kop(K), G32) if ctx.needs_input_grad[1] else None           # G^T K  (L,M)
        dk, dw = ctx.dtypes
        return (gK.to(dk) if gK is not None else None, gW.to(dw) if gW is not None else None)


def k_matmul(K, Wm):
    return _KMatmul.apply(K, Wm)
