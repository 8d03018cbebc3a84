"""float64 CPU stand-in for svgp_vae_b200.backend.CudaBackend -- TEST INFRASTRUCTURE ONLY.

Implements every backend primitive with plain torch on the CPU (kernel formulas from
oracle/tfp_kernels.py) so that the *host logic* of the product (ops.py autograd formulas, step.py
pass structure, svgp.py method composition, dist sharding) can be checked against the literal
oracle without a GPU.  Injected with ``backend.set_backend_for_tests``; the product never imports it.
On the GPU box the same primitives are compared one by one against the CUDA kernels.
"""
import torch

from oracle import tfp_kernels as tfk
from svgp_vae_b200.backend import Kop

K_NONE, K_SE, K_EXPSIN, K_LINEAR, K_COSINE = 0, 1, 2, 3, 4
F64 = torch.float64


def _factor(t, x, z, amp, length, pairwise):
    if t == K_NONE:
        return None
    if t == K_SE:
        k = tfk.ExponentiatedQuadratic(amp, length)
    elif t == K_EXPSIN:
        k = tfk.ExpSinSquared(amp, length, 2 * torch.pi)
    else:
        k = tfk.Linear()
    out = k.matrix(x, z) if pairwise else k.apply(x, z)
    if t == K_COSINE:
        nx, nz = x.norm(dim=1), z.norm(dim=1)
        out = out / (nx[:, None] * nz[None, :]) if pairwise else out / (nx * nz)
    return out


def kernel_value(spec, Fx, Fz, hyp, pairwise=True):
    ta, da, tb, db = spec
    Fx, Fz, hyp = Fx.to(F64), Fz.to(F64), hyp.to(F64)
    ka = _factor(ta, Fx[:, :da], Fz[:, :da], hyp[0], hyp[1], pairwise)
    kb = _factor(tb, Fx[:, da:da + db], Fz[:, da:da + db], hyp[2], hyp[3], pairwise)
    if ka is None:
        return kb
    return ka if kb is None else ka * kb


class OracleBackend:
    name = "oracle-cpu-f64"

    def __init__(self):
        self.launches = 0

    def want_tc(self, N, M):
        return False

    # K1
    def kernel_fwd(self, spec, Fx, Fz, hyp, tc=False):
        return Kop(kernel_value(spec, Fx, Fz, hyp).float())

    def kernel_fwd_f64(self, spec, Fx, Fz, hyp):
        return kernel_value(spec, Fx, Fz, hyp)

    def kernel_bwd(self, spec, Fx, Fz, hyp, G, need_x=True, need_z=True):
        with torch.enable_grad():
            Fx_, Fz_, h_ = (t.detach().to(F64).requires_grad_(True) for t in (Fx, Fz, hyp))
            K = kernel_value(spec, Fx_, Fz_, h_)
            gx, gz, gh = torch.autograd.grad(K, [Fx_, Fz_, h_], G.to(F64), allow_unused=True)
        z = lambda g, ref: torch.zeros_like(ref) if g is None else g
        return (z(gx, Fx_).float() if need_x else None, z(gz, Fz_) if need_z else None, z(gh, h_))

    def kernel_diag_fwd(self, spec, Fx, Fy, hyp):
        return kernel_value(spec, Fx, Fy, hyp, pairwise=False).float()

    def kernel_diag_bwd(self, spec, Fx, Fy, hyp, g):
        with torch.enable_grad():
            Fx_, Fy_, h_ = (t.detach().to(F64).requires_grad_(True) for t in (Fx, Fy, hyp))
            k = kernel_value(spec, Fx_, Fy_, h_, pairwise=False)
            gx, gy, gh = torch.autograd.grad(k, [Fx_, Fy_, h_], g.to(F64), allow_unused=True)
        z = lambda gg, ref: torch.zeros_like(ref) if gg is None else gg
        return z(gx, Fx_).float(), z(gy, Fy_).float(), z(gh, h_)

    def gather_rows(self, table, ids):
        return table[ids].float()

    def scatter_add_rows(self, g, ids, rows):
        out = torch.zeros((rows, g.shape[1]), dtype=F64)
        out.index_add_(0, ids, g.to(F64))
        return out

    # GEMM class (all in float64 internally)
    def syrk(self, kop, W, impl=0, chunk_rows=0):
        K = kop.value().to(F64)
        return torch.einsum('il,ia,ib->lab', W.to(F64), K, K)

    def gemm_tn(self, kop, X):
        return X.to(F64).t() @ kop.value().to(F64)

    def gemm_nn(self, kop, Wm):
        return (kop.value().to(F64) @ Wm.to(F64).t()).float()

    def rowquad(self, kop, S64, tri=False, impl=0, out=None):
        K = kop.value().to(F64)
        if tri:
            T = torch.einsum('ia,lca->ilc', K, S64)
            q = (T * T).sum(-1).float()
        else:
            q = torch.einsum('ia,lab,ib->il', K, S64, K).float()
        if out is not None:
            out.copy_(q)
            return out
        return q

    def scaled_gemm(self, kop, W, G64, out=None, ndot=0, impl=0):
        K = kop.value().to(F64)
        r = torch.einsum('il,ia,lac->ic', W.to(F64), K, G64).float()
        if out is not None:
            out += r
            r = out
        if ndot:
            return r, torch.einsum('ia,lab,ib->il', K, G64[:ndot], K).float()
        return r

    def gemm_f32(self, A, B, out=None):
        r = (A.to(F64) @ B.to(F64)).float()
        if out is not None:
            out += r
            return out
        return r

    # K3
    def chol(self, X):
        Lf, info = torch.linalg.cholesky_ex(X)
        return Lf, info.to(torch.int32)

    def trinv(self, Lf):
        eye = torch.eye(Lf.shape[-1], dtype=Lf.dtype).expand_as(Lf)
        return torch.linalg.solve_triangular(Lf, eye, upper=False)

    def ltl(self, T):
        return T.transpose(-1, -2) @ T

    def bmm64(self, A, B, transA=False, transB=False):
        A = A.transpose(-1, -2) if transA else A
        B = B.transpose(-1, -2) if transB else B
        return (A @ B).contiguous()

    # K4 row terms
    def rowstats(self, y, noise, kappa):
        p = torch.where(noise == 0, torch.zeros_like(noise), 1.0 / torch.where(noise == 0, torch.ones_like(noise), noise))
        py = p * y
        p64, y64 = p.to(F64), y.to(F64)
        sums = torch.stack([(p64 * kappa.to(F64)[:, None]).sum(0), (p64 * y64 * y64).sum(0), torch.log(noise.to(F64)).sum(0)])
        return p, py, sums

    def predictive(self, kappa, h, q1, p, clip=None):
        raw = kappa[:, None] - h[:, None] + q1
        L = q1.shape[1]
        if not clip:
            return raw, torch.zeros(L, dtype=F64), None
        pv = raw.clamp(clip[0], clip[1])
        mask = (pv != raw).to(torch.uint8)
        clipsum = (p.to(F64) * (pv.to(F64) - raw.to(F64))).sum(0)
        return pv, clipsum, mask

    def rowterms_bwd_pre(self, g_pv, g_pm, p, y, clip=None):
        G_q1 = torch.zeros_like(y) if g_pv is None else g_pv.clone()
        G_p_clip = None
        if clip:
            mask, pv, kappa, h, q1raw, gce = clip
            m = mask.bool()
            pv_raw = kappa[:, None] - h[:, None] + q1raw
            G_q1 = torch.where(m, 0.5 * gce[None, :] * p, G_q1)
            G_p_clip = -0.5 * gce[None, :] * (pv - pv_raw)
        return (G_q1.contiguous(), torch.cat([p, 2.0 * G_q1], 1).contiguous(), torch.cat([p * y, g_pm], 1).contiguous(), G_p_clip,
                G_q1.sum(1))

    def rowterms_bwd_post(self, y, noise, p, kappa, kGk, G_py, gsums, G_p_clip, G_kappa):
        gs = gsums.float()
        G_p = 0.5 * kGk + y * G_py + kappa[:, None] * gs[0][None, :] + (y * y) * gs[1][None, :]
        if G_p_clip is not None:
            G_p = G_p + G_p_clip
        G_y = p * G_py + 2.0 * p * y * gs[1][None, :]
        nz = noise != 0
        safe = torch.where(nz, noise, torch.ones_like(noise))
        G_noise = torch.where(nz, -p * p * G_p, torch.zeros_like(p)) + gs[2][None, :] / safe
        return G_y, G_noise, G_kappa + (p * gs[0][None, :]).sum(1)
