"""CPU simulation of operand formats for the tensor-core products of the SVGP step -- DESIGN TOOL, not product code.

Runs step.py's pass structure on the float64 CPU stand-in backend (tests/oracle_backend.py) with the three
N x M x M contractions (SYRK, row quads, scaled GEMM) and the two skinny products replaced by models of a
tensor-core operand format whose ACCUMULATION IS EXACT (integer MMA, int32 accumulators):

  i8:<bits>:<sa>x<sb>:<cut>   fixed-point operands relative to the per-row / per-column maximum (scale taken
                              along the reduction axis), cut into signed 8-bit slices (round-to-nearest digits);
                              slice pairs (t, u) with t + u <= cut are multiplied, the rest dropped.
  f16hl                       fp16 hi + lo operands (22 bits floating), exact accumulation (round 1's format
                              without its truncating TMEM accumulation)
  exact                       fp32 K_nm, float64 products (the SIMT path)

and reports per-tensor parity against the streamlined float64 oracle exactly like tests/probes/parity_probe.py.
Usage: python tools/numerics/sim_parity.py N M L model [model ...]

A model may override single ops: 'i8:0:4x4:3+syrk=...+syrkg=...+scaledA=...+scaledS=...+quad=...+nn=...' (syrk = the forward
SYRK, syrkg = its adjoint, scaledA / scaledS = the dA + dA^T / S - Kinv halves of pass D); the cut may be a list of digit pairs
('i8:0:4x4:00,01,02,10,11,12,20,21' = the eight pairs of three leading digits).  Environment switches:
  SIM_DIGITS=twos      the device's digits (d in [-128, 127]: the lower ones have mean -1/2) instead of round-to-nearest balanced
                       ones, everywhere; SIM_TWOS=op,op,... the same for single ops (this is what reproduced the GPU's 7e-4 of the
                       hyper-parameter gradients at M = 4096: DESIGN.md section 7)
  SIM_EPI=fma|int64    the scaled GEMM's own epilogue arithmetic (order accumulators -> fp32 by three fmas / one int64 conversion)
  SIM_CORR=1           with SIM_EPI: the expected value of the dropped order-4 pairs added before rounding (svgp_i8_pair_bias)
  SIM_SYRK_SYM=mirror|avg   how the SYRK result is symmetrised (lower triangle mirrored / both triangles averaged)
  SIM_SYRK_EXACT_PRODUCT=1, SIM_KNOISE, SIM_KMMNOISE   earlier experiments (section 7)
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refs  # noqa: E402
from oracle import svgp_streamlined as st  # noqa: E402
from oracle_backend import OracleBackend, kernel_value  # noqa: E402
from svgp_vae_b200 import backend, configs  # noqa: E402
from svgp_vae_b200.backend import Kop  # noqa: E402

F64 = torch.float64


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


TWOS_OPS = set(x for x in os.environ.get("SIM_TWOS", "").split(",") if x)      # ops whose digits are the device's (biased) ones
if os.environ.get("SIM_DIGITS") == "twos":
    TWOS_OPS = {"syrk", "syrkg", "scaledA", "scaledS", "scaled", "nn", "quad"}


class Fixed:
    """Fixed-point image of X: X ~= Xi * scale, |Xi| <= 127 * 2^(bits-7); scale constant along `axis` (the reduction axis)."""
    twos = False           # set per simulated op (SIM_TWOS): balanced round-to-nearest digits otherwise

    def __init__(self, X, axis, bits, nslices):
        self.bits, self.ns = bits, nslices
        mx = X.abs().amax(axis, keepdim=True).clamp_min(1e-300)
        top = 127.0 * 2.0 ** (bits - 7)
        self.scale = mx / top
        self.Xi = torch.round(X / self.scale)

    def upto(self, j):
        """Sum of the top j + 1 signed-digit slices (= Xi rounded to the grid of slice j)."""
        j = min(j, self.ns - 1)
        drop = self.bits + 1 - 8 * (j + 1)
        if drop <= 0:
            return self.Xi
        g = 2.0 ** drop
        if Fixed.twos:       # the device's digits: d in [-128, 127] (residue 128 -> -128 + carry), mean -0.5 for the lower ones
            return torch.floor(self.Xi / g + 0.5) * g
        return torch.round(self.Xi / g) * g

    def digit(self, t):
        return self.upto(t) - (self.upto(t - 1) if t > 0 else 0.0)

    def value(self):
        return self.Xi * self.scale


def sliced_matmul(A, B, cut, mm):
    """sum over slice pairs t + u <= cut of A_t B_u (integers, exact up to float64), mm(a, b) the contraction.
    cut may also be an explicit list of (t, u) pairs."""
    out = None
    if isinstance(cut, (list, tuple)):
        for t in sorted(set(t for t, _ in cut)):
            Bsum = sum(B.digit(u) for tt, u in cut if tt == t)
            term = mm(A.digit(t), Bsum)
            out = term if out is None else out + term
        return out
    for t in range(A.ns):
        if cut - t < 0:
            break
        term = mm(A.digit(t), B.upto(cut - t))
        out = term if out is None else out + term
    return out


class Model:
    """'default[+op=model...]', op in syrk / syrkg (the adjoint SYRK) / quad / scaled / nn / tn."""

    def __init__(self, text):
        self.text = text
        parts = text.split("+")
        self.sub = {}
        for q in parts[1:]:
            k, v = q.split("=")
            self.sub[k] = Model(v)
        p = parts[0].split(":")
        self.kind = p[0]
        if self.kind in ("sym", "symf"):          # SYRK of ONE quantised operand sqrt(w) K against itself, all slice pairs
            self.bits, self.sa = int(p[1]), int(p[2])
            self.sb, self.cut = self.sa, (int(p[3]) if len(p) > 3 else 99)
        if self.kind == "e8":                     # ONE two-sided equilibrated integer K (power-of-two row and column scales)
            self.bits = int(p[1])
            self.sa, self.sb = (int(x) for x in p[2].split("x"))
            self.cut = int(p[3])
        if self.kind == "i8":
            self.bits = int(p[1])
            self.sa, self.sb = (int(x) for x in p[2].split("x"))
            self.cut = int(p[3]) if "," not in p[3] and len(p[3]) < 2 else [(int(q[0]), int(q[1])) for q in p[3].split(",")]

    def of(self, op):
        return self.sub.get(op, self)

    def fixed(self, X, axis, which):
        ns = self.sa if which == "a" else self.sb
        return Fixed(X, axis, 8 * ns - 1 if self.bits == 0 else min(self.bits, 8 * ns - 1), ns)


def f16hl(X, axis=None):
    """fp16 hi + lo image of X with a power-of-two scale per matrix (max * s < 2^14), as round 1's svgp_split_f16."""
    mx = X.abs().max().clamp_min(1e-300)
    s = 2.0 ** torch.floor(torch.log2(16384.0 / mx) - 1e-9)
    Xs = X * s
    hi = Xs.to(torch.float16).to(F64)
    lo = (Xs - hi).to(torch.float16).to(F64)
    return (hi + lo) / s


class SimBackend(OracleBackend):
    name = "sim"

    def __init__(self, model):
        super().__init__()
        self.m = model
        self.nsyrk = 0

    def want_tc(self, N, M):
        return True

    def kernel_fwd(self, spec, Fx, Fz, hyp, tc=False):
        K = kernel_value(spec, Fx, Fz, hyp)
        noise = float(os.environ.get("SIM_KNOISE", "0"))      # relative error of the fp32 kernel evaluation (exp argument rounding)
        mmnoise = float(os.environ.get("SIM_KMMNOISE", "0"))  # the same for the fp32 K_mm
        if not tc and mmnoise > 0 and K.shape[0] == K.shape[1]:
            g = torch.Generator().manual_seed(12)
            E = torch.randn(K.shape, generator=g, dtype=F64)
            K = K * (1.0 + mmnoise * 0.5 * (E + E.t()))
        if tc and noise > 0:
            g = torch.Generator().manual_seed(11)
            K = K * (1.0 + noise * torch.randn(K.shape, generator=g, dtype=F64))
        kop = Kop(K.float())
        kop.sim = bool(tc)
        kop.cache = {}
        return kop

    @staticmethod
    def _kfixed(kop, m, axis):
        key = (axis, m.bits, m.sa, m.sb)
        if key not in kop.cache:
            kop.cache[key] = m.fixed(kop.K.double(), axis, "a" if axis == 1 else "b")
        return kop.cache[key]

    @staticmethod
    def _kequil(kop, m):
        """-> (Kint float64 integers, r (N,1), c (1,M)) with K ~= Kint r c 2^-b, r / c powers of two, |Kint| <= 2^b."""
        key = ("equil", m.bits)
        if key not in kop.cache:
            K = kop.K.double()
            b = m.bits if m.bits else 30
            r = 2.0 ** torch.ceil(torch.log2(K.abs().amax(1, keepdim=True).clamp_min(1e-300)))
            c = 2.0 ** torch.ceil(torch.log2((K / r).abs().amax(0, keepdim=True).clamp_min(1e-300)))
            Kint = torch.round(K / (r * c) * 2.0 ** b)
            kop.cache[key] = (Kint, r, c, b)
        return kop.cache[key]

    @staticmethod
    def _khl(kop):
        if "hl" not in kop.cache:
            kop.cache["hl"] = f16hl(kop.K.double())
        return kop.cache["hl"]

    def _kval(self, kop, m, axis=1):
        if not getattr(kop, "sim", False) or m.kind == "exact":
            return kop.K.double()
        if m.kind == "e8":
            Kint, r, c, b = self._kequil(kop, m)
            return Kint * r * c * 2.0 ** -b
        return self._khl(kop) if m.kind == "f16hl" else self._kfixed(kop, m, axis).value()

    # ---- reductions over datapoints ---------------------------------------------------------------------------
    def syrk(self, kop, W, impl=0, chunk_rows=0):
        self.nsyrk += 1
        m = self.m.of("syrk" if self.nsyrk == 1 else "syrkg")
        Fixed.twos = ("syrk" if self.nsyrk == 1 else "syrkg") in TWOS_OPS
        if not getattr(kop, "sim", False) or m.kind == "exact":
            return super().syrk(kop, W)
        if m.kind == "f16hl":
            K = self._khl(kop)
            return torch.stack([(f16hl((W[:, l:l + 1].float() * K.float()).double())).t() @ K for l in range(W.shape[1])])
        out = []
        if m.kind == "e8":
            Kint, r, c, b = self._kequil(kop, m)
            Kf = Fixed.__new__(Fixed)
            Kf.bits, Kf.ns, Kf.Xi, Kf.scale = 31, m.sb, Kint * 2.0, None          # digits of the shared integer (31-bit container)
            for l in range(W.shape[1]):
                wp = (W[:, l:l + 1].double() * r * r)                               # w r^2 per datapoint
                wn = (wp / wp.abs().max()).float()
                V = (wn * Kint.float()).double()                                    # fp32 product, |V| <= 2^b
                Vf = Fixed.__new__(Fixed)
                Vf.bits, Vf.ns, Vf.Xi, Vf.scale = 31, m.sa, torch.round(V * 2.0), None
                rr = sliced_matmul(Vf, Kf, m.cut, lambda a_, b_: a_.t() @ b_) / 4.0
                out.append(rr * wp.abs().max() * (c.t() * c) * 2.0 ** (-2 * b))
            return torch.stack(out)
        if m.kind in ("sym", "symf"):
            K = kop.K.double()
            for l in range(W.shape[1]):
                w = W[:, l:l + 1].double()
                U = (w.abs().sqrt().float() * K.float()).double()
                if m.kind == "sym":
                    Uf = m.fixed(U, 0, "a")
                    if m.cut < 2 * (m.sa - 1):
                        sg = torch.sign(w)
                        r = None
                        for t in range(m.sa):
                            if m.cut - t < 0:
                                break
                            term = (Uf.digit(t) * sg).t() @ Uf.upto(m.cut - t)
                            r = term if r is None else r + term
                        out.append(r * Uf.scale.t() * Uf.scale)
                        continue
                    Uq = Uf.value()
                else:                                   # floating: round to `bits` significand bits
                    mant, ex = torch.frexp(U)
                    Uq = torch.ldexp(torch.round(mant * 2.0 ** m.bits) / 2.0 ** m.bits, ex.to(torch.int32))
                out.append((Uq * torch.sign(w)).t() @ Uq)
            return torch.stack(out)
        Kc = self._kfixed(kop, m, 0)
        Kv = Kc.value()
        for l in range(W.shape[1]):
            if os.environ.get("SIM_SYRK_EXACT_PRODUCT", "0") == "1":     # 64-bit integer product in the transform (one rounding, onto the grid)
                V = W[:, l:l + 1].double() * Kv
            else:
                V = (W[:, l:l + 1].float() * Kv.float()).double()      # fp32 product in the transform
            Vf = m.fixed(V, 0, "a")
            r = sliced_matmul(Vf, Kc, m.cut, lambda a, b: a.t() @ b)
            r = r * Vf.scale.t() * Kc.scale                              # V^T K: weighted operand = row index
            sym = os.environ.get("SIM_SYRK_SYM", "")                     # how the kernel turns it into a symmetric matrix
            if sym == "mirror":                                          # lower triangle of K^T V (weighted = column index), mirrored: the GPU kernel
                rt = r.t()
                r = torch.tril(rt) + torch.tril(rt, -1).t()
            elif sym == "avg":                                           # both triangles computed, averaged
                r = 0.5 * (r + r.t())
            out.append(r)
        return torch.stack(out)

    def gemm_tn(self, kop, X):
        return X.to(F64).t() @ self._kval(kop, self.m.of("tn"), 0)

    # ---- reductions over inducing points ----------------------------------------------------------------------
    def _kg(self, kop, G, m):
        """K_nm @ G for one (M, C) float64 matrix G in the simulated format (float64 result of exact accumulation)."""
        if not getattr(kop, "sim", False) or m.kind == "exact":
            return kop.K.double() @ G
        if m.kind == "f16hl":
            return self._khl(kop) @ f16hl(G)
        if m.kind == "e8":
            Kint, r, c, b = self._kequil(kop, m)
            Gp = G * c.t()                                             # rows of G scaled by the column scales of K
            Gf = Fixed(Gp, 0, 8 * m.sb - 1, m.sb)
            Kf = Fixed.__new__(Fixed)
            Kf.bits, Kf.ns, Kf.Xi, Kf.scale = 31, m.sa, Kint * 2.0, None
            rr = sliced_matmul(Kf, Gf, m.cut, lambda a_, b_: a_ @ b_) / 2.0
            return rr * r * 2.0 ** -b * Gf.scale
        Gf = m.fixed(G, 0, "b")                                        # scale per output column
        Kr = self._kfixed(kop, m, 1)
        r = sliced_matmul(Kr, Gf, m.cut, lambda a, b: a @ b)
        return r * Kr.scale * Gf.scale

    def gemm_nn(self, kop, Wm):
        Fixed.twos = "nn" in TWOS_OPS
        return self._kg(kop, Wm.to(F64).t(), self.m.of("nn")).float()

    def rowquad(self, kop, S64, tri=False, impl=0, out=None):
        m = self.m.of("quad")
        Fixed.twos = "quad" in TWOS_OPS
        qs = []
        for l in range(S64.shape[0]):
            if tri:
                T = self._kg(kop, S64[l].t(), m).float()               # fp32 epilogue: sum of squares
                qs.append((T * T).sum(-1, dtype=torch.float32))
            else:
                T = self._kg(kop, S64[l], m).float()
                qs.append((T * self._kval(kop, m).float()).sum(-1, dtype=torch.float32))
        q = torch.stack(qs, 1)
        if out is not None:
            out.copy_(q)
            return out
        return q

    def scaled_gemm(self, kop, W, G64, out=None, ndot=0, impl=0):
        N, M = kop.K.shape
        r = torch.zeros(N, G64.shape[2], dtype=torch.float32)
        dots = []
        L2 = G64.shape[0]
        for s in range(L2):
            m = self.m.of("scaledA" if s < ndot else "scaledS")
            Fixed.twos = ("scaledA" if s < ndot else "scaledS") in TWOS_OPS
            if m is self.m:
                m = self.m.of("scaled")
            if os.environ.get("SIM_EPI") and getattr(kop, "sim", False) and m.kind == "i8" and m.cut == 3:
                # the kernel's own epilogue arithmetic (scaled8_chunk): four int32 order accumulators, each converted to fp32,
                # a three-fma recombination, the column scale, an fma into the running sum
                Gf = m.fixed(G64[s], 0, "b")
                Kr = self._kfixed(kop, m, 1)
                acc = [sum(Kr.digit(t) @ Gf.digit(o - t) for t in range(o + 1)) for o in range(4)]
                unit = [2.0 ** -(8 * (3 - o)) for o in range(4)]          # digit(t) carries its weight: strip it -> plain integers
                top = 2.0 ** (Kr.bits + 1 - 8) * 2.0 ** (Gf.bits + 1 - 8)
                a = [(acc[o] / (top * 2.0 ** (-8 * o))) for o in range(4)]
                if os.environ.get("SIM_CORR"):
                    # expectation of the dropped order-4 pairs (1,3) (2,2) (3,1) of digits d = e - 1/2 (e symmetric): the device's
                    # rank-1 correction  -(SK_123[i] + SG_123[a]) / 512 - 3 M / 1024  in units of the order-3 accumulator
                    u = lambda F, t: F.digit(t) / 2.0 ** (F.bits + 1 - 8 * (t + 1))        # plain integer digits
                    sk = sum(u(Kr, t).sum(1, keepdim=True) for t in (1, 2, 3))
                    sg = sum(u(Gf, t).sum(0, keepdim=True) for t in (1, 2, 3))
                    a[3] = a[3] - (sk + sg) / 512.0 - 3.0 * Kr.Xi.shape[1] / 1024.0
                af = [x.float().double() for x in a]
                mode = os.environ["SIM_EPI"]
                if mode == "int64":                                      # exact integer recombination, ONE rounding
                    tv = (((a[0] * 256 + a[1]) * 256 + a[2]) * 256 + a[3]).float().double()
                else:
                    tv = (af[2] * 256.0 + af[3]).float().double()
                    tv = (af[1] * 65536.0 + tv).float().double()
                    tv = (af[0] * 16777216.0 + tv).float().double()
                tv = (tv * Gf.scale.float().double()).float().double()
                wgt = (W[:, s:s + 1].double() * Kr.scale * top * 2.0 ** -24).float().double()
                r = (wgt * tv + r.double()).float()
                T = (tv * (Kr.scale * top * 2.0 ** -24)).float()
            else:
                T = self._kg(kop, G64[s], m).float()                       # one rounding to fp32 out of the accumulators
                r += W[:, s:s + 1].float() * T                              # fp32 running sum in the epilogue registers
            if s < ndot:
                dots.append((T * self._kval(kop, m).float()).sum(-1, dtype=torch.float32))
        if out is not None:
            out += r
            r = out
        if ndot:
            return r, torch.stack(dots, 1)
        return r


def main():
    N, M, L = (int(x) for x in sys.argv[1:4])
    torch.manual_seed(0)
    cfg = configs.sweep_inputs(N, M, L)
    o, s, op, sp = refs.make_pair("sweep", cfg, "cpu")
    X, y, nz = cfg["aux"].double(), cfg["y"].double().requires_grad_(True), cfg["noise"].double().requires_grad_(True)
    for t in op:
        t.requires_grad_(True)
    Z = o.inducing_index_points
    t0c = time.time()
    t0 = st.streamlined_terms(o.kernel_matrix(X, Z), o.kernel_matrix(Z, Z), o.kernel_matrix(X, X, diag_only=True), y, nz,
                              cfg["ctor"]["N_train"], cfg["ctor"]["jitter"])
    g0 = st.glue_from_terms(t0, float(N), cfg["ctor"]["N_train"])
    gm, gv = refs.upstream(tuple(y.shape))
    J0 = g0["KL_term"] + (gm * t0["p_m"]).sum() + (gv * t0["p_v"]).sum()
    gr0 = torch.autograd.grad(J0, [y, nz] + op)
    ref = {k: t0[k].detach() for k in ("p_m", "p_v")}
    refs0 = {k: float(g0[k]) for k in ("inside_elbo_recon", "inside_elbo_kl", "ce_term", "KL_term")}
    del t0, g0, J0
    print("# oracle %.1fs" % (time.time() - t0c), flush=True)
    for text in sys.argv[4:]:
        t1 = time.time()
        old = backend.set_backend_for_tests(SimBackend(Model(text)))
        try:
            r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"], cfg["y"], cfg["noise"])
        finally:
            backend.set_backend_for_tests(old)
        out = dict(N=N, M=M, L=L, model=text, p_m=rel(r1["p_m"], ref["p_m"]), p_v=rel(r1["p_v"], ref["p_v"]))
        for k in refs0:
            out[k] = abs(float(r1[k]) - refs0[k]) / abs(refs0[k])
        for name, a, b in zip(["dy", "dnoise", "dZ", "dhyp"], gr0, g1):
            out[name] = rel(b, a)
        out["sec"] = round(time.time() - t1, 1)
        print(json.dumps({k: (float("%.3g" % x) if isinstance(x, float) else x) for k, x in out.items()}), flush=True)


if __name__ == "__main__":
    main()
