"""Which of the remaining fp16 / fp32 pieces limits parity on the integer tensor-core path?  Selected backend calls of one
step are replaced by float64 torch computations on the same operands and the step's parity is reported per subset.
Usage: python tests/probes/ablate_i8_probe.py N M L [subset,subset,...]   (subsets of: scaledS tn quad nn f32 kbwd)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refs  # noqa: E402
from oracle import svgp_streamlined as st  # noqa: E402
from oracle_backend import OracleBackend  # noqa: E402
from svgp_vae_b200 import backend, configs  # noqa: E402

F64 = torch.float64


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def main():
    N, M, L = (int(x) for x in sys.argv[1:4])
    subsets = [tuple(x for x in s.split("+") if x) for s in (sys.argv[4].split(",") if len(sys.argv) > 4 else [""])]
    be = backend.get_backend()
    ob = OracleBackend()
    cfg = configs.sweep_inputs(N, M, L)
    o, s, op, sp = refs.make_pair("sweep", cfg, "cuda")
    X, y, nz = cfg["aux"].double(), cfg["y"].double().requires_grad_(True), cfg["noise"].double().requires_grad_(True)
    for t in op:
        t.requires_grad_(True)
    Z = o.inducing_index_points
    t0 = st.streamlined_terms(o.kernel_matrix(X, Z), o.kernel_matrix(Z, Z), o.kernel_matrix(X, X, diag_only=True), y, nz,
                              cfg["ctor"]["N_train"], cfg["ctor"]["jitter"])
    g0 = st.glue_from_terms(t0, float(N), cfg["ctor"]["N_train"])
    gm, gv = refs.upstream(tuple(y.shape))
    J0 = g0["KL_term"] + (gm * t0["p_m"]).sum() + (gv * t0["p_v"]).sum()
    gr0 = torch.autograd.grad(J0, [y, nz] + op)
    orig = dict(scaled=be.scaled_gemm, tn=be.gemm_tn, quad=be.rowquad, nn=be.gemm_nn, f32=be.gemm_f32, kbwd=be.kernel_bwd)
    state = {"which": ()}

    from oracle_backend import kernel_value
    Ktrue = {}

    def Kval(kop, which="r"):
        if "ktrue" in state["which"]:                        # the float64 kernel values themselves instead of the operand planes
            if "K" not in Ktrue:
                Ktrue["K"] = kernel_value((1, 4, 1, 4), cfg["aux"].cuda(), torch.as_tensor(cfg["ctor"]["initial_inducing_points"]).cuda(),
                                          torch.ones(4, device="cuda"))
            return Ktrue["K"]
        return kop.value_i8(which) if kop.i8 else kop.value().double()

    orig_syrk, orig_si8 = be.syrk, be.scaled_gemm_i8

    def x_syrk(kop, W, impl=0, chunk_rows=0):
        if "syrk" not in state["which"]:
            return orig_syrk(kop, W, impl=impl, chunk_rows=chunk_rows)
        K = Kval(kop, "r" if "kr" in state["which"] else "c")      # "kr": the SAME dequantised K as every other product
        return torch.stack([(K * W[:, l:l + 1].double()).t() @ K for l in range(W.shape[1])])

    def x_si8(kop, W, G, out=None, ndot=0):
        which = state["which"]
        if ndot and ("dotsA" in which or "outA" in which):
            # only ONE of the kernel's two outputs from float64: "dotsA" = the k-dots, "outA" = the weighted sums
            k_out, k_dots = orig_si8(kop, W, G, out=(out.clone() if out is not None else None), ndot=ndot)
            state["which"] = tuple(w for w in which if w not in ("dotsA", "outA")) + ("scaledA",)
            try:
                f_out, f_dots = x_si8(kop, W, G, out=out, ndot=ndot)
            finally:
                state["which"] = which
            return (f_out if "outA" in which else k_out), (f_dots if "dotsA" in which else k_dots)
        if "scaledA" not in state["which"]:
            return orig_si8(kop, W, G, out=out, ndot=ndot)
        K = Kval(kop)
        Gv = G.value()
        r = torch.zeros(K.shape[0], G.R, dtype=F64, device=K.device)
        dots = torch.zeros(K.shape[0], max(ndot, 1), dtype=F64, device=K.device)
        for t in range(G.B):
            T = K @ Gv[t].t()
            r += T if W is None else W[:, t:t + 1].double() * T
            if t < ndot:
                dots[:, t] = (T * K).sum(1)
        if out is not None:
            out += r.float()
            r = out
        else:
            r = r.float()
        return (r, dots[:, :ndot].float()) if ndot else r

    be.syrk, be.scaled_gemm_i8 = x_syrk, x_si8

    def x_scaled(kop, W, G64, out=None, ndot=0, impl=0):
        if "scaledS" not in state["which"] or ndot:
            return orig["scaled"](kop, W, G64, out=out, ndot=ndot, impl=impl)
        K = Kval(kop)
        G = G64 if torch.is_tensor(G64) else (G64.hi.double() + G64.lo.double()) * G64.inv[:G64.hi.shape[0], None, None].double()
        r = torch.zeros(K.shape[0], G.shape[2], dtype=F64, device=K.device)
        for t in range(G.shape[0]):
            r += W[:, t:t + 1].double() * (K @ G[t])
        if out is not None:
            out += r.float()
            return out
        return r.float()

    def x_tn(kop, Xm):
        if "tn" not in state["which"]:
            return orig["tn"](kop, Xm)
        return Xm.double().t() @ Kval(kop)

    def x_quad(kop, S64, tri=False, impl=0, out=None):
        if "quad" not in state["which"]:
            return orig["quad"](kop, S64, tri=tri, impl=impl, out=out)
        K = Kval(kop)
        S = S64 if torch.is_tensor(S64) else (S64.hi.double() + S64.lo.double()) * S64.inv[:S64.hi.shape[0], None, None].double()
        q = torch.stack([((K @ S[l].t()) ** 2).sum(1) if tri else ((K @ S[l]) * K).sum(1) for l in range(S.shape[0])], 1).float()
        if out is not None:
            out.copy_(q)
            return out
        return q

    def x_nn(kop, Wm, impl=0):
        if "nn" not in state["which"]:
            return orig["nn"](kop, Wm, impl=impl)
        return (Kval(kop) @ Wm.double().t()).float()

    def x_f32(A, B, out=None):
        if "f32" not in state["which"]:
            return orig["f32"](A, B, out=out)
        r = (A.double() @ B.double())
        if out is not None:
            out.copy_((out.double() + r).float())
            return out
        return r.float()

    def x_kbwd(spec, Fx, Fz, hyp, G, need_x=True, need_z=True):
        if "kbwd" not in state["which"]:
            return orig["kbwd"](spec, Fx, Fz, hyp, G, need_x=need_x, need_z=need_z)
        outs = []
        for r0 in range(0, Fx.shape[0], 8192):                      # float64 autograd through the oracle kernel formulas
            outs.append(ob.kernel_bwd(spec, Fx[r0:r0 + 8192], Fz, hyp, G[r0:r0 + 8192], need_x=need_x, need_z=need_z))
        dFx = torch.cat([o_[0] for o_ in outs]) if need_x else None
        return dFx, sum(o_[1] for o_ in outs), sum(o_[2] for o_ in outs)

    be.scaled_gemm, be.gemm_tn, be.rowquad, be.gemm_nn, be.gemm_f32, be.kernel_bwd = x_scaled, x_tn, x_quad, x_nn, x_f32, x_kbwd
    for sub in subsets:
        state["which"] = sub
        r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), tc=True)
        out = dict(N=N, M=M, L=L, exact="+".join(sub), p_m=rel(r1["p_m"], t0["p_m"]), p_v=rel(r1["p_v"], t0["p_v"]))
        for name, a, b in zip(["dy", "dnoise", "dZ", "dhyp"], gr0, g1):
            out[name] = rel(b, a)
        print(json.dumps({k: (float("%.3g" % x) if isinstance(x, float) else x) for k, x in out.items()}), flush=True)


if __name__ == "__main__":
    main()
