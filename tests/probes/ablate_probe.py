"""Error attribution for the tcgen05 path: selected backend calls of one step are replaced by float64 torch
computations ON THE SAME OPERAND PLANES (kop.value()), and the step's parity against the float64 oracle is
reported for every subset.  Usage: python tests/probes/ablate_probe.py N M L  -> JSON lines."""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refs  # noqa: E402
from oracle import svgp_streamlined as st  # noqa: E402
from svgp_vae_b200 import backend, configs  # noqa: E402

F64 = torch.float64


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def main():
    N, M, L = (int(x) for x in sys.argv[1:4])
    be = backend.get_backend()
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

    orig = dict(syrk=be.syrk, scaled=be.scaled_gemm, quad=be.rowquad, nn=be.gemm_nn, tn=be.gemm_tn, kfwd=be.kernel_fwd)
    Kx = {}

    def Kval(kop):
        if id(kop) not in Kx:
            Kx.clear()
            Kx[id(kop)] = kop.value().double()
        return Kx[id(kop)]

    state = {"n": 0, "which": ()}

    def x_syrk(kop, W, impl=0, chunk_rows=0):
        state["n"] += 1
        want = ("syrkA" in state["which"] and state["n"] == 1) or ("syrkG" in state["which"] and state["n"] == 2)
        if not want:
            return orig["syrk"](kop, W, impl=impl, chunk_rows=chunk_rows)
        return torch.einsum('il,ia,ib->lab', W.double(), Kval(kop), Kval(kop))

    def x_scaled(kop, W, G64, out=None, ndot=0, impl=0):
        K = Kval(kop)
        if "scaledP" in state["which"] and torch.is_tensor(G64):
            G64 = be.planes(G64)                       # exact product, but on the fp16 hi/lo planes of G
        G = G64 if torch.is_tensor(G64) else (G64.hi.double() + G64.lo.double()) * G64.inv[:G64.hi.shape[0], None, None].double()
        r = torch.zeros(K.shape[0], G.shape[2], dtype=F64, device=K.device)
        for t in range(G.shape[0]):
            r += W[:, t:t + 1].double() * (K @ G[t])
        r = r.float()
        if out is not None:
            out += r
            r = out
        if ndot:
            return r, torch.stack([((K @ G[t]) * K).sum(1) for t in range(ndot)], 1).float()
        return r

    def x_quad(kop, S64, tri=False, impl=0, out=None):
        K = Kval(kop)
        if tri:
            T = torch.einsum('ia,lca->ilc', K, S64)
            q = (T * T).sum(-1).float()
        else:
            q = torch.einsum('ia,lab,ib->il', K, S64, K).float()
        if out is not None:
            out.copy_(q)
            return out
        return q

    def x_nn(kop, Wm, impl=0):
        return (Kval(kop) @ Wm.double().t()).float()

    def x_tn(kop, Xm):
        return Xm.double().t() @ Kval(kop)

    exact = dict(syrk=x_syrk, scaled=x_scaled, quad=x_quad, nn=x_nn, tn=x_tn)
    names = ["syrkA", "syrkG", "scaled", "quad", "nn", "tn"]
    subsets = [(), tuple(names), ("syrkA",), ("scaled",), ("scaled", "scaledP"), ("syrkA", "syrkG", "scaled"), ("syrkA", "syrkG", "scaled", "scaledP")]
    for sub in subsets:
        state["n"], state["which"] = 0, sub
        be.syrk = exact["syrk"]
        be.scaled_gemm = exact["scaled"] if "scaled" in sub else orig["scaled"]
        Kx.clear()
        be.rowquad = exact["quad"] if "quad" in sub else orig["quad"]
        be.gemm_nn = exact["nn"] if "nn" in sub else orig["nn"]
        be.gemm_tn = exact["tn"] if "tn" in sub else orig["tn"]
        r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), tc=True)
        out = dict(exact=list(sub), p_m=rel(r1["p_m"], t0["p_m"]), p_v=rel(r1["p_v"], t0["p_v"]))
        for name, a, b in zip(["dy", "dnoise", "dZ", "dhyp"], gr0, g1):
            out[name] = rel(b, a)
        print(json.dumps({k: (float("%.3g" % v) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
    be.syrk, be.scaled_gemm, be.rowquad, be.gemm_nn, be.gemm_tn = (orig[k] for k in ("syrk", "scaled", "quad", "nn", "tn"))


if __name__ == "__main__":
    main()
