"""Per-tensor parity of the tcgen05 path against the streamlined float64 oracle for a list of shapes and SYRK
chain lengths.  Usage: python tests/probes/parity_probe.py N,M,L[,chunk_rows[,mm_chunk[,tc]]] ...   -> one JSON line per case."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refs  # noqa: E402
from oracle import svgp_streamlined as st  # noqa: E402
from svgp_vae_b200 import configs  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def main():
    for arg in sys.argv[1:]:
        v = [int(x) for x in arg.split(",")]
        N, M, L = v[:3]
        chunk = v[3] if len(v) > 3 else 0
        mmc = v[4] if len(v) > 4 else 0
        tc = bool(v[5]) if len(v) > 5 else True
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
        kw = dict(tc=tc, chunk_rows=chunk)
        if mmc:
            kw["mm_chunk"] = mmc
        r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), **kw)
        out = dict(N=N, M=M, L=L, chunk_rows=chunk, mm_chunk=mmc, tc=tc, p_m=rel(r1["p_m"], t0["p_m"]), p_v=rel(r1["p_v"], t0["p_v"]))
        for k in ("inside_elbo_recon", "inside_elbo_kl", "ce_term", "KL_term"):
            out[k] = abs(float(r1[k]) - float(g0[k])) / abs(float(g0[k]))
        for name, a, b in zip(["dy", "dnoise", "dZ", "dhyp"], gr0, g1):
            out[name] = rel(b, a)
        print(json.dumps({k: (float("%.3g" % x) if isinstance(x, float) else x) for k, x in out.items()}), flush=True)


if __name__ == "__main__":
    main()
