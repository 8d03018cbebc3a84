"""Where the tcgen05 path loses accuracy: every tensor-core call of one real step is captured and re-computed in
float64 from the SAME operand planes (so only the accumulation differs), and from the unsplit float64 operands
(so the fp16 hi/lo split shows up).  Usage: python tests/probes/accum_probe.py N M L [chunk_rows] -> JSON lines."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refs  # noqa: E402
from svgp_vae_b200 import backend, configs  # noqa: E402


def stats(name, got, ref, **kw):
    got, ref = got.double(), ref.double()
    err = got - ref
    den = ref.abs().max()
    out = dict(op=name, max_rel=float(err.abs().max() / den), rms_rel=float(err.pow(2).mean().sqrt() / den),
               mean_signed_rel=float(err.mean() / den), ref_rms_rel=float(ref.pow(2).mean().sqrt() / den),
               shrink=float(-(err * ref).sum() / (ref * ref).sum()), **kw)     # least-squares beta of got = (1 - beta) ref
    print(json.dumps({k: (float("%.3g" % v) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)


def main():
    N, M, L = (int(x) for x in sys.argv[1:4])
    chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    be = backend.get_backend()
    cfg = configs.sweep_inputs(N, M, L)
    o, s, op, sp = refs.make_pair("sweep", cfg, "cuda")
    cap = {}
    orig_scaled, orig_syrk, orig_quad = be.scaled_gemm, be.syrk, be.rowquad

    def scaled(kop, W, G64, out=None, ndot=0, impl=0):
        r = orig_scaled(kop, W, G64, out=out, ndot=ndot, impl=impl)
        if ndot:
            cap["scaled"] = (kop, W.clone(), G64.clone() if torch.is_tensor(G64) else G64, r[0].clone(), r[1].clone(), ndot)
        return r

    def syrk(kop, W, impl=0, chunk_rows=0):
        r = orig_syrk(kop, W, impl=impl, chunk_rows=chunk_rows)
        if W.shape[1] == L:
            cap.setdefault("syrk", []).append((kop, W.clone(), r.clone(), chunk_rows))
        return r

    def quad(kop, S64, tri=False, impl=0, out=None):
        r = orig_quad(kop, S64, tri=tri, impl=impl, out=out)
        if (S64.hi if isinstance(S64, backend.Planes) else S64).shape[0] == L:
            cap["quad"] = (kop, S64.clone(), tri, r.clone())
        return r

    be.scaled_gemm, be.syrk, be.rowquad = scaled, syrk, quad
    refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), tc=True, chunk_rows=chunk)
    be.scaled_gemm, be.syrk, be.rowquad = orig_scaled, orig_syrk, orig_quad

    kop = cap["scaled"][0]
    K = kop.value().double()                                   # exact value of the operand planes
    for i, (_, W, A_tc, cr) in enumerate(cap["syrk"]):
        A_ref = torch.einsum('il,ia,ib->lab', W.double(), K, K)
        stats("syrk[%d]" % i, A_tc, A_ref, chunk_rows=cr, positive_weights=bool((W >= 0).all()))
        for c2 in (256, 512, 1024, 2048, 4096):
            stats("syrk[%d]" % i, orig_syrk(kop, W, chunk_rows=c2), A_ref, chunk_rows=c2)
    _, S64, tri, q_tc = cap["quad"]
    pl = be.planes(S64)
    Sp = (pl.hi.double() + pl.lo.double()) * pl.inv[:L, None, None].double()
    for nm, Sx in (("planes", Sp), ("float64", S64)):
        if tri:
            T = torch.einsum('ia,lca->ilc', K, Sx)
            q_ref = (T * T).sum(-1)
        else:
            q_ref = torch.einsum('ia,lab,ib->il', K, Sx, K)
        stats("rowquad vs " + nm, q_tc, q_ref, tri=bool(tri))
    _, W, G64, out_tc, dots_tc, ndot = cap["scaled"]
    pl = be.planes(G64)
    nb = G64.shape[0]
    Gp = (pl.hi.double() + pl.lo.double()) * pl.inv[:nb, None, None].double()
    Fx = s._features(cfg["aux"].cuda(), False).float().contiguous()
    Fz = s._features(s.inducing_index_points, True).float().contiguous()
    hyp = s._hyp().float().contiguous()
    # per matrix family (with the default ordering: [dA + dA^T (L) ; S - Kinv (L)]), against the same planes
    for fam, sl in (("dA+dA^T", slice(0, L)), ("S-Kinv", slice(L, nb))):
        o_tc = orig_scaled(kop, W[:, sl].contiguous(), G64[sl].contiguous())
        o_ref = torch.zeros(N, M, dtype=torch.float64, device="cuda")
        for t in range(sl.start, sl.stop):
            o_ref += W[:, t:t + 1].double() * (K @ Gp[t])
        stats("scaled_gemm family " + fam + " vs planes", o_tc, o_ref)
        _, dZ_a, _ = be.kernel_bwd(s._spec(), Fx, Fz, hyp, o_tc.float().contiguous(), need_x=False)
        _, dZ_b, _ = be.kernel_bwd(s._spec(), Fx, Fz, hyp, o_ref.float().contiguous(), need_x=False)
        stats("dZ (K_nm path) family " + fam, dZ_a, dZ_b)
    for nm, Gx in (("planes", Gp), ("float64", G64)):
        out_ref = torch.zeros(N, M, dtype=torch.float64, device="cuda")
        for t in range(nb):
            out_ref += W[:, t:t + 1].double() * (K @ Gx[t])
        stats("scaled_gemm out vs " + nm, out_tc, out_ref)
        dots_ref = torch.stack([((K @ Gx[t]) * K).sum(1) for t in range(ndot)], 1)
        stats("scaled_gemm dots vs " + nm, dots_tc, dots_ref)
        # what the error of dK_nm does to the inducing-point gradient (K_nm path only)
        _, dZ_tc, dh_tc = be.kernel_bwd(s._spec(), Fx, Fz, hyp, out_tc.float().contiguous(), need_x=False)
        _, dZ_ref, dh_ref = be.kernel_bwd(s._spec(), Fx, Fz, hyp, out_ref.float().contiguous(), need_x=False)
        stats("dZ (K_nm path) from dK_nm vs " + nm, dZ_tc, dZ_ref)
        stats("dhyp (K_nm path) from dK_nm vs " + nm, dh_tc, dh_ref)


if __name__ == "__main__":
    main()
