"""Print per-tensor relative errors (product on cuda vs float64 literal oracle) for every config, no asserts."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refs  # noqa: E402
from conftest import MNIST_FIXTURE, rel_err  # noqa: E402
from svgp_vae_b200 import backend, configs  # noqa: E402


def report(name, kind, cfg, clip=False, **kw):
    o, s, op, sp = refs.make_pair(kind, cfg, "cuda")
    r0, J0, g0 = refs.oracle_objective(o, op, cfg["aux"], cfg["y"], cfg["noise"], clip_pv=clip)
    try:
        r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), clip_pv=clip, **kw)
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(cfg=name, error=str(e)[:200])), flush=True)
        return
    out = dict(cfg=name, p_m=rel_err(r1["p_m"], r0["p_m"]), p_v=rel_err(r1["p_v"], r0["p_v"]))
    for k in ("inside_elbo_recon", "inside_elbo_kl", "ce_term", "KL_term"):
        out[k] = abs(float(r1[k]) - float(r0[k])) / abs(float(r0[k]))
    out["J"] = abs(float(J1) - float(J0)) / abs(float(J0))
    out["grads"] = [None if a is None else float("%.3g" % rel_err(b, a)) for a, b in zip(g0, g1)]
    print(json.dumps(out), flush=True)


def syrk_accuracy():
    be = backend.get_backend()
    g = torch.Generator(device="cuda").manual_seed(0)
    N, M, L = 16384, 256, 2
    Fx = torch.randn(N, 8, generator=g, device="cuda"); Fz = torch.randn(M, 8, generator=g, device="cuda")
    kop = be.kernel_fwd((1, 4, 1, 4), Fx, Fz, torch.ones(4, device="cuda"), tc=True)
    K64 = kop.value().double()
    for sign in ("pos", "mixed"):
        W = torch.rand(N, L, generator=g, device="cuda") + 0.1 if sign == "pos" else torch.randn(N, L, generator=g, device="cuda")
        ref = torch.einsum('il,ia,ib->lab', W.double(), K64, K64)
        for chunk in (128, 512, 1024, 4096, 16384):
            A = be.syrk(kop, W, chunk_rows=chunk)
            d = (A - ref)
            print(json.dumps(dict(op="syrk_acc", sign=sign, chunk=chunk, maxrel=float(d.abs().max() / ref.abs().max()),
                                  mean_signed_rel=float((d / ref).mean()))), flush=True)


def rowquad_accuracy():
    be = backend.get_backend()
    g = torch.Generator(device="cuda").manual_seed(0)
    for M in (128, 256, 1024):
        N, L = 4096, 2
        Fx = torch.randn(N, 8, generator=g, device="cuda"); Fz = torch.randn(M, 8, generator=g, device="cuda")
        kop = be.kernel_fwd((1, 4, 1, 4), Fx, Fz, torch.ones(4, device="cuda"), tc=True)
        K64 = kop.value().double()
        Lt = torch.tril(torch.rand(L, M, M, generator=g, device="cuda", dtype=torch.float64)).contiguous()   # all positive
        T = torch.einsum('ia,lca->ilc', K64, Lt)
        ref = (T * T).sum(-1)
        q = be.rowquad(kop, Lt, tri=True)
        d = q.double() - ref
        print(json.dumps(dict(op="rowquad_tri_acc_pos", M=M, maxrel=float(d.abs().max() / ref.abs().max()),
                              mean_signed_rel=float((d / ref).mean()))), flush=True)


if __name__ == "__main__":
    syrk_accuracy()
    rowquad_accuracy()
    report("mnist", "mnist", configs.mnist_inputs(MNIST_FIXTURE, L=16))
    report("mnist_norm", "mnist", configs.mnist_inputs(MNIST_FIXTURE, L=16, normalize=True))
    report("sprites72", "sprites", configs.sprites_inputs(M=72, L=64), clip=True)
    report("sprites72_raw", "sprites", configs.sprites_inputs(M=72, L=8, normalize=False), clip=True)
    report("sprites500", "sprites", configs.sprites_inputs(M=500, L=8), clip=True)
    report("sweep_simt", "sweep", configs.sweep_inputs(2304, 200, 3), tc=False)
    report("sweep_tc", "sweep", configs.sweep_inputs(2304, 200, 3), tc=True)
    report("sweep_tc_full", "sweep", configs.sweep_inputs(2304, 200, 3), tc=True, tri=False)
    pass
