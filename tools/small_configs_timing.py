"""Step latency of the reference's own configurations (BASELINE.json configs[0..2]) on the B200: fused elbo_step
(fwd + bwd of KL_term + <g, p_m> + <g, p_v>) and the reference's per-channel calling pattern, CUDA events, JSON lines."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import svgp_vae_b200 as pkg  # noqa: E402
from svgp_vae_b200 import configs  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def fused(svgp, cfg, clip, graphed=False):
    aux, y, nz = cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    gm, gv = torch.randn(y.shape, generator=g, device="cuda"), torch.randn(y.shape, generator=g, device="cuda")
    call = pkg.GraphedElboStep(svgp, aux, y, nz, clip_pv=clip) if graphed else (lambda a, b, c: svgp.elbo_step(a, b, c, clip_pv=clip))

    def step():
        yy, nn = y.clone().requires_grad_(True), nz.clone().requires_grad_(True)
        res = call(aux, yy, nn)
        J = res["KL_term"] + (gm.to(res["p_m"].dtype) * res["p_m"]).sum().double() + (gv.to(res["p_v"].dtype) * res["p_v"]).sum().double()
        J.backward()
    return step


def main():
    fx = os.path.join(ROOT, "tests", "golden", "mnist_aux.npz")
    cases = [("mnist b=256 m=32 L=16", "mnist", configs.mnist_inputs(fx, L=16), False),
             ("sprites b=500 M=72 L=64", "sprites", configs.sprites_inputs(M=72, L=64), True),
             ("sprites b=500 M=500 L=64", "sprites", configs.sprites_inputs(M=500, L=64), True)]
    for name, kind, cfg, clip in cases:
        cls = pkg.mnistSVGP if kind == "mnist" else pkg.spritesSVGP
        svgp = cls(name="t", **cfg["ctor"]).cuda()
        ms = timeit(fused(svgp, cfg, clip))
        b = cfg["aux"].shape[0]
        print(json.dumps({"config": name, "path": "elbo_step fwd+bwd", "ms": ms, "datapoints_per_s": b / ms * 1e3}), flush=True)
        try:
            ms = timeit(fused(svgp, cfg, clip, graphed=True), reps=100)
            print(json.dumps({"config": name, "path": "GraphedElboStep fwd+bwd (CUDA graphs)", "ms": ms, "datapoints_per_s": b / ms * 1e3}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"config": name, "path": "GraphedElboStep", "error": repr(e)[:300]}), flush=True)
    # moving ball: two SVGP objects, the reference's per-object calls (batch 35 x tmax 30, m = 15)
    cfg = configs.ball_inputs()
    sx, sy = pkg.SVGP(name="x", **cfg["ctor"]).cuda(), pkg.SVGP(name="y", **cfg["ctor"]).cuda()
    x, y, nz = cfg["x"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda()

    def ball():
        yy, nn = y.clone().requires_grad_(True), nz.clone().requires_grad_(True)
        tot = 0.0
        for ch, s in enumerate((sx, sy)):
            pm, B, mu, Ah = s.approximate_posterior_params(x, y=yy[:, :, ch], noise=nn[:, :, ch])
            a, b_ = s.variational_loss(x, yy[:, :, ch], nn[:, :, ch], mu_hat=mu, A_hat=Ah)
            tot = tot + (a - b_).sum() + pm.sum() + torch.diagonal(B, dim1=-2, dim2=-1).sum()
        tot.backward()
    ms = timeit(ball, reps=10)
    print(json.dumps({"config": "ball batch=35 tmax=30 m=15", "path": "per-object calls fwd+bwd", "ms": ms, "videos_per_s": 35 / ms * 1e3}), flush=True)
    # the same SVGP part through the product glue (ball_svgp_terms), eager and captured as CUDA graphs (GraphedBallStep)
    mu0, var0 = cfg["y"].cuda().float(), cfg["noise"].cuda().float()

    def ball_glue(fn):
        def run():
            m, v = mu0.clone().requires_grad_(True), var0.clone().requires_grad_(True)
            pm, pv, kl = fn(m, v)
            (kl.sum() + pm.sum() + pv.sum()).backward()
        return run

    def eager(m, v):
        t = pkg.ball_svgp_terms(sx, sy, m, v)
        return t["full_p_mu"], t["full_p_var"], t["KL_term"]
    try:
        ms = timeit(ball_glue(eager), reps=10)
        print(json.dumps({"config": "ball batch=35 tmax=30 m=15", "path": "ball_svgp_terms fwd+bwd (eager)", "ms": ms, "videos_per_s": 35 / ms * 1e3}), flush=True)
        step = pkg.GraphedBallStep(sx, sy, mu0.shape[0], mu0.shape[1])
        ms = timeit(ball_glue(step), reps=100)
        step.check()
        print(json.dumps({"config": "ball batch=35 tmax=30 m=15", "path": "GraphedBallStep fwd+bwd (CUDA graphs)", "ms": ms, "videos_per_s": 35 / ms * 1e3}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"config": "ball", "path": "GraphedBallStep", "error": repr(e)[:300]}), flush=True)


if __name__ == "__main__":
    main()
