"""N-sharded SVGP step over 2 ranks (gloo, CPU, float64 stand-in backend) == the single-process step.

Covers the host logic of SURVEY 8(e): all-reduce of A_l / v_l / row sums in the forward, of their
adjoints in the backward, the global batch size in N_train / b, and the gradient convention
(per-rank loss = local terms + global scalars / world; replicated parameters get per-rank partial
gradients that sum to the single-process gradient)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import refs
from svgp_vae_b200 import configs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, cfg, clip, out, mm_chunk=None):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    from oracle_backend import OracleBackend
    from svgp_vae_b200 import backend
    backend.set_backend_for_tests(OracleBackend())
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, s, _, sp = refs.make_pair(kind, cfg, "cpu")
    n = cfg["aux"].shape[0] // world
    sl = slice(rank * n, (rank + 1) * n)
    y = cfg["y"][sl].clone().requires_grad_(True)
    nz = cfg["noise"][sl].clone().requires_grad_(True)
    res = s.elbo_step(cfg["aux"][sl], y, nz, clip_pv=clip, group=dist.group.WORLD, mm_chunk=mm_chunk)
    gm, gv = refs.upstream(tuple(cfg["y"].shape))
    J = (gm[sl].to(res["p_m"].dtype) * res["p_m"]).sum().double() + (gv[sl].to(res["p_v"].dtype) * res["p_v"]).sum().double() \
        + res["KL_term"] / world
    grads = torch.autograd.grad(J, [y, nz] + list(sp))
    pg = [g.double().clone() for g in grads[2:]]
    for g in pg:
        dist.all_reduce(g)
    Jt = J.detach().clone()
    dist.all_reduce(Jt)
    out[rank] = dict(p_m=res["p_m"].detach(), p_v=res["p_v"].detach(), KL_term=float(res["KL_term"]), J=float(Jt),
                     gy=grads[0], gn=grads[1], pg=pg, mu_hat=res["mu_hat"], A_hat=res["A_hat"])
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,clip,L,mm_chunk", [("mnist", False, 3, None), ("sprites", True, 3, None), ("mnist", False, 4, None),
                                                  ("sprites", True, 4, None), ("mnist", False, 4, 1)])
def test_sharded_step_matches_single_process(oracle_backend, kind, clip, L, mm_chunk):
    """L = 3: the float64 M x M stage stays replicated (3 % 2 != 0); L = 4: it is channel-sharded (two channels per rank,
    results all-gathered, dK_mm all-reduced); mm_chunk = 1: the channel-chunked stage under a group (replicated)."""
    from conftest import MNIST_FIXTURE, rel_err
    cfg = configs.mnist_inputs(MNIST_FIXTURE, L=L) if kind == "mnist" else configs.sprites_inputs(M=72, L=L, normalize=False)
    _, s, _, sp = refs.make_pair(kind, cfg, "cpu")
    r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"], cfg["y"], cfg["noise"], clip_pv=clip)
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, cfg, clip, out, mm_chunk), nprocs=world, join=True)
    n = cfg["aux"].shape[0] // world
    tol = 2e-6
    for r in range(world):
        sl = slice(r * n, (r + 1) * n)
        o = out[r]
        assert rel_err(o["p_m"], r1["p_m"][sl]) < tol and rel_err(o["p_v"], r1["p_v"][sl]) < tol
        assert abs(o["KL_term"] - float(r1["KL_term"])) < tol * abs(float(r1["KL_term"]))
        assert abs(o["J"] - float(J1)) < tol * abs(float(J1))
        assert rel_err(o["mu_hat"], r1["mu_hat"]) < tol and rel_err(o["A_hat"], r1["A_hat"]) < tol
        assert rel_err(o["gy"], g1[0][sl]) < 1e-5 and rel_err(o["gn"], g1[1][sl]) < 1e-5
        for a, b in zip(o["pg"], g1[2:]):
            if b.abs().max() > 0:
                assert rel_err(a, b) < 1e-5


# ---- forward_pass_SVGPVAE under a group: per-rank elbo shares sum to the single-process elbo, gradients likewise -----
class _ShardVAE:
    dtype = torch.float64

    def __init__(self, mu, var):
        self.mu, self.var = mu, var

    def encode(self, images):
        return self.mu, self.var

    def decode(self, z):
        base = torch.linspace(-1.0, 1.0, 28 * 28, dtype=z.dtype).reshape(1, 28, 28, 1)
        return torch.tanh(z.sum(1)).reshape(-1, 1, 1, 1) * base + 0.1 * z[:, :1].reshape(-1, 1, 1, 1)


def _shard_images(b):
    i = torch.arange(b, dtype=torch.float64).reshape(-1, 1, 1, 1)
    return torch.sin(0.1 * i + 3.0 * torch.linspace(0.0, 1.0, 28 * 28, dtype=torch.float64).reshape(1, 28, 28, 1))


def _glue_worker(rank, world, port, cfg, geco, eps, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    from oracle_backend import OracleBackend
    import svgp_vae_b200 as pkg
    from svgp_vae_b200 import backend
    backend.set_backend_for_tests(OracleBackend())
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, s, _, sp = refs.make_pair("mnist", cfg, "cpu")
    n = cfg["aux"].shape[0] // world
    sl = slice(rank * n, (rank + 1) * n)
    mu, var = cfg["y"][sl].clone().requires_grad_(True), cfg["noise"][sl].clone().requires_grad_(True)
    r = pkg.forward_pass_SVGPVAE((_shard_images(cfg["aux"].shape[0])[sl], cfg["aux"][sl]), beta=0.7, vae=_ShardVAE(mu, var), svgp=s,
                                 C_ma=0.3, lagrange_mult=1.5, alpha=0.99, kappa=0.02, clipping_qs=True, GECO=geco, epsilon=eps[sl],
                                 group=dist.group.WORLD)
    g = torch.autograd.grad(r[0], [mu, var] + list(sp))
    pg = [t.double().clone() for t in g[2:]]
    for t in pg:
        dist.all_reduce(t)
    e = r[0].detach().clone()
    dist.all_reduce(e)
    out[rank] = dict(elbo=float(e), C_ma=float(r[13]), lm=float(r[14]), gmu=g[0], gvar=g[1], pg=pg)
    dist.destroy_process_group()


@pytest.mark.parametrize("geco", [False, True])
def test_glue_sharded_matches_single_process(oracle_backend, geco):
    """the global KL_term must enter each rank's elbo share divided by the world size (GECO: the moving average
    and the multiplier from the all-reduced reconstruction loss)."""
    from conftest import MNIST_FIXTURE, rel_err
    import svgp_vae_b200 as pkg
    cfg = configs.mnist_inputs(MNIST_FIXTURE, L=4)
    b = cfg["aux"].shape[0]
    eps = torch.randn(b, 4, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    _, s, _, sp = refs.make_pair("mnist", cfg, "cpu")
    mu, var = cfg["y"].clone().requires_grad_(True), cfg["noise"].clone().requires_grad_(True)
    r1 = pkg.forward_pass_SVGPVAE((_shard_images(b), cfg["aux"]), beta=0.7, vae=_ShardVAE(mu, var), svgp=s, C_ma=0.3, lagrange_mult=1.5,
                                  alpha=0.99, kappa=0.02, clipping_qs=True, GECO=geco, epsilon=eps)
    g1 = torch.autograd.grad(r1[0], [mu, var] + list(sp))
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_glue_worker, args=(world, _free_port(), cfg, geco, eps, out), nprocs=world, join=True)
    n = b // world
    for r in range(world):
        o, sl = out[r], slice(r * n, (r + 1) * n)
        assert abs(o["elbo"] - float(r1[0])) < 1e-8 * abs(float(r1[0]))
        assert abs(o["C_ma"] - float(r1[13])) < 1e-10 * abs(float(r1[13])) and abs(o["lm"] - float(r1[14])) < 1e-10 * abs(float(r1[14]))
        assert rel_err(o["gmu"], g1[0][sl]) < 1e-6 and rel_err(o["gvar"], g1[1][sl]) < 1e-6
        for a, c in zip(o["pg"], g1[2:]):
            if c.abs().max() > 0:
                assert rel_err(a, c) < 1e-6
