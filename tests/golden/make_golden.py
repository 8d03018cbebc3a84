"""Generate tests/golden/golden_outputs.npz from the literal float64 oracle (oracle/svgp_literal.py).

    python tests/golden/make_golden.py

The reference itself cannot be executed here (TensorFlow 1.15 / TFP 0.8 are not installable), so these
vectors pin the ORACLE, not the reference; they exist so that (i) the oracle cannot drift silently and (ii) the
GPU tests can compare against committed numbers as well as against a live oracle run.  Inputs are the
deterministic generators of svgp_vae_b200/configs.py (MNIST aux data from tests/golden/mnist_aux.npz).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refs  # noqa: E402
from oracle import svgp_literal as lit  # noqa: E402
from svgp_vae_b200 import configs  # noqa: E402


def cases():
    fx = os.path.join(HERE, "mnist_aux.npz")
    yield "mnist", "mnist", configs.mnist_inputs(fx, L=4), False
    yield "mnist_norm", "mnist", configs.mnist_inputs(fx, L=4, normalize=True), False
    yield "mnist_train_last", "mnist", configs.mnist_inputs(fx, L=2, b=210, rows="train", batch_index=15), False
    yield "sprites72", "sprites", configs.sprites_inputs(M=72, L=4), True
    yield "sprites72_raw", "sprites", configs.sprites_inputs(M=72, L=4, normalize=False), True
    yield "sweep_small", "sweep", configs.sweep_inputs(700, 40, 3), False


def main():
    out = {}
    for name, kind, cfg, clip in cases():
        o, _, op, _ = refs.make_pair(kind, cfg, "cpu")
        r, J, g = refs.oracle_objective(o, op, cfg["aux"], cfg["y"], cfg["noise"], clip_pv=clip)
        out[name + "/p_m"] = r["p_m"].detach().numpy()
        out[name + "/p_v"] = r["p_v"].detach().numpy()
        out[name + "/scalars"] = np.array([float(r[k]) for k in ("inside_elbo_recon", "inside_elbo_kl", "ce_term", "KL_term")] + [float(J)])
        out[name + "/grad_y"] = g[0].numpy()
        out[name + "/grad_noise"] = g[1].numpy()
        out[name + "/grad_Z"] = g[2].numpy()
    cfg = configs.ball_inputs()
    ox, oy = lit.BallSVGP(name="x", **cfg["ctor"]), lit.BallSVGP(name="y", **cfg["ctor"])
    r = lit.ball_glue(ox, oy, cfg["y"].double(), cfg["noise"].double())
    out["ball/p_m"], out["ball/p_v"] = r["p_m"].numpy(), r["p_v"].numpy()
    out["ball/KL_term"], out["ball/recon"], out["ball/kl"] = r["KL_term"].numpy(), r["inside_elbo_recon"].numpy(), r["inside_elbo_kl"].numpy()
    np.savez_compressed(os.path.join(HERE, "golden_outputs.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
