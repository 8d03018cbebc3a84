"""Actual gradient errors of the small reference configurations on the CUDA path against the golden outputs of the reference
source (the two places tests/test_gpu_e2e.py allows 3e-4).  Usage: python tests/probes/small_grad_errors.py -> JSON lines."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refs  # noqa: E402
from conftest import GOLDEN, MNIST_FIXTURE, rel_err  # noqa: E402
from svgp_vae_b200 import configs  # noqa: E402


def main():
    gold = np.load(os.path.join(GOLDEN, "reference_golden.npz"))
    T = lambda k: torch.from_numpy(gold[k])
    cases = [("mnist", lambda: configs.mnist_inputs(MNIST_FIXTURE, L=4), False),
             ("mnist_norm", lambda: configs.mnist_inputs(MNIST_FIXTURE, L=4, normalize=True), False),
             ("mnist_train_last", lambda: configs.mnist_inputs(MNIST_FIXTURE, L=2, b=210, rows="train", batch_index=15), False)]
    for name, maker, clip in cases:
        cfg = maker()
        _, s, _, sp = refs.make_pair("mnist", cfg, "cuda")
        r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda(), clip_pv=clip)
        out = {"case": name}
        for gname, g in zip(["y", "noise", "Z", "table", "amplitude", "length"], g1):
            ref_g = T(name + "/grad_" + gname)
            if ref_g.abs().max() > 0:
                out[gname] = float("%.3g" % rel_err(g, ref_g))
        print(json.dumps(out), flush=True)
    cfg = configs.mnist_inputs(MNIST_FIXTURE, L=2, b=64)
    cfg["ctor"]["titsias"] = True
    _, s, _, sp = refs.make_pair("mnist", cfg, "cuda")
    r1, J1, g1 = refs.product_objective(s, sp, cfg["aux"].cuda(), cfg["y"].cuda(), cfg["noise"].cuda())
    out = {"case": "mnist_titsias"}
    for g, n in zip(g1, ["y", "noise", "Z", "table"]):
        out[n] = float("%.3g" % rel_err(g, T("mnist_titsias/grad_" + n)))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
