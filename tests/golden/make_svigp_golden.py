"""Golden vectors for SVIGP_Hensman from the REFERENCE SOURCE (SVIGP_Hensman_model.py, unmodified) executed under
tests/golden/tf_shim.py -- same method as make_reference_golden.py.

    python tests/golden/make_svigp_golden.py          # needs /root/reference (this container only)

Inputs: the rotated-MNIST configuration of svgp_vae_b200/configs.py (eval aux rows, 32 inducing points, PCA table),
L = 3 channels, variational parameters seeded below (the reference initialises them to zeros / identity, which would
leave most terms trivial).  Output: tests/golden/svigp_golden.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REFERENCE = os.environ.get("SVGP_REFERENCE_DIR", "/root/reference")

import tf_shim  # noqa: E402

tf = tf_shim.install()
sys.path.insert(0, REFERENCE)
import SVIGP_Hensman_model as ref  # noqa: E402  (the reference, unmodified)

from svgp_vae_b200 import configs  # noqa: E402

F64 = torch.float64
MNIST_FIXTURE = os.path.join(HERE, "mnist_aux.npz")


def variational_init(L, m, seed=3):
    g = torch.Generator().manual_seed(seed)
    mu = 0.5 * torch.randn(L, m, generator=g, dtype=F64)
    A = torch.eye(m, dtype=F64).repeat(L, 1, 1) * 0.7 + 0.05 * torch.tril(torch.randn(L, m, m, generator=g, dtype=F64))
    return mu, A


def upstream(b, L, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, L, generator=g, dtype=F64)


def case(out, name, normalize):
    L = 3
    cfg = configs.mnist_inputs(MNIST_FIXTURE, L=L, b=192, normalize=normalize)
    c = cfg["ctor"]
    svgp = ref.SVIGP_Hensman(fixed_inducing_points=False, initial_inducing_points=c["initial_inducing_points"], name="ref",
                             jitter=c["jitter"], N_train=c["N_train"], dtype=np.float64, L=L, fixed_gp_params=False,
                             object_vectors_init=c["object_vectors_init"], K_obj_normalize=normalize)
    m = svgp.nr_inducing
    mu0, A0 = variational_init(L, m)
    mus = [mu0[l].clone().requires_grad_(True) for l in range(L)]
    As = [A0[l].clone().requires_grad_(True) for l in range(L)]
    svgp.variational_inducing_observations_loc = mus
    svgp.variational_inducing_observations_scale = As
    svgp.variational_inducing_observations_cov_mat = [tf.matmul(x, tf.transpose(x)) for x in As]        # :72-73
    svgp.noise = torch.tensor(0.3, dtype=F64, requires_grad=True)
    aux = cfg["aux"].to(F64)
    rec, kl, means = [], [], []
    for l in range(L):                                                                                  # :246-252
        r_l, k_l, m_l = svgp.variational_loss(x=aux, z=None, lat_channel=l)
        rec.append(r_l); kl.append(k_l); means.append(m_l)
    rec, kl, means = torch.stack(rec), torch.stack(kl), torch.stack(means, 1)
    gm = upstream(aux.shape[0], L)
    b = float(aux.shape[0])
    J = rec.sum() - (b / svgp.N_train) * kl.sum() + (gm * means).sum()                                  # :254-257 + decoder stand-in
    leaves = [svgp.inducing_index_points, svgp.object_vectors, svgp.amplitude, svgp.l_GP, svgp.noise] + mus + As
    grads = torch.autograd.grad(J, leaves, allow_unused=True)
    out[name + "/L3"], out[name + "/KL"], out[name + "/mean"] = rec.detach().numpy(), kl.detach().numpy(), means.detach().numpy()
    out[name + "/J"] = np.array(float(J))
    for n, g in zip(["Z", "table", "amplitude", "length", "noise"], grads[:5]):
        out[name + "/grad_" + n] = g.detach().numpy()
    out[name + "/grad_mu"] = torch.stack(grads[5:5 + L]).detach().numpy()
    out[name + "/grad_A"] = torch.stack(grads[5 + L:]).detach().numpy()
    with torch.no_grad():
        test_aux = aux[:40].clone()
        test_aux[:, 1] = test_aux[:, 1] + 0.3
        mv, B = svgp.approximate_posterior_params(test_aux, 1)                                           # :202-227
        out[name + "/post_mean"], out[name + "/post_B"] = mv.numpy(), B.numpy()


def main():
    out = {}
    case(out, "svigp", False)
    case(out, "svigp_norm", True)
    path = os.path.join(HERE, "svigp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("svigp/")})


if __name__ == "__main__":
    main()
