"""Deterministic inputs for the five configurations of BASELINE.json / SURVEY section 8(d).

The reference ships no data generator usable here (its loaders need TF and the missing
train_data3.p), so every config is reproduced shape-for-shape from the reference's own settings:

  ball     batch 35 x tmax 30, m = 15, f32, jitter 1e-9, l = 2  (BALL_experiment.py:38-44,96-102,326)
  mnist    b = 256 rows of the real auxiliary data, m = 32, L = 16, f64, jitter 1e-6, N_train 4050
           (MNIST_experiment.py:40-43,92-115,1128-1153); needs tests/golden/mnist_aux.npz
  sprites  b = 500, M in {72, 500}, L = 64, L_action 8, L_character 16, jitter 1e-2, N_train 50000
           (SPRITES_experiment.py:33-38,100-105,183,595-616)
  sweep    N rows, M inducing points, L channels, product-SE kernel d = 4 + 4

Everything is generated on the CPU with a seeded torch.Generator and moved to ``device``.
"""
import math

import numpy as np
import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def ball_inputs(device="cpu", seed=0, batch=35, tmax=30):
    g = _gen(seed)
    y = torch.randn(batch, tmax, 2, generator=g)
    noise = torch.exp(-3.0 + torch.randn(batch, tmax, 2, generator=g)).clamp(1e-6, 1e3)
    x = (torch.arange(tmax, dtype=torch.float32) + 1.0).repeat(batch, 1)          # SVGPVAE_model.py:663-664
    return dict(x=x.to(device), y=y.to(device), noise=noise.to(device),
                ctor=dict(titsias=False, num_inducing_points=15, fixed_inducing_points=True, tmin=1, tmax=tmax,
                          vidlt=2.0, fixed_gp_params=True, jitter=1e-9, ip_min=1, ip_max=tmax, GP_init=2.0))


def mnist_train_aux(fix):
    """Rebuild the (4050, 10) training auxiliary data from the shipped mask + PCA table."""
    pca, mask = fix["pca_ov_init"], fix["train_mask"]
    angles = np.linspace(0, 2 * np.pi, 17)[:-1]
    test_angle = 7
    train_angles = np.array([a for k, a in enumerate(angles) if k != test_angle])
    rows = []
    for i in range(360):
        for k in range(15):
            if mask[i * 15 + k]:
                rows.append(np.concatenate([[float(i), train_angles[k]], pca[i]]))
    return np.asarray(rows)


def mnist_inputs(fixture_path, device="cpu", L=16, b=256, rows="eval", normalize=False, batch_index=0):
    fix = np.load(fixture_path)
    pca = fix["pca_ov_init"]
    if rows == "eval":
        aux_all = fix["eval_aux"]
    else:
        aux_all = mnist_train_aux(fix)
    aux = aux_all[batch_index * b: batch_index * b + b]
    r = np.arange(32)
    Z = np.concatenate([r[:, None].astype(np.float64), (2 * np.pi * (r // 2) / 16.0)[:, None],
                        pca[(7 * r) % 400] * (1.0 + 0.05 * np.sin(r))[:, None]], axis=1)
    i = torch.arange(aux.shape[0], dtype=torch.float64)[:, None]
    l = torch.arange(L, dtype=torch.float64)[None, :]
    y = torch.sin(0.37 * i + l)
    noise = 0.05 + 0.25 * (1.0 + torch.cos(0.11 * i + 2.0 * l))
    return dict(aux=torch.from_numpy(aux).to(device), y=y.to(device), noise=noise.to(device),
                ctor=dict(titsias=False, fixed_inducing_points=False, initial_inducing_points=Z, fixed_gp_params=False,
                          object_vectors_init=pca, jitter=1e-6, N_train=4050, L=L, K_obj_normalize=normalize))


def sprites_inputs(device="cpu", seed=0, M=72, L=64, b=500, K_SE=False, normalize=True):
    g = _gen(seed)
    La, Lc = 8, 16
    action = 1.5 * torch.randn(72, La, generator=g)                              # SPRITES_experiment.py:101
    Z = 1.5 * torch.randn(M, La + Lc, generator=g)                               # :102-105
    ids = torch.randint(0, 72, (b,), generator=g)
    chars = torch.randn(b // 50, Lc, generator=g).repeat_interleave(50, dim=0)   # SVGPVAE_model.py:1107-1110
    aux = torch.cat([ids[:, None].float(), chars], dim=1)
    y = torch.randn(b, L, generator=g)
    noise = torch.exp(-2.0 + 0.5 * torch.randn(b, L, generator=g)).clamp(1e-3, 10.0)
    return dict(aux=aux.to(device), y=y.to(device), noise=noise.to(device),
                ctor=dict(titsias=False, fixed_inducing_points=False, initial_inducing_points=Z.numpy(), jitter=1e-2,
                          N_train=50000, L_action=La, initial_GPLVM_action=action.numpy(), L_character=Lc, L=L,
                          K_obj_normalize=normalize, K_SE=K_SE))


def sweep_inputs(N, M, L, device="cpu", seed=1234, rank=0, N_train=None):
    """Product-SE kernel d = 4 + 4, sigma = l = 1, jitter 1e-2; X ~ N(0,1) per shard, Z ~ N(0,1) seed 7."""
    gz = _gen(7)
    Z = torch.randn(M, 8, generator=gz)
    if str(device).startswith("cuda"):
        g = torch.Generator(device=device)
        g.manual_seed(seed + rank)
        X = torch.randn(N, 8, generator=g, device=device)
        y = torch.randn(N, L, generator=g, device=device)
        noise = torch.exp(-2.0 + 0.5 * torch.randn(N, L, generator=g, device=device)).clamp_(1e-3, 10.0)
    else:
        g = _gen(seed + rank)
        X = torch.randn(N, 8, generator=g)
        y = torch.randn(N, L, generator=g)
        noise = torch.exp(-2.0 + 0.5 * torch.randn(N, L, generator=g)).clamp_(1e-3, 10.0)
    return dict(aux=X, y=y, noise=noise,
                ctor=dict(initial_inducing_points=Z.numpy(), dim_a=4, dim_b=4, jitter=1e-2,
                          N_train=N if N_train is None else N_train, L=L))
