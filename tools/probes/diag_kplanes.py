import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from svgp_vae_b200 import backend, configs
be = backend.get_backend()
N, M = 4096, 256
cfg = configs.sweep_inputs(N, M, 2, device="cuda")
Z = torch.from_numpy(cfg["ctor"]["initial_inducing_points"]).float().cuda()
kop = be.kernel_fwd((1, 4, 1, 4), cfg["aux"].float().contiguous(), Z.contiguous(), torch.ones(4, device="cuda"), tc=True, i8=True)
K = ((kop.Kh.double() + kop.Kl.double()) * kop.kscale[1].double())[:N, :M]
Kr, Kc = kop.value_i8("r"), kop.value_i8("c")
ur = (Kr - K).abs() / (K.abs().amax(1, keepdim=True) / 8323072.0)
uc = (Kc - K).abs() / (K.abs().amax(0, keepdim=True) / 8323072.0)
print("row units max", ur.max().item(), "col units max", uc.max().item())
print("rscale vs", (kop.rscale.double() / (K.abs().amax(1) / 8323072.0)).min().item(), (kop.rscale.double() / (K.abs().amax(1) / 8323072.0)).max().item())
print("cscale vs", (kop.cscale.double() / (K.abs().amax(0) / 8323072.0)).min().item(), (kop.cscale.double() / (K.abs().amax(0) / 8323072.0)).max().item())
i = ur.argmax() // M; j = ur.argmax() % M
print(i.item(), j.item(), K[i, j].item(), Kr[i, j].item(), K[i].abs().max().item(), kop.Kr[:, i, j])
