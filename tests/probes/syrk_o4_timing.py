"""Cost of the thirteen-pair forward SYRK (SVGP_IMPL_TC_I8_O4) against the ten-pair one at configs[4]'s shapes on one GPU.
Usage: python tests/probes/syrk_o4_timing.py [N M L]   -> one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from svgp_vae_b200 import backend, configs  # noqa: E402
from svgp_vae_b200._lib import IMPL_TC_I8, IMPL_TC_I8_O4  # noqa: E402


def main():
    N, M, L = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (250000, 4096, 128)
    be = backend.get_backend()
    cfg = configs.sweep_inputs(N, M, 2, device="cuda")
    Z = torch.from_numpy(cfg["ctor"]["initial_inducing_points"]).float().cuda()
    kop = be.kernel_fwd((1, 4, 1, 4), cfg["aux"].float().contiguous(), Z.contiguous(), torch.ones(4, device="cuda"), tc=True, i8=True)
    W = torch.rand(N, L, device="cuda") + 0.5
    out = dict(N=N, M=M, L=L)
    for name, impl in (("ten_pairs_ms", IMPL_TC_I8), ("thirteen_pairs_ms", IMPL_TC_I8_O4)):
        A = be.syrk(kop, W, impl=impl)            # warm-up
        del A
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A = be.syrk(kop, W, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1), 1)
        del A
    print(json.dumps(out))


if __name__ == "__main__":
    main()
