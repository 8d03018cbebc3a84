"""Time the three tcgen05 kernels and K1 in isolation (CUDA events) and report algorithmic TFLOP/s / GB/s.
Usage: python tools/tc_probe.py [N M L]   -> JSON lines on stdout."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svgp_vae_b200 import backend  # noqa: E402


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    N, M, L = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (131072, 1024, 8)
    only = sys.argv[4] if len(sys.argv) >= 5 else None          # "syrk": time the SYRK only
    be = backend.get_backend()
    g = torch.Generator(device="cuda").manual_seed(0)
    Fx = torch.randn(N, 8, generator=g, device="cuda"); Fz = torch.randn(M, 8, generator=g, device="cuda")
    hyp = torch.ones(4, device="cuda")
    spec = (1, 4, 1, 4)
    if not only:
        t = timeit(lambda: be.kernel_fwd(spec, Fx, Fz, hyp, tc=True, i8=False))
        print(json.dumps(dict(op="kernel_fwd_planes", N=N, M=M, ms=t, GBs=4 * N * M * 2 / t / 1e6)), flush=True)
        t = timeit(lambda: be.kernel_fwd(spec, Fx, Fz, hyp, tc=False))
        print(json.dumps(dict(op="kernel_fwd_f32", N=N, M=M, ms=t, GBs=N * M * 4 / t / 1e6)), flush=True)
    kop = be.kernel_fwd(spec, Fx, Fz, hyp, tc=True, i8=False)
    W = torch.randn(N, L, generator=g, device="cuda")
    S = torch.randn(L, M, M, generator=g, device="cuda", dtype=torch.float64); S = (S + S.transpose(1, 2)).contiguous()
    Lt = torch.tril(S).contiguous()
    tag = os.environ.get("SVGP_TC_BK", "default")
    flush = os.environ.get("SVGP_SYRK_FLUSH", "default")
    for chunk in ((128, 256, 512, 1024, 2048) if only != "quad" else ()):
        t = timeit(lambda: be.syrk(kop, W, chunk_rows=chunk))
        print(json.dumps(dict(op="syrk_tc", bk=tag, flush=flush, chunk=chunk, N=N, M=M, L=L, ms=t,
                              alg_TFLOPs=N * M * M * L / t / 1e9)), flush=True)
    if only == "syrk":
        return
    if only == "quad":
        Ltpl = be.planes(Lt)
        t = timeit(lambda: be.rowquad(kop, Ltpl, tri=True))
        print(json.dumps(dict(op="rowquad_tc_tri", N=N, M=M, L=L, ms=t, alg_TFLOPs=N * M * M * L / t / 1e9)), flush=True)
        return
    Spl, Ltpl = be.planes(S), be.planes(Lt)
    t = timeit(lambda: be.planes(S))
    print(json.dumps(dict(op="split_f16", M=M, L=L, ms=t)), flush=True)
    t = timeit(lambda: be.rowquad(kop, Spl))
    print(json.dumps(dict(op="rowquad_tc_full", bk=tag, N=N, M=M, L=L, ms=t, alg_TFLOPs=2 * N * M * M * L / t / 1e9)), flush=True)
    t = timeit(lambda: be.rowquad(kop, Ltpl, tri=True))
    print(json.dumps(dict(op="rowquad_tc_tri", bk=tag, N=N, M=M, L=L, ms=t, alg_TFLOPs=N * M * M * L / t / 1e9)), flush=True)
    t = timeit(lambda: be.scaled_gemm(kop, W, Spl))
    print(json.dumps(dict(op="scaled_gemm_tc", bk=tag, N=N, M=M, L=L, ms=t, alg_TFLOPs=2 * N * M * M * L / t / 1e9)), flush=True)
    t = timeit(lambda: be.scaled_gemm(kop, W, Spl, ndot=L))
    print(json.dumps(dict(op="scaled_gemm_tc_dots", bk=tag, N=N, M=M, L=L, ms=t, alg_TFLOPs=2 * N * M * M * L / t / 1e9)), flush=True)
    G = torch.randn(N, M, generator=g, device="cuda")
    t = timeit(lambda: be.kernel_bwd(spec, Fx, Fz, hyp, G))
    print(json.dumps(dict(op="kernel_bwd", N=N, M=M, ms=t)), flush=True)
    X = torch.randn(64, M, M, generator=g, device="cuda", dtype=torch.float64)
    X = X @ X.transpose(1, 2) / M + torch.eye(M, device="cuda", dtype=torch.float64)
    t = timeit(lambda: be.chol(X)); print(json.dumps(dict(op="chol64", M=M, batch=64, ms=t)), flush=True)
    Lf, _ = be.chol(X)
    t = timeit(lambda: be.trinv(Lf)); print(json.dumps(dict(op="trinv64", M=M, batch=64, ms=t)), flush=True)
    t = timeit(lambda: be.bmm64(X, X)); print(json.dumps(dict(op="bmm64", M=M, batch=64, ms=t, TFLOPs=2 * 64 * M ** 3 / t / 1e9)), flush=True)
    t = timeit(lambda: be.gemm_tn(kop, W)); print(json.dumps(dict(op="gemm_tn", ms=t)), flush=True)
    t = timeit(lambda: be.gemm_nn(kop, torch.randn(L, M, device="cuda"))); print(json.dumps(dict(op="gemm_nn", ms=t)), flush=True)


if __name__ == "__main__":
    main()
