"""Manual diagnostic (not collected by pytest): localise tcgen05 GEMM layout bugs on a real B200.

python tests/diag_gemm.py  -> prints error maps for a single 128xBN tile with structured inputs.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videomv_b200 import ops  # noqa: E402


def errmap(out, ref, rb=16, cb=16):
    d = (out.float() - ref).abs()
    M, N = d.shape
    m = d[: M // rb * rb, : N // cb * cb].reshape(M // rb, rb, N // cb, cb).amax(dim=(1, 3))
    return m


def run(M, N, K, bn, stages, kind):
    g = torch.Generator(device="cuda").manual_seed(0)
    if kind == "rand":
        a = torch.randn(M, K, generator=g, device="cuda").half()
        w = torch.randn(N, K, generator=g, device="cuda").half() * K ** -0.5
    elif kind == "eye":       # out[m, n] = a[m, n] for n < K
        a = torch.randn(M, K, generator=g, device="cuda").half()
        w = torch.zeros(N, K, device="cuda").half()
        for i in range(min(N, K)):
            w[i, i] = 1
    elif kind.startswith("kblk"):   # only one 16-wide K slice non-zero
        j = int(kind[4:])
        a = torch.zeros(M, K, device="cuda").half()
        a[:, 16 * j:16 * j + 16] = torch.randn(M, 16, generator=g, device="cuda").half()
        w = torch.randn(N, K, generator=g, device="cuda").half() * 0.25
    try:
        out = ops.gemm(a, w, block_n=bn, stages=stages)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"  {kind}: EXCEPTION {e}")
        return False
    ref = a.float() @ w.float().t()
    err = (out.float() - ref).abs().max().item()
    print(f"  {kind:8s} M{M} N{N} K{K} bn{bn} st{stages}: max err {err:.4e}  (max ref {ref.abs().max().item():.3f})")
    if err > 1e-2:
        em = errmap(out, ref)
        print("   error map (rows of 16 x cols of 16), >1e-2 marked X:")
        for r in range(em.shape[0]):
            print("   ", "".join("X" if v > 1e-2 else "." for v in em[r].tolist()))
        return False
    return True


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    ok = True
    for bn in (128, 160, 64, 256):
        print(f"block_n {bn}")
        for kind in ("eye", "rand", "kblk0", "kblk1", "kblk3"):
            ok &= run(128, bn, 64, bn, 3, kind)
        ok &= run(256, 2 * bn, 256, bn, 3, "rand")
        ok &= run(256, 2 * bn, 1024, bn, 0, "rand")
    print("DIAG", "PASS" if ok else "FAIL")
