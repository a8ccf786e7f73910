"""Shared helpers for the parity tests."""
import torch

# north_star tolerance for fp16 storage / fp32 accumulate, per kernel, on identical fp16 inputs
RTOL, ATOL = 1e-3, 1e-4


def report(name, got, ref, rtol=RTOL, atol=ATOL):
    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    viol = (diff > bound).float().mean().item()
    worst = (diff / bound).max().item()
    print(f"[parity] {name}: max|d|={diff.max().item():.3e} max|ref|={ref.abs().max().item():.3e} "
          f"worst(d/bound)={worst:.2f} frac_viol={viol:.2e}")
    return worst, viol


def assert_close(name, got, ref, rtol=RTOL, atol=ATOL):
    assert torch.isfinite(got.float()).all(), f"{name}: non-finite output"
    worst, viol = report(name, got, ref, rtol, atol)
    assert worst <= 1.0, f"{name}: parity violated (worst diff/bound {worst:.2f}, {viol:.2e} of elements)"
