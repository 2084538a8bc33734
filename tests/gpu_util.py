"""Shared helpers for the -m gpu tests: oracle weights on the device, tolerance gates."""
import torch

from oracle import reference_path as R

COS_MIN = 0.999        # BASELINE.json north_star tolerance
REL_MAX = 2e-2         # max|x-y| / max|y|


def gate(got: torch.Tensor, want: torch.Tensor, what: str = "", cos_min: float = COS_MIN, rel_max: float = REL_MAX):
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    cos = float(R.cosine_rows(got, want).min())
    rel = R.max_rel_err(got, want)
    print(f"[gate] {what}: cos_min={cos:.6f} rel_max={rel:.5f}")
    assert cos >= cos_min, f"{what}: cosine {cos:.6f} < {cos_min}"
    assert rel <= rel_max, f"{what}: max relative error {rel:.5f} > {rel_max}"
    return cos, rel


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
    return R.max_rel_err(got.detach().float().cpu(), want.detach().float().cpu())
