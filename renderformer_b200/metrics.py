"""Parity metrics of BASELINE.json's north_star: max relative error on pre-tonemap
HDR pixels and PSNR on log-HDR, both against the fp32 oracle."""
from __future__ import annotations

import math

import torch

REL_TOL = 2e-2   # max|a-b| / max|b| on HDR pixels
PSNR_MIN = 45.0  # dB on log10(x+1), peak = data range of the reference log image


def hdr_rel_err(test: torch.Tensor, ref: torch.Tensor) -> float:
    ref = ref.double().cpu()
    return ((test.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def log_psnr(test: torch.Tensor, ref: torch.Tensor) -> float:
    lt = torch.log10(test.double().cpu().clamp_min(-0.999) + 1.0)
    lr = torch.log10(ref.double().cpu().clamp_min(-0.999) + 1.0)
    mse = ((lt - lr) ** 2).mean().item()
    peak = (lr.max() - lr.min()).item()
    if mse == 0:
        return float("inf")
    return 10.0 * math.log10(max(peak, 1e-12) ** 2 / mse)


def rel_l2(test: torch.Tensor, ref: torch.Tensor) -> float:
    ref = ref.double().cpu()
    return ((test.double().cpu() - ref).norm() / ref.norm().clamp_min(1e-12)).item()
