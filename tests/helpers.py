"""Shared test utilities: deterministic synthetic weights / inputs and parity metrics."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def synth_state_dict(shapes: dict, seed: int = 0) -> dict:
    """Deterministic (CPU mt19937) weights for a name -> shape map, in sorted-key order.

    Weight matrices ~ N(0, 1/fan_in) (so activations stay O(1) through 24 layers), LayerNorm gains
    ~ 1 + 0.1 N(0,1), biases and tokens ~ 0.02-0.1 N(0,1).  ``pos_emb`` entries are NOT generated
    here (they are fixed sin-cos tables; callers keep the model's own).
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if name.endswith("pos_emb"):
            continue
        r = torch.randn(shp, generator=g)
        if "norm" in name and name.endswith("weight"):
            out[name] = 1.0 + 0.1 * r
        elif name.endswith("bias"):
            out[name] = 0.1 * r
        elif name.endswith("weight") and len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            out[name] = r * fan_in ** -0.5
        else:
            out[name] = 0.05 * r if "class_emb" not in name else 0.5 * r
    return out


def load_synth(model: torch.nn.Module, seed: int = 0) -> dict:
    """Fill ``model`` with synth_state_dict weights (pos_emb untouched); returns the full state_dict."""
    sd = model.state_dict()
    new = synth_state_dict({k: v.shape for k, v in sd.items()}, seed)
    for k, v in new.items():
        sd[k] = v.to(sd[k].dtype)
    model.load_state_dict(sd)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def synth_images(batch: int, modalities, seed: int = 1234, size: int = 512) -> dict:
    g = torch.Generator().manual_seed(seed)
    x = {}
    for m in modalities:
        if m == "bscanlayermap":
            x[m] = torch.randint(0, 13, (batch, size // 4, size // 4), generator=g)
        else:
            x[m] = torch.rand(batch, 1, size, size, generator=g)
    return x


def parity(got: torch.Tensor, ref: torch.Tensor) -> dict:
    """max|d|/max|ref|, relative Frobenius error and the worst per-row cosine."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    d = (got - ref)
    g2 = got.reshape(-1, got.shape[-1])
    r2 = ref.reshape(-1, ref.shape[-1])
    cos = torch.nn.functional.cosine_similarity(g2, r2, dim=-1)
    return {
        "max_rel": (d.abs().max() / (ref.abs().max() + 1e-12)).item(),
        "rel_fro": (d.norm() / (ref.norm() + 1e-12)).item(),
        "min_cos": cos.min().item(),
    }


# bf16 tolerance stated by BASELINE.json north_star: max-rel <= 2e-2, cosine >= 0.999 vs the fp32 reference
TOL_MAX_REL = 2e-2
TOL_COS = 0.999


def assert_parity(got, ref, what="", max_rel=TOL_MAX_REL, cos=TOL_COS):
    m = parity(got, ref)
    assert m["max_rel"] <= max_rel and m["min_cos"] >= cos, f"{what}: {m}"
    return m
