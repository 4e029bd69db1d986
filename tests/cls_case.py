"""Classification fine-tune case (BASELINE.json configs[4]): builders shared by the CPU
oracle test and the GPU parity test.  The wrapper classes need a checkpoint file like the reference's
(mirage_wrapper.py:48-89), so one is synthesised in a temp dir with the recipe of SURVEY.md 8(c)."""
from __future__ import annotations

import argparse
import contextlib
import io
import tempfile

import torch

from helpers import load_synth


def ckpt_args(size="base"):
    return argparse.Namespace(model=f"miragepre_{size}", out_domains=[], decoder_dim=256, decoder_depth=2,
                              decoder_num_heads=8, decoder_use_task_queries=True, decoder_use_xattn=True,
                              num_global_tokens=1, drop_path=0.0, grid_sizes={"bscan": [16, 16]})


def build_cls_model(pool: str, weights_seed: int, device="cpu", num_classes=5, size="base"):
    """mirage_b200.mirage_wrapper.miragecls_factory[pool] with synthetic weights (ViT-B/L encoder)."""
    from mirage_b200.mirage_wrapper import miragecls_factory
    cls_t = miragecls_factory[pool]
    with contextlib.redirect_stdout(io.StringIO()):
        helper = cls_t.__new__(cls_t)
        torch.nn.Module.__init__(helper)
        a2 = argparse.Namespace(**vars(ckpt_args(size)))
        a2.in_domains = ["bscan"]
        a2.patch_size = {"bscan": (32, 32)}
        a2.input_size = {"bscan": (512, 512)}
        helper.args = a2
        enc = helper.get_model()
        with tempfile.NamedTemporaryFile(suffix=".pth") as f:
            torch.save({"model": enc.state_dict(), "args": ckpt_args(size)}, f.name)
            m = cls_t(num_classes=num_classes, input_size=512, patch_size=32, modalities="bscan", weights=f.name,
                      device="cpu")
    sd = load_synth(m, weights_seed)
    return m.to(device), sd


def oracle_cls_logits(x: torch.Tensor, sd: dict, pool: str, size: str = "base"):
    """fp32 CPU oracle: patch embed -> (+ global token last) -> 12 blocks -> LayerNorm -> pool -> head.
    The reference feeds the encoder a random permutation of the 256 patch tokens; every
    pooling variant is invariant to it, so the oracle keeps the natural order."""
    from oracle import mirage_oracle as O
    msd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    tok = O.light_forward({"bscan": x}, msd, *((12, 12) if size == "base" else (24, 16)))
    return O.cls_head(tok, sd, pool)
