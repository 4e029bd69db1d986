"""Segmentation-head case: MIRAGELight (small encoder) + LinearSegAdapter / ConvNeXtAdapter,
shared by the CPU oracle test and the GPU parity test."""
from __future__ import annotations

import argparse

import torch

from helpers import load_synth


def seg_args():
    a = argparse.Namespace()
    a.in_domains = ["bscan"]
    a.patch_size = {"bscan": (32, 32)}
    a.input_size = {"bscan": (512, 512)}
    a.grid_sizes = {"bscan": [16, 16]}
    return a


def build_seg_model(kind: str, dim=128, depth=2, heads=2):
    from mirage_b200.input_adapters import PatchedInputAdapter
    from mirage_b200.model import MIRAGELight
    from mirage_b200.output_adapters import ConvNeXtAdapter, LinearSegAdapter
    if kind == "linear":
        head = LinearSegAdapter(num_classes=13, main_tasks=("bscan",), patch_size=[32, 32], task="bscan",
                                image_size=(512, 512))
    else:
        head = ConvNeXtAdapter(num_classes=13, embed_dim=2048, preds_per_patch=16, main_tasks=("bscan",),
                               patch_size=[32, 32], depth=2, task="bscan", image_size=(512, 512))
    ins = {"bscan": PatchedInputAdapter(num_channels=1, stride_level=1, patch_size_full=(32, 32), image_size=(512, 512))}
    m = MIRAGELight(seg_args(), input_adapters=ins, output_adapters={"bscan": head}, num_global_tokens=1,
                    dim_tokens=dim, depth=depth, num_heads=heads, drop_path_rate=0.0)
    return m


def oracle_seg(x, sd, kind, depth=2, heads=2):
    from oracle import mirage_oracle as O
    enc_sd = {k: v for k, v in sd.items() if not k.startswith("output_adapters.")}
    tok = O.light_forward(x, enc_sd, depth, heads)
    pre = "output_adapters.bscan."
    if kind == "linear":
        return O.linear_seg_adapter(tok, sd, pre, (16, 16), (512, 512))
    return O.convnext_adapter(tok, sd, pre, (16, 16), (512, 512), 16, 2)
