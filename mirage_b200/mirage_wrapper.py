"""Checkpoint-driven wrappers with the API of the reference's ``mirage_wrapper.py``.

``MIRAGEWrapper(input_size, patch_size, modalities, weights, device)`` rebuilds the pretraining
model (encoder + SpatialOutputAdapters) from a checkpoint ``{'model': state_dict, 'args': Namespace}``
and reconstructs all modalities from whichever are given; ``miragecls_factory['global'|'cls'|
'token_mix']`` adds LayerNorm + pooling + Linear classification heads (mirage_wrapper.py:187-244).
"""
from __future__ import annotations

import copy
from functools import partial
from typing import Union

import numpy as np
import torch
from torch import nn

from . import functional as Fn
from .factory import get_factory_adder
from .input_adapters import PatchedInputAdapter, SemSegInputAdapter
from .model import MIRAGEModel
from .output_adapters import SpatialOutputAdapter
from .utils import pair

DEFAULT_CONF = {
    'channels': 1,
    'stride_level': 1,
    'input_adapter': partial(PatchedInputAdapter, num_channels=1),
    'output_adapter': partial(SpatialOutputAdapter, num_channels=1),
}

DOMAIN_CONF = {
    'bscan': copy.deepcopy(DEFAULT_CONF),
    'slo': copy.deepcopy(DEFAULT_CONF),
    'bscanlayermap': {
        'num_classes': 13,
        'stride_level': 1,
        'input_adapter': partial(SemSegInputAdapter, num_classes=13, dim_class_emb=64,
                                 interpolate_class_emb=False),
        'output_adapter': partial(SpatialOutputAdapter, num_channels=13),
    },
}

_ENCODER = {'large': dict(dim_tokens=1024, depth=24, num_heads=16), 'base': {}}


class MIRAGEWrapper(nn.Module):
    def __init__(self, input_size=512, patch_size=32, modalities='bscan-slo-bscanlayermap', weights=None,
                 device='cuda'):
        super().__init__()
        assert weights is not None
        ckpt = torch.load(weights, map_location=device, weights_only=False)
        args = ckpt['args']
        args.in_domains = modalities.split('-')
        input_size, patch_size = pair(input_size), pair(patch_size)
        assert input_size is not None and patch_size is not None
        args.patch_size, args.input_size, args.grid_size = {}, {}, {}
        for domain in args.in_domains:
            if domain != 'bscanlayermap':
                args.patch_size[domain] = patch_size
                args.input_size[domain] = input_size
            else:
                args.patch_size[domain] = (8, 8)
                args.input_size[domain] = (128, 128)
            args.grid_size[domain] = [input_size[i] // patch_size[i] for i in range(len(input_size))]
        self.args = args
        self.model = self.get_model()
        msg = self.model.load_state_dict(ckpt['model'], strict=False)
        assert len(msg.missing_keys) == 0, f'missing keys: {msg.missing_keys[:8]}'
        self.load_report = msg

    def get_output_adapters(self) -> Union[None, dict]:
        return {
            domain: DOMAIN_CONF[domain]['output_adapter'](
                stride_level=DOMAIN_CONF[domain]['stride_level'],
                patch_size_full=tuple(self.args.patch_size[domain]),
                dim_tokens=self.args.decoder_dim, depth=self.args.decoder_depth,
                num_heads=self.args.decoder_num_heads,
                use_task_queries=self.args.decoder_use_task_queries, task=domain,
                context_tasks=list(self.args.in_domains), use_xattn=self.args.decoder_use_xattn,
                image_size=self.args.input_size[domain])
            for domain in self.args.out_domains
        }

    def get_model(self):
        input_adapters = {
            domain: DOMAIN_CONF[domain]['input_adapter'](
                stride_level=DOMAIN_CONF[domain]['stride_level'],
                patch_size_full=tuple(self.args.patch_size[domain]),
                image_size=self.args.input_size[domain])
            for domain in self.args.in_domains
        }
        for size in ('large', 'base'):
            if size in self.args.model:
                return MIRAGEModel(args=self.args, input_adapters=input_adapters,
                                   output_adapters=self.get_output_adapters(),
                                   num_global_tokens=self.args.num_global_tokens,
                                   drop_path_rate=self.args.drop_path, **_ENCODER[size])
        raise ValueError('Unknown model size:', self.args.model)

    def forward(self, x: dict):
        """x: {modality: tensor in [0, 1]} for any subset of the model's modalities (batch 1, like the
        reference: missing modalities are zero-filled and fully masked).  Returns the predictions
        dict, or the encoder tokens when ``self.model.output_adapters`` is None.  Mutates ``x``."""
        masks = {}
        for k in self.args.in_domains:
            if k not in x:
                if k == 'bscanlayermap':
                    x[k] = torch.zeros((1, *self.args.input_size[k])).long()
                else:
                    x[k] = torch.zeros((1, 1, *self.args.input_size[k]))
                fill = 1
            else:
                fill = 0
            masks[k] = torch.LongTensor(np.full(self.args.grid_size[k], fill)).flatten()[None].to(self.device)
            x[k] = x[k].to(self.device)
        preds, _masks = self.model(x, mask_inputs=False, task_masks=masks)
        return preds

    @property
    def device(self):
        return next(self.parameters()).device


add_miragecls, miragecls_factory = get_factory_adder()


@add_miragecls('global')
class MIRAGEClsGlobal(MIRAGEWrapper):
    """Encoder -> LayerNorm -> mean over patch tokens -> Linear (mirage_wrapper.py:190-227)."""

    def __init__(self, num_classes=0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert num_classes > 0
        assert len(self.args.in_domains) == 1
        self.num_classes = num_classes
        self.model.output_adapters = None
        self.embed_dim = self.model.encoder[0].norm1.normalized_shape[0]
        self.norm = nn.LayerNorm(self.embed_dim, eps=1e-06, elementwise_affine=True)
        self.build_head()

    def build_head(self, factor=1):
        self.head = nn.Linear(self.embed_dim * factor, self.num_classes)

    def forward(self, x):
        x_d = {self.args.in_domains[0]: x}
        out, _masks = self.model(x_d, mask_inputs=False)
        B, N, D = out.shape
        ranges = self._pool_ranges(N)
        if ranges is not None and out.is_cuda and D % 128 == 0 and D <= 1024:
            # fused LayerNorm + token mean (csrc/pool.cu): one pass over the encoder output
            out = Fn.ln_meanpool(out, self.norm.weight, self.norm.bias, self.norm.eps, ranges)
        else:
            # a user-overridden pool(): normalise, then call it (mirage_wrapper.py:219-221)
            out = Fn.layer_norm(out.reshape(B * N, D), self.norm.weight, self.norm.bias, self.norm.eps,
                                out_f32=True).reshape(B, N, D)
            out = self.pool(out)
        # [B, D] x [C, D]^T: a few MFLOP, left to the PyTorch library
        return self.head(out)

    def pool(self, x):
        return x[:, :-self.args.num_global_tokens, :].mean(dim=1)

    _POOL_KIND = 'global'

    def _pool_ranges(self, N):
        """Token row ranges whose LayerNorm-ed mean makes up the pooled feature, or None when ``pool`` has been
        overridden outside this module."""
        g = self.args.num_global_tokens
        known = {'global': MIRAGEClsGlobal.pool, 'cls': MIRAGEClsCLS.pool, 'token_mix': MIRAGEClsTokenMix.pool}
        if type(self).pool is not known.get(self._POOL_KIND):
            return None
        return {'global': [(0, N - g)], 'cls': [(N - g, N)], 'token_mix': [(0, N - g), (N - g, N)]}[self._POOL_KIND]

    def get_output_adapters(self):
        return None


@add_miragecls('cls')
class MIRAGEClsCLS(MIRAGEClsGlobal):
    _POOL_KIND = 'cls'

    def pool(self, x):
        return x[:, -self.args.num_global_tokens:, :].mean(dim=1)


@add_miragecls('token_mix')
class MIRAGEClsTokenMix(MIRAGEClsGlobal):
    _POOL_KIND = 'token_mix'

    def build_head(self, factor=2):
        super().build_head(factor)

    def pool(self, x):
        patch = x[:, :-self.args.num_global_tokens, :].mean(dim=1)
        global_ = x[:, -self.args.num_global_tokens:, :].mean(dim=1)
        return torch.cat([patch, global_], dim=1)
