"""MIRAGE MultiMAE encoder with the reference's API (mirage/model.py), running on B200 kernels.

``MIRAGEModel``  masking forward (pretraining, mirage_wrapper.MIRAGEWrapper, cls wrappers)
``MIRAGELight``  "MultiViT": no masking (HF wrapper, segmentation tuning)
``model_factory['miragepre_base' | 'miragepre_large' | 'miragelight_base' | 'miragelight_large']``

Constructor signatures, attribute names (``input_adapters``, ``output_adapters``, ``global_tokens``,
``encoder``), ``state_dict`` keys, forward signatures and return structures follow the reference, so
checkpoints and calling code carry over.  What differs is underneath: tokens live in one flat fp32
[B*N, D] residual stream, the per-modality tokens are written straight into their slice of the
concatenated buffer by the patch-embedding GEMM epilogue, visible-token selection is one
vectorised row-gather kernel, and every Block is one fused autograd node (functional._Block).
"""
from __future__ import annotations

import os

import itertools
import math
from collections import OrderedDict
from functools import partial
from typing import Dict, List, Optional, Union

import torch
from torch import Tensor, nn
from torch.distributions.dirichlet import Dirichlet

from . import functional as Fn
from . import ops
from .factory import get_factory_adder
from .utils import Block, trunc_normal_

add_model, model_factory = get_factory_adder()


class MIRAGEModel(nn.Module):
    """Reference: mirage/model.py:22-431."""

    def __init__(self, args, input_adapters: Dict[str, nn.Module],
                 output_adapters: Optional[Dict[str, nn.Module]], num_global_tokens: int = 1,
                 dim_tokens: int = 768, depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0,
                 qkv_bias: bool = True, drop_rate: float = 0.0, attn_drop_rate: float = 0.0,
                 drop_path_rate: float = 0.0, norm_layer=partial(nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.args = args
        for adapter in input_adapters.values():
            adapter.init(dim_tokens=dim_tokens)
        self.input_adapters = nn.ModuleDict(input_adapters)
        if output_adapters is not None:
            for adapter in output_adapters.values():
                adapter.init(dim_tokens_enc=dim_tokens)
            self.output_adapters = nn.ModuleDict(output_adapters)
        else:
            self.output_adapters = None

        self.num_global_tokens = num_global_tokens
        self.global_tokens = nn.Parameter(torch.zeros(1, num_global_tokens, dim_tokens))
        trunc_normal_(self.global_tokens, std=0.02)

        rates = torch.linspace(0, drop_path_rate, depth).tolist()
        self.encoder = nn.Sequential(*[
            Block(dim=dim_tokens, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=rates[i], norm_layer=norm_layer)
            for i in range(depth)])
        self.dim_tokens = dim_tokens

        self._reference_init()
        self.input_info = None
        self.token_dist = None
        # 'reference': the reference's torch op sequence (bit-exact masks given the same generators);
        # 'device': one fused sampling kernel (same distribution, own random stream; see set_mask_sampler)
        self.mask_sampler = 'reference'
        # masked forward: embed the kept patches only (functional.embed_visible); MB_EMBED_VISIBLE=0 or
        # model.visible_embedding = False restores embed-everything-then-gather
        self.visible_embedding = os.environ.get('MB_EMBED_VISIBLE', '1') != '0'
        self._mask_rng = None

    # -- initialisation: same distributions as mirage/model.py:95-121 -----------------------------
    def _reference_init(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear):
                # fused qkv / kv projections are initialised as 3 / 2 separate Xavier matrices
                parts = 3 if 'qkv' in name else (2 if 'kv' in name else 0)
                if parts:
                    bound = math.sqrt(6. / float(m.weight.shape[0] // parts + m.weight.shape[1]))
                    nn.init.uniform_(m.weight, -bound, bound)
            elif isinstance(m, nn.Conv2d) and '.proj' in name:
                w = m.weight.data
                nn.init.xavier_uniform_(w.view([w.shape[0], -1]))  # patch projection as a Linear (MAE)

    def get_num_layers(self):
        return len(self.encoder)

    @torch.jit.ignore
    def no_weight_decay(self):
        skip = {'global_tokens'}
        for task, adapter in self.input_adapters.items():
            if hasattr(adapter, 'no_weight_decay'):
                skip |= {f'input_adapters.{task}.{n}' for n in adapter.no_weight_decay()}
        if self.output_adapters is not None:
            for task, adapter in self.output_adapters.items():
                if hasattr(adapter, 'no_weight_decay'):
                    skip |= {f'output_adapters.{task}.{n}' for n in adapter.no_weight_decay()}
        return skip

    # -- mask sampling ------------------------------------------------------------------------------
    def sample_alphas(self, B: int, n_tasks: int, alphas: Union[float, List[float], Tensor] = 1.0,
                      eps: float = 1e-5):
        """Uniformly pick a non-empty task subset per sample, then scale (mirage/model.py:145-166)."""
        choices = torch.Tensor([list(c) for c in itertools.product([0, 1], repeat=n_tasks)][1:])
        pick = torch.randint(0, len(choices), (B,))
        return torch.index_select(choices, 0, pick) * torch.tensor(alphas) + eps

    def generate_random_masks(self, input_tokens: Dict[str, Tensor], num_encoded_tokens: int,
                              alphas: Union[float, List[float], Tensor] = 1.0,
                              sample_tasks_uniformly: bool = False):
        """Dirichlet split of ``num_encoded_tokens`` over the tasks, random subset per task, global
        shuffle putting visible tokens first (mirage/model.py:168-239).

        The torch op sequence and RNG consumption (CPU generator for the Dirichlet draw, device
        generator for the uniform noise) are the reference's, so identical seeds give bit-identical
        ``task_masks`` / ``ids_keep`` / ``ids_restore``.  ``input_tokens`` is only inspected for
        shapes and device (it may hold meta-like placeholders).
        """
        first = next(iter(input_tokens.values()))
        B, device = first.shape[0], first.device
        counts = [t.shape[1] for t in input_tokens.values()]
        if self.mask_sampler == 'device' and first.is_cuda:
            return self._device_masks(input_tokens, num_encoded_tokens, alphas, sample_tasks_uniformly)

        if self.token_dist is None:
            total = sum(counts)
            dist = {d: t.shape[1] / total for d, t in input_tokens.items()}
            self.token_dist = dict(sorted(dist.items(), key=lambda kv: kv[1], reverse=True))

        alphas = [alphas] * len(input_tokens) if isinstance(alphas, float) else alphas
        if sample_tasks_uniformly:
            conc = self.sample_alphas(B, len(input_tokens), alphas=alphas)
            share = self._to_device_async(Dirichlet(conc).sample(), device)
        else:
            share = self._to_device_async(Dirichlet(torch.Tensor(alphas)).sample((B,)), device)
        per_task = (share * num_encoded_tokens).round().long()

        masks = []
        for i, n in enumerate(counts):
            noise = torch.rand(B, n, device=device)
            order = torch.argsort(noise, dim=1)
            rank = torch.arange(n, device=device).unsqueeze(0).expand(B, -1)
            rank = torch.gather(rank, dim=1, index=order)
            masks.append(torch.where(rank < per_task[:, i].unsqueeze(1), 0, 1))

        mask_all = torch.cat(masks, dim=1)
        ids_shuffle = torch.argsort(mask_all + torch.rand_like(mask_all.float()), dim=1)
        ids_restore = torch.argsort(ids_shuffle, dim=1)
        ids_keep = ids_shuffle[:, :num_encoded_tokens]

        # the per-task rounding need not sum to num_encoded_tokens: recompute the binary mask
        mask_all = torch.ones_like(mask_all)
        mask_all[:, :num_encoded_tokens] = 0
        mask_all = torch.gather(mask_all, dim=1, index=ids_restore)
        task_masks = dict(zip(input_tokens.keys(), torch.split(mask_all, counts, dim=1)))
        return task_masks, ids_keep, ids_restore

    # -- on-device sampler ------------------------------------------------------------
    def set_mask_sampler(self, kind: str = 'device', seed: Optional[int] = None):
        """``'device'``: ``generate_random_masks`` becomes ONE kernel launch (csrc/masks.cu) -- Dirichlet split,
        per-task random subsets and the global visible-first order computed on the GPU from a Philox stream
        keyed by ``seed`` (default: drawn once from torch's CPU generator) and a device-resident draw counter
        the kernel advances itself, so the call is CUDA-graph capturable and every replay draws fresh masks.
        Same distribution as the reference sampler (tests/test_gpu_masks.py), different random numbers.
        ``'reference'`` restores the bit-exact reference path."""
        if kind not in ('reference', 'device'):
            raise ValueError(f'unknown mask sampler {kind!r}')
        self.mask_sampler = kind
        if kind == 'device':
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            dev = self.global_tokens.device
            self._mask_rng = {'seed': int(seed), 'draw': torch.zeros(1, dtype=torch.int64, device=dev),
                              'done': torch.zeros(1, dtype=torch.int32, device=dev)}
        return self

    def _device_masks(self, input_tokens: Dict[str, Tensor], num_encoded_tokens: int, alphas,
                      sample_tasks_uniformly: bool):
        first = next(iter(input_tokens.values()))
        B = first.shape[0]
        counts = [t.shape[1] for t in input_tokens.values()]
        rng = self._mask_rng
        if rng is None or rng['draw'].device != first.device:
            self.set_mask_sampler('device', seed=None if rng is None else rng['seed'])
            rng = self._mask_rng
            if rng['draw'].device != first.device:
                rng['draw'] = rng['draw'].to(first.device)
                rng['done'] = rng['done'].to(first.device)
        if isinstance(alphas, Tensor):
            alphas = alphas.flatten().tolist()
        al = [float(alphas)] * len(counts) if isinstance(alphas, (int, float)) else [float(a) for a in alphas]
        mask_all, ids_keep, ids_restore = ops.sample_masks(rng['seed'], rng['draw'], rng['done'], counts, al, B,
                                                           num_encoded_tokens, sample_tasks_uniformly)
        task_masks = dict(zip(input_tokens.keys(), torch.split(mask_all, counts, dim=1)))
        return task_masks, ids_keep, ids_restore

    def _to_device_async(self, t: Tensor, device) -> Tensor:
        """Host -> device copy of the (CPU-sampled, as in the reference) Dirichlet shares without stalling
        the host: a pageable ``.to(device)`` makes the CPU wait until the GPU has drained everything queued
        before it, i.e. the previous training step, so step k+1 could not be enqueued while step k runs.
        The values, and therefore the masks, are unchanged.  A small ring of pinned staging buffers is
        reused; a slot is recycled only after its previous copy has completed."""
        if torch.device(device).type != 'cuda':
            return t.to(device)
        ring = getattr(self, '_share_ring', None)
        if ring is None or ring['shape'] != tuple(t.shape) or ring['dtype'] != t.dtype:
            ring = {'shape': tuple(t.shape), 'dtype': t.dtype, 'i': 0,
                    'buf': [torch.empty(t.shape, dtype=t.dtype).pin_memory() for _ in range(4)],
                    'ev': [None] * 4}
            self._share_ring = ring
        i = ring['i']
        ring['i'] = (i + 1) % 4
        if ring['ev'][i] is not None:
            ring['ev'][i].synchronize()
        ring['buf'][i].copy_(t)
        out = ring['buf'][i].to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        ring['ev'][i] = ev
        return out

    @staticmethod
    def make_mask(N_H, N_W, xy_idxs, full_tasks=[], indicate_visible=True, flatten=True, device='cuda'):
        """Per-task masks from lists of visible (x, y) patch coordinates (mirage/model.py:241-277)."""
        out = {}
        for k, v in xy_idxs.items():
            m = torch.ones(N_H, N_W).to(device)
            idx = torch.LongTensor(v)
            if len(idx) > 0:
                m[idx[:, 1], idx[:, 0]] = 0
            out[k] = m
        for task in full_tasks:
            out[task][:] = 0
        if not indicate_visible:
            out = {k: 1 - v for k, v in out.items()}
        if flatten:
            out = {k: v.flatten().unsqueeze(0) for k, v in out.items()}
        return out

    def generate_input_info(self, input_task_tokens, image_size):
        """Token bookkeeping consumed by the output adapters (mirage/model.py:279-303).
        ``input_task_tokens`` maps domain -> tensor or int (number of tokens)."""
        info = OrderedDict()
        info['tasks'] = {}
        start = 0
        for domain, t in input_task_tokens.items():
            n = t if isinstance(t, int) else t.shape[1]
            d = {'num_tokens': n, 'has_posemb': True, 'start_idx': start, 'end_idx': start + n}
            if isinstance(image_size, dict):
                d['image_size'] = image_size[domain]
            if self.args.grid_sizes is not None:
                d['grid_size'] = self.args.grid_sizes[domain]
            start += n
            info['tasks'][domain] = d
        if isinstance(image_size, int):
            info['image_size'] = image_size
        info['num_task_tokens'] = start
        info['num_global_tokens'] = self.num_global_tokens
        return info

    # -- tokenisation -------------------------------------------------------------------------------
    def _token_counts(self, x: Dict[str, Tensor]) -> Dict[str, int]:
        counts = OrderedDict()
        for domain, t in x.items():
            if domain not in self.input_adapters:
                continue
            ad = self.input_adapters[domain]
            H, W = t.shape[-2:]
            counts[domain] = (H // ad.P_H) * (W // ad.P_W)
        return counts

    def _embed_all(self, x: Dict[str, Tensor], extra_rows: int = 0):
        """All modalities -> one fp32 buffer [B, N_all + extra_rows, D] (the reference's per-adapter
        forward + torch.cat, model.py:352-356/:384).  Without autograd every adapter's GEMM epilogue
        writes its slice in place; with autograd the adapters' Functions run and are concatenated."""
        counts = self._token_counts(x)
        B = next(iter(x.values())).shape[0]
        n_all = sum(counts.values())
        params = [p for d in counts for p in self.input_adapters[d].parameters()]
        if Fn.grad_needed(*params):
            toks = [self.input_adapters[d](x[d]) for d in counts]
            if extra_rows:
                toks.append(toks[0].new_zeros(B, extra_rows, self.dim_tokens))
            return torch.cat(toks, dim=1), counts
        dev = self.global_tokens.device
        buf = torch.empty((B, n_all + extra_rows, self.dim_tokens), dtype=torch.float32, device=dev)
        off = 0
        flat = buf.view(B * (n_all + extra_rows), self.dim_tokens)
        for d, n in counts.items():
            self.input_adapters[d].write_tokens(x[d], flat, (n, n_all + extra_rows, off))
            off += n
        return buf, counts

    def _run_encoder(self, x2, B, N, collect=False):
        outs = []
        for blk in self.encoder:
            x2 = blk.forward_flat(x2, B, N)
            if collect:
                outs.append(x2)
        return outs if collect else x2

    # -- forward ------------------------------------------------------------------------------------
    def forward(self, x: Union[Dict[str, Tensor], Tensor], mask_inputs: bool = True,
                task_masks: Optional[Dict[str, Tensor]] = None, num_encoded_tokens: int = 128,
                alphas: Union[float, List[float]] = 1.0, sample_tasks_uniformly: bool = False,
                return_all_layers: bool = False, reshape: bool = False):
        """Input adapters -> (random | given) masking -> encoder -> output adapters.
        Returns ``(preds | encoder_tokens | features, task_masks)`` as mirage/model.py:305-431."""
        x = {'bscan': x} if isinstance(x, Tensor) else x
        counts = self._token_counts(x)
        n_all = sum(counts.values())
        B = next(iter(x.values())).shape[0]
        D = self.dim_tokens
        if self.input_info is None:
            self.input_info = self.generate_input_info(dict(counts), image_size=self.args.input_size)
        input_info = self.input_info

        if not mask_inputs:
            num_encoded_tokens = n_all

        # The masks are drawn from the token COUNTS alone (model.py:168-239), so they can be sampled before the
        # input adapters run -- and then only the kept patches need embedding (functional.embed_visible).
        if task_masks is None:
            shapes = {d: self.global_tokens.new_empty((B, n, 0)) for d, n in counts.items()}
            task_masks, ids_keep, ids_restore = self.generate_random_masks(
                shapes, num_encoded_tokens, alphas=alphas, sample_tasks_uniformly=sample_tasks_uniformly)
        else:
            # reference semantics incl. its batch-wide count of visible tokens (model.py:379-382)
            mask_all = torch.cat([task_masks[t] for t in counts], dim=1)
            ids_shuffle = torch.argsort(mask_all, dim=1)
            ids_restore = torch.argsort(ids_shuffle, dim=1)
            ids_keep = ids_shuffle[:, :(mask_all == 0).sum()]

        n_glob = self.num_global_tokens
        specs = None
        if (self.visible_embedding and mask_inputs and 0 < ids_keep.shape[1] < n_all and n_glob > 0
                and len(counts) <= 4 and len(counts) + n_glob <= 8 and D <= 1024 and D % 4 == 0):   # limits of visible.cu
            specs = [self.input_adapters[d].visible_spec(x[d]) if hasattr(self.input_adapters[d], 'visible_spec')
                     else None for d in counts]
        if specs is not None and all(sp is not None for sp in specs) and len(specs) <= 4:
            tok = Fn.embed_visible([sp[0] for sp in specs], [t for sp in specs for t in sp[1]], ids_keep,
                                   self.global_tokens).reshape(B, ids_keep.shape[1] + n_glob, D)
        else:
            tokens_all, counts = self._embed_all(x)
            tok = Fn.token_gather(tokens_all, ids_keep, self.global_tokens[0])     # [B, n_keep + n_glob, D]
        N = tok.shape[1]
        x2 = tok.reshape(B * N, D)

        if return_all_layers:
            gh, gw = self.args.grid_sizes['bscan']
            feats = OrderedDict()
            for i, t in enumerate(self._run_encoder(x2, B, N, collect=True)):
                t = t.reshape(B, N, D)[:, :-n_glob]
                feats[f'layer_{i}'] = t.reshape(B, gh, gw, D).permute(0, 3, 1, 2)
            return feats

        enc = self._run_encoder(x2, B, N).reshape(B, N, D)

        if self.output_adapters is None:
            if reshape:
                gh, gw = self.args.grid_sizes['bscan']
                enc = enc[:, :-n_glob].reshape(B, gh, gw, D).permute(0, 3, 1, 2)
            return enc, task_masks

        preds = {
            domain: self.output_adapters[domain](encoder_tokens=enc, input_info=input_info,
                                                 ids_keep=ids_keep, ids_restore=ids_restore)
            for domain in self.output_adapters
        }
        return preds, task_masks


def _sized(cls, size):
    cfg = {'base': dict(dim_tokens=768, depth=12, num_heads=12),
           'large': dict(dim_tokens=1024, depth=24, num_heads=16)}[size]

    def build(input_adapters: Dict[str, nn.Module], output_adapters: Optional[Dict[str, nn.Module]],
              args, **kwargs):
        return cls(args, input_adapters=input_adapters, output_adapters=output_adapters, mlp_ratio=4,
                   qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **cfg, **kwargs)
    return build


class MIRAGELight(MIRAGEModel):
    """MultiViT: MIRAGE without masking (mirage/model.py:478-567, hf/mirage_hf.py:363-579)."""

    def process_input(self, x):
        x = {'bscan': x} if isinstance(x, Tensor) else x
        if 'bscan' in x:
            _, _, H, W = x['bscan'].shape
        elif 'semseg' in x:
            _, H, W = x['semseg'].shape
            H *= self.input_adapters['semseg'].stride_level
            W *= self.input_adapters['semseg'].stride_level
        else:
            _, _, H, W = list(x.values())[0].shape
        n_glob = self.num_global_tokens
        params = [p for d in x if d in self.input_adapters for p in self.input_adapters[d].parameters()]
        if Fn.grad_needed(self.global_tokens, *params):
            buf, counts = self._embed_all(x)
            buf = torch.cat([buf, self.global_tokens.expand(buf.shape[0], -1, -1)], dim=1)
        else:
            # tokens and global rows written in place: no concat copies on the inference path
            buf, counts = self._embed_all(x, extra_rows=n_glob)
            ops.fill_global_rows(self.global_tokens.detach()[0], buf, sum(counts.values()))
        input_info = self.generate_input_info(dict(counts), image_size=(H, W))
        return buf, input_info

    def forward(self, x: Union[Dict[str, Tensor], Tensor], return_all_layers=False, **kwargs):
        tokens, input_info = self.process_input(x)
        B, N, D = tokens.shape
        x2 = tokens.reshape(B * N, D)
        if not return_all_layers:
            encoder_tokens = self._run_encoder(x2, B, N).reshape(B, N, D)
        else:
            encoder_tokens = [t.reshape(B, N, D) for t in self._run_encoder(x2, B, N, collect=True)]
        if self.output_adapters is None:
            return encoder_tokens
        return {domain: self.output_adapters[domain](encoder_tokens=encoder_tokens, input_info=input_info)
                for domain in self.output_adapters}


add_model('miragepre_base')(_sized(MIRAGEModel, 'base'))
add_model('miragepre_large')(_sized(MIRAGEModel, 'large'))
add_model('miragelight_base')(_sized(MIRAGELight, 'base'))
add_model('miragelight_large')(_sized(MIRAGELight, 'large'))
