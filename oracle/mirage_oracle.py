"""CPU oracle for the MIRAGE MultiViT hot path -- TEST INFRASTRUCTURE ONLY.

A plain fp32 PyTorch restatement (functional, driven by a reference-compatible ``state_dict``) of
the arithmetic on the path BASELINE.json names.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product
package ``mirage_b200`` never does.

Parity status: PINNED against the reference itself.  ``oracle/make_golden.py`` imports the
unmodified reference from ``/root/reference`` (this container only), runs it on seeded inputs and
synthetic weights, and commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function here against those vectors.  The reference ships no tests or golden vectors
of its own, and all its arithmetic on this path is PyTorch library code
(torch 2.5.1 pinned by the reference's requirements.txt:52; 2.11.0 here).

Each function cites the reference file:line it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-6  # norm_layer=partial(nn.LayerNorm, eps=1e-6), mirage/model.py:57


# --------------------------------------------------------------------------------------------
# positional embedding
# --------------------------------------------------------------------------------------------
def sincos_posemb_2d(h: int, w: int, dim: int, temperature: float = 10000.0) -> Tensor:
    """Fixed 2-D sin-cos table, returned as [1, dim, h, w].

    mirage/utils.py:24-41.  Note the reference's axis convention: the first quarter pairs
    (sin, cos) are driven by the index that runs along the FIRST spatial axis after the final
    reshape (it builds the mesh with indexing='ij' over (w, h) and then reads the flat axis as
    (h w)).
    """
    assert dim % 4 == 0
    quarter = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(quarter, dtype=torch.float32) / quarter))
    a = torch.arange(w, dtype=torch.float32)
    b = torch.arange(h, dtype=torch.float32)
    # flat index f = i * h + j with i in [0, w), j in [0, h)  (meshgrid 'ij' over (w, h))
    first = a.repeat_interleave(h)            # value of the w-grid at flat position f
    second = b.repeat(w)                      # value of the h-grid at flat position f
    ang1 = first[:, None] * omega[None, :]
    ang2 = second[:, None] * omega[None, :]
    table = torch.cat([ang1.sin(), ang1.cos(), ang2.sin(), ang2.cos()], dim=1)  # [w*h, dim]
    # the reference then reinterprets the flat axis as (h w) and moves dim first
    return table.reshape(1, h, w, dim).permute(0, 3, 1, 2).contiguous()


def _tokens_from_posemb(pos: Tensor, nh: int, nw: int, mode: str) -> Tensor:
    """[1, D, h, w] table -> [nh*nw, D] rows, resized if the token grid differs.

    mirage/input_adapters.py:104-105 (bicubic), :232-233 (bilinear); output_adapters.py:176-177.
    """
    if pos.shape[-2:] != (nh, nw):
        pos = F.interpolate(pos, size=(nh, nw), mode=mode, align_corners=False)
    else:
        # identity at the native grid for both modes (checked in tests against F.interpolate)
        pos = pos
    return pos[0].flatten(1).t().contiguous()


# --------------------------------------------------------------------------------------------
# input adapters
# --------------------------------------------------------------------------------------------
def patch_embed(img: Tensor, weight: Tensor, bias: Tensor, pos_emb: Tensor) -> Tensor:
    """PatchedInputAdapter.forward, mirage/input_adapters.py:87-110.

    img [B, C, H, W]; weight [D, C, P, P]; returns [B, (H/P)*(W/P), D].  The stride-P convolution
    is a GEMM over flattened patches in (c, ph, pw) order.
    """
    B, Cin, H, W = img.shape
    D, _, P, Q = weight.shape
    nh, nw = H // P, W // Q
    patches = img.reshape(B, Cin, nh, P, nw, Q).permute(0, 2, 4, 1, 3, 5).reshape(B, nh * nw, Cin * P * Q)
    tok = patches @ weight.reshape(D, -1).t() + bias
    return tok + _tokens_from_posemb(pos_emb, nh, nw, "bicubic")


def semseg_embed(labels: Tensor, class_emb: Tensor, weight: Tensor, bias: Tensor, pos_emb: Tensor) -> Tensor:
    """SemSegInputAdapter.forward, mirage/input_adapters.py:211-238.

    labels [B, H, W] int64; class_emb [n_cls, E]; weight [D, E, P, P]; returns [B, N, D].
    """
    B, H, W = labels.shape
    D, E, P, Q = weight.shape
    nh, nw = H // P, W // Q
    emb = class_emb[labels]                                   # [B, H, W, E]
    patches = emb.reshape(B, nh, P, nw, Q, E).permute(0, 1, 3, 5, 2, 4).reshape(B, nh * nw, E * P * Q)
    tok = patches @ weight.reshape(D, -1).t() + bias
    return tok + _tokens_from_posemb(pos_emb, nh, nw, "bilinear")


def semseg_embed_interp(labels: Tensor, class_emb: Tensor, weight: Tensor, bias: Tensor, pos_emb: Tensor,
                        patch: Tuple[int, int]) -> Tensor:
    """SemSegInputAdapter.forward with interpolate_class_emb=True, mirage/input_adapters.py:194-200, :226-238:
    the embedded class map is down-sampled bilinearly by the patch size (nn.Upsample(scale_factor=1/P),
    align_corners=False) and projected by a 1x1 convolution.  weight [D, E, 1, 1]; returns [B, N, D]."""
    B, H, W = labels.shape
    D, E = weight.shape[:2]
    nh, nw = H // patch[0], W // patch[1]
    emb = class_emb[labels].permute(0, 3, 1, 2)                # [B, E, H, W]
    small = F.interpolate(emb, scale_factor=(1 / patch[0], 1 / patch[1]), mode="bilinear")   # [B, E, nh, nw]
    tok = small.permute(0, 2, 3, 1).reshape(B, nh * nw, E) @ weight.reshape(D, E).t() + bias
    return tok + _tokens_from_posemb(pos_emb, nh, nw, "bilinear")


# --------------------------------------------------------------------------------------------
# transformer pieces
# --------------------------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """nn.LayerNorm(eps=1e-6) as used at mirage/utils.py:241,250,260-261."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + LN_EPS) * w + b


def gelu(x: Tensor) -> Tensor:
    """nn.GELU() (exact erf form), mirage/utils.py:143,156."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def sdpa(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """softmax(q k^T / sqrt(hd)) v without mask or dropout; mirage/utils.py:181-185, :216-220.

    q [B, H, Nq, hd], k/v [B, H, Nk, hd].
    """
    s = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    return torch.softmax(s, dim=-1) @ v


def self_attention(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """Attention.forward, mirage/utils.py:174-188."""
    B, N, D = x.shape
    qkv = x @ sd[pre + "qkv.weight"].t() + sd[pre + "qkv.bias"]
    qkv = qkv.reshape(B, N, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
    o = sdpa(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(B, N, D)
    return o @ sd[pre + "proj.weight"].t() + sd[pre + "proj.bias"]


def cross_attention(x: Tensor, ctx: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """CrossAttention.forward, mirage/utils.py:205-223."""
    B, N, D = x.shape
    M = ctx.shape[1]
    q = (x @ sd[pre + "q.weight"].t() + sd[pre + "q.bias"]).reshape(B, N, heads, D // heads).transpose(1, 2)
    kv = (ctx @ sd[pre + "kv.weight"].t() + sd[pre + "kv.bias"]).reshape(B, M, 2, heads, D // heads)
    k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
    o = sdpa(q, k, v).transpose(1, 2).reshape(B, N, D)
    return o @ sd[pre + "proj.weight"].t() + sd[pre + "proj.bias"]


def mlp(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """Mlp.forward (dropout p=0), mirage/utils.py:154-159."""
    h = gelu(x @ sd[pre + "fc1.weight"].t() + sd[pre + "fc1.bias"])
    return h @ sd[pre + "fc2.weight"].t() + sd[pre + "fc2.bias"]


def block(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """Block.forward (DropPath = identity at rate 0), mirage/utils.py:259-262."""
    x = x + self_attention(layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]), sd, pre + "attn.", heads)
    x = x + mlp(layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"]), sd, pre + "mlp.")
    return x


def encoder(x: Tensor, sd: Dict[str, Tensor], depth: int, heads: int, pre: str = "encoder.",
            return_all_layers: bool = False):
    """self.encoder = nn.Sequential(Block...), mirage/model.py:81-93, :409, :545-553."""
    outs = []
    for i in range(depth):
        x = block(x, sd, f"{pre}{i}.", heads)
        outs.append(x)
    return outs if return_all_layers else x


# --------------------------------------------------------------------------------------------
# input stage + masking
# --------------------------------------------------------------------------------------------
def embed_inputs(x: Dict[str, Tensor], sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Per-domain tokenisation in dict order, mirage/model.py:352-356 / :513-517."""
    toks = {}
    for d, t in x.items():
        p = f"input_adapters.{d}."
        if p + "class_emb.weight" in sd:
            toks[d] = semseg_embed(t, sd[p + "class_emb.weight"], sd[p + "proj.weight"], sd[p + "proj.bias"],
                                   sd[p + "pos_emb"])
        else:
            toks[d] = patch_embed(t, sd[p + "proj.weight"], sd[p + "proj.bias"], sd[p + "pos_emb"])
    return toks


def random_masks(num_tokens: Sequence[int], batch: int, num_visible: int, alphas=1.0,
                 device: torch.device | str = "cpu") -> Tuple[List[Tensor], Tensor, Tensor]:
    """MIRAGEModel.generate_random_masks (sample_tasks_uniformly=False), mirage/model.py:202-239.

    Consumes the global CPU generator (Dirichlet) and the generator of ``device`` (uniform noise)
    in exactly the reference's order, so under the same seeds the outputs are bit-identical.
    Returns (per-task masks [B, n_t] int64 with 0 = visible, ids_keep [B, num_visible],
    ids_restore [B, sum n_t]).
    """
    n_tasks = len(num_tokens)
    conc = [alphas] * n_tasks if isinstance(alphas, float) else alphas
    share = torch.distributions.Dirichlet(torch.Tensor(conc)).sample((batch,)).to(device)
    want = (share * num_visible).round().long()                          # :209
    per_task = []
    for i, n in enumerate(num_tokens):
        noise = torch.rand(batch, n, device=device)                      # :216
        order = torch.argsort(noise, dim=1)                              # :217
        ranks = torch.arange(n, device=device).unsqueeze(0).expand(batch, -1)
        ranks = torch.gather(ranks, dim=1, index=order)                  # :218-219
        per_task.append(torch.where(ranks < want[:, i].unsqueeze(1), 0, 1))   # :221
    flat = torch.cat(per_task, dim=1)
    shuffle = torch.argsort(flat + torch.rand_like(flat.float()), dim=1)      # :225
    restore = torch.argsort(shuffle, dim=1)                                   # :226
    keep = shuffle[:, :num_visible]
    final = torch.ones_like(flat)
    final[:, :num_visible] = 0
    final = torch.gather(final, dim=1, index=restore)                         # :230-233
    return list(torch.split(final, list(num_tokens), dim=1)), keep, restore


def select_visible(tokens: Tensor, ids_keep: Tensor, global_tokens: Tensor) -> Tensor:
    """gather(ids_keep) then append the global token LAST, mirage/model.py:387-391."""
    B, _, D = tokens.shape
    vis = torch.gather(tokens, 1, ids_keep.unsqueeze(-1).expand(-1, -1, D))
    return torch.cat([vis, global_tokens.expand(B, -1, -1)], dim=1)


def light_forward(x: Dict[str, Tensor], sd: Dict[str, Tensor], depth: int, heads: int,
                  return_all_layers: bool = False):
    """MIRAGELight.forward with output_adapters=None, mirage/model.py:497-556 (hf/mirage_hf.py:510-568)."""
    toks = torch.cat(list(embed_inputs(x, sd).values()), dim=1)
    B = toks.shape[0]
    toks = torch.cat([toks, sd["global_tokens"].expand(B, -1, -1)], dim=1)
    return encoder(toks, sd, depth, heads, return_all_layers=return_all_layers)


# --------------------------------------------------------------------------------------------
# spatial output adapter (decoder)
# --------------------------------------------------------------------------------------------
def decoder_queries_and_context(ctx: Tensor, sd: Dict[str, Tensor], pre: str, task: str,
                                in_tasks: Sequence[str], tokens_per_task: Sequence[int],
                                grid: Tuple[int, int], ids_keep: Tensor, ids_restore: Tensor,
                                n_global: int = 1) -> Tuple[Tensor, Tensor]:
    """SpatialOutputAdapter.get_queries_and_context + generate_context_embeddings,
    mirage/output_adapters.py:164-246 (use_task_queries=True and task among the inputs).

    ctx [B, n_vis + n_global, Dd] are the projected encoder tokens.
    """
    B, _, Dd = ctx.shape
    n_all = int(sum(tokens_per_task))
    body = ctx[:, :-n_global] if n_global else ctx
    fill = sd[pre + "mask_token"].expand(B, n_all - body.shape[1], Dd)
    full = torch.cat([body, fill], dim=1)
    full = torch.gather(full, 1, ids_restore.unsqueeze(-1).expand(-1, -1, Dd))        # :206-207
    pos = _tokens_from_posemb(sd[pre + "pos_emb"], grid[0], grid[1], "bilinear")       # :176-177
    embs = []
    for t, n in zip(in_tasks, tokens_per_task):
        key = f"{pre}task_embeddings.{t}"
        te = sd[key].expand(1, n, Dd) if key in sd else torch.zeros(1, n, Dd, dtype=ctx.dtype)
        assert n == pos.shape[0]
        embs.append(te + pos)                                                          # :180
    full = full + torch.cat(embs, dim=1)                                               # :212
    start = int(sum(tokens_per_task[: list(in_tasks).index(task)]))
    queries = full[:, start:start + tokens_per_task[list(in_tasks).index(task)]]       # :216-218
    vis = torch.gather(full, 1, ids_keep.unsqueeze(-1).expand(-1, -1, Dd))             # :233-237
    context = torch.cat([vis, ctx[:, -n_global:]], dim=1) if n_global else vis         # :240-242
    return queries, context


def spatial_output_adapter(enc: Tensor, sd: Dict[str, Tensor], pre: str, task: str,
                           in_tasks: Sequence[str], tokens_per_task: Sequence[int],
                           grid: Tuple[int, int], patch: Tuple[int, int], channels: int,
                           ids_keep: Tensor, ids_restore: Tensor, heads: int = 8, depth: int = 2,
                           n_global: int = 1) -> Tensor:
    """SpatialOutputAdapter.forward, mirage/output_adapters.py:248-296 (use_xattn=True)."""
    ctx = enc @ sd[pre + "proj_context.weight"].t() + sd[pre + "proj_context.bias"]    # :272
    q, c = decoder_queries_and_context(ctx, sd, pre, task, in_tasks, tokens_per_task, grid,
                                       ids_keep, ids_restore, n_global)
    qn = layer_norm(q, sd[pre + "query_norm.weight"], sd[pre + "query_norm.bias"])
    cn = layer_norm(c, sd[pre + "context_norm.weight"], sd[pre + "context_norm.bias"])
    x = cross_attention(qn, cn, sd, pre + "decoder.", heads)                           # :279 (no residual)
    x = x + mlp(layer_norm(x, sd[pre + "out_norm.weight"], sd[pre + "out_norm.bias"]), sd, pre + "mlp.")  # :280
    for j in range(depth):
        x = block(x, sd, f"{pre}decoder_transformer.{j}.", heads)                      # :285
    x = x @ sd[pre + "out_proj.weight"].t() + sd[pre + "out_proj.bias"]                # :288
    B = x.shape[0]
    nh, nw = grid
    ph, pw = patch
    x = x.reshape(B, nh, nw, channels, ph, pw).permute(0, 3, 1, 4, 2, 5)               # :291-294
    return x.reshape(B, channels, nh * ph, nw * pw)


# --------------------------------------------------------------------------------------------
# masked criteria
# --------------------------------------------------------------------------------------------
def _upsampled_mask(mask: Tensor, H: int, W: int, scale: int) -> Tensor:
    nh, nw = H // scale, W // scale
    m = mask.reshape(mask.shape[0], nh, nw).float()
    return m.repeat_interleave(scale, dim=1).repeat_interleave(scale, dim=2)


def masked_mse(pred: Tensor, target: Tensor, mask: Optional[Tensor], patch: int) -> Tensor:
    """MaskedMSELoss.forward (norm_pix=False), mirage/criterion.py:87-117."""
    se = (pred - target) ** 2
    if mask is None:
        return se.mean()
    if mask.sum() == 0:
        return torch.tensor(0)
    H, W = pred.shape[-2:]
    m = _upsampled_mask(mask, H, W, patch)
    per = (se.mean(dim=1) * m).flatten(1).sum(1) / m.flatten(1).sum(1)
    return per.nanmean()


def masked_ce(logits: Tensor, target: Tensor, mask: Optional[Tensor], patch: int,
              label_smoothing: float = 0.0) -> Tensor:
    """MaskedCrossEntropyLoss.forward, mirage/criterion.py:31-51."""
    logp = torch.log_softmax(logits, dim=1)
    # F.cross_entropy(reduction='none') with its default ignore_index = -100: such pixels get zero loss (the
    # smoothing term included) but are still counted by the mask sum below
    keep = target != -100
    nll = -logp.gather(1, target.clamp(min=0).unsqueeze(1)).squeeze(1)
    if label_smoothing > 0.0:
        nll = (1.0 - label_smoothing) * nll + label_smoothing * (-logp.mean(dim=1))
    nll = torch.where(keep, nll, torch.zeros_like(nll))
    if mask is None:
        return nll.mean()
    if mask.sum() == 0:
        return torch.tensor(0)
    H, W = logits.shape[-2:]
    m = _upsampled_mask(mask, H, W, patch)
    per = (nll * m).flatten(1).sum(1) / m.flatten(1).sum(1)
    return per.nanmean()


# --------------------------------------------------------------------------------------------
# whole pretraining forward
# --------------------------------------------------------------------------------------------
DOMAIN_SPEC = {
    # domain: (output channels, patch size on the domain's own grid); mirage_wrapper.py:22-44, :73-78
    "bscan": (1, (32, 32)),
    "slo": (1, (32, 32)),
    "bscanlayermap": (13, (8, 8)),
}


def pretrain_forward(x: Dict[str, Tensor], sd: Dict[str, Tensor], depth: int, heads: int,
                     masks: Tuple[List[Tensor], Tensor, Tensor], out_domains: Sequence[str],
                     dec_heads: int = 8, dec_depth: int = 2, grid: Tuple[int, int] = (16, 16)):
    """MIRAGEModel.forward with given (task_masks, ids_keep, ids_restore), mirage/model.py:352-431."""
    toks = embed_inputs(x, sd)
    in_tasks = list(toks.keys())
    per_task = [t.shape[1] for t in toks.values()]
    _, ids_keep, ids_restore = masks
    vis = select_visible(torch.cat(list(toks.values()), dim=1), ids_keep, sd["global_tokens"])
    enc = encoder(vis, sd, depth, heads)
    preds = {}
    for d in out_domains:
        ch, patch = DOMAIN_SPEC[d]
        preds[d] = spatial_output_adapter(enc, sd, f"output_adapters.{d}.", d, in_tasks, per_task, grid, patch,
                                          ch, ids_keep, ids_restore, heads=dec_heads, depth=dec_depth)
    return preds, enc


def pretrain_loss(preds: Dict[str, Tensor], targets: Dict[str, Tensor], task_masks: Dict[str, Tensor]):
    """Per-task masked losses summed, run_pretraining.py:714-723."""
    losses = {}
    for d, p in preds.items():
        ch, patch = DOMAIN_SPEC[d]
        if ch == 1:
            losses[d] = masked_mse(p.float(), targets[d], task_masks[d], patch[0])
        else:
            losses[d] = masked_ce(p.float(), targets[d], task_masks[d], patch[0])
    return sum(losses.values()), losses


def cls_head(tokens: Tensor, sd: Dict[str, Tensor], pool: str = "global", n_global: int = 1) -> Tensor:
    """MIRAGEClsGlobal/CLS/TokenMix forward tail, mirage_wrapper.py:219-244."""
    t = layer_norm(tokens, sd["norm.weight"], sd["norm.bias"])
    patch = t[:, :-n_global].mean(dim=1)
    glob = t[:, -n_global:].mean(dim=1)
    feat = {"global": patch, "cls": glob, "token_mix": torch.cat([patch, glob], dim=1)}[pool]
    return feat @ sd["head.weight"].t() + sd["head.bias"]


# --------------------------------------------------------------------------------------------
# reference-compatible state_dict layouts (SURVEY.md 8(b) "state_dict keys"), built WITHOUT the product
# package: bench.py's CPU legs and the tests fill them with seeded synthetic weights.
# --------------------------------------------------------------------------------------------
MODEL_SIZES = {"tiny": (128, 2, 2), "base": (768, 12, 12), "large": (1024, 24, 16)}


def _block_keys(sd: Dict[str, Tensor], pre: str, dim: int, ratio: int = 4):
    """Block parameter layout, mirage/utils.py:236-257 (+ Attention :167-174, Mlp :143-151)."""
    z = torch.zeros
    sd[pre + "norm1.weight"], sd[pre + "norm1.bias"] = torch.ones(dim), z(dim)
    sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"] = z(3 * dim, dim), z(3 * dim)
    sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"] = z(dim, dim), z(dim)
    sd[pre + "norm2.weight"], sd[pre + "norm2.bias"] = torch.ones(dim), z(dim)
    sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"] = z(ratio * dim, dim), z(ratio * dim)
    sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"] = z(dim, ratio * dim), z(dim)


def light_state_dict_shapes(mods: Sequence[str], size: str, grid: Tuple[int, int] = (16, 16)) -> Dict[str, Tensor]:
    """Encoder-only (MIRAGELight / hf MIRAGEWrapper.model) state_dict: zero weights of the right shapes and the
    real frozen sin-cos ``pos_emb`` tables (input_adapters.py:66-72; model.py:61-87)."""
    dim, depth, _ = MODEL_SIZES[size]
    sd: Dict[str, Tensor] = {"global_tokens": torch.zeros(1, 1, dim)}
    for d in mods:
        pre = f"input_adapters.{d}."
        sd[pre + "pos_emb"] = sincos_posemb_2d(grid[0], grid[1], dim)
        if d == "bscanlayermap":
            sd[pre + "class_emb.weight"] = torch.zeros(13, 64)
            sd[pre + "proj.weight"], sd[pre + "proj.bias"] = torch.zeros(dim, 64, 8, 8), torch.zeros(dim)
        else:
            sd[pre + "proj.weight"], sd[pre + "proj.bias"] = torch.zeros(dim, 1, 32, 32), torch.zeros(dim)
    for i in range(depth):
        _block_keys(sd, f"encoder.{i}.", dim)
    return sd


def pretrain_state_dict_shapes(mods: Sequence[str], size: str, dec_dim: int = 256, dec_depth: int = 2,
                               grid: Tuple[int, int] = (16, 16)) -> Dict[str, Tensor]:
    """MIRAGEModel + one SpatialOutputAdapter per task (output_adapters.py:50-145)."""
    dim = MODEL_SIZES[size][0]
    sd = light_state_dict_shapes(mods, size, grid)
    z = torch.zeros
    for d in mods:
        pre = f"output_adapters.{d}."
        sd[pre + "mask_token"] = z(1, 1, dec_dim)
        sd[pre + "pos_emb"] = sincos_posemb_2d(grid[0], grid[1], dec_dim)
        for t in mods:
            sd[f"{pre}task_embeddings.{t}"] = z(1, 1, dec_dim)
        sd[pre + "decoder.q.weight"], sd[pre + "decoder.q.bias"] = z(dec_dim, dec_dim), z(dec_dim)
        sd[pre + "decoder.kv.weight"], sd[pre + "decoder.kv.bias"] = z(2 * dec_dim, dec_dim), z(2 * dec_dim)
        sd[pre + "decoder.proj.weight"], sd[pre + "decoder.proj.bias"] = z(dec_dim, dec_dim), z(dec_dim)
        for n in ("context_norm", "query_norm", "out_norm"):
            sd[f"{pre}{n}.weight"], sd[f"{pre}{n}.bias"] = torch.ones(dec_dim), z(dec_dim)
        sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"] = z(4 * dec_dim, dec_dim), z(4 * dec_dim)
        sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"] = z(dec_dim, 4 * dec_dim), z(dec_dim)
        for j in range(dec_depth):
            _block_keys(sd, f"{pre}decoder_transformer.{j}.", dec_dim)
        out = (13 * 8 * 8) if d == "bscanlayermap" else 32 * 32
        sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"] = z(out, dec_dim), z(out)
        sd[pre + "proj_context.weight"], sd[pre + "proj_context.bias"] = z(dec_dim, dim), z(dec_dim)
    return sd


def cls_state_dict_shapes(size: str, num_classes: int = 5, factor: int = 1) -> Dict[str, Tensor]:
    """miragecls_factory[...] (mirage_wrapper.py:190-206): encoder under ``model.``, ``norm`` and ``head``."""
    dim = MODEL_SIZES[size][0]
    sd = {"model." + k: v for k, v in light_state_dict_shapes(["bscan"], size).items()}
    sd["norm.weight"], sd["norm.bias"] = torch.ones(dim), torch.zeros(dim)
    sd["head.weight"], sd["head.bias"] = torch.zeros(num_classes, dim * factor), torch.zeros(num_classes)
    return sd


# --------------------------------------------------------------------------------------------
# segmentation heads of the MIRAGELight caller
# --------------------------------------------------------------------------------------------
def linear_seg_adapter(tokens: Tensor, sd: Dict[str, Tensor], pre: str, grid: Tuple[int, int],
                       image_hw: Tuple[int, int], start: int = 0, mode: str = "bilinear") -> Tensor:
    """LinearSegAdapter.forward, mirage/output_adapters.py:556-575: task tokens -> 'b (nh nw) d -> b d nh nw' ->
    1x1 conv -> F.interpolate to the image size.  ``tokens`` [B, N_all + global, D]."""
    nh, nw = grid
    B, _, D = tokens.shape
    x = tokens[:, start:start + nh * nw].reshape(B, nh, nw, D).permute(0, 3, 1, 2)           # :565-567
    x = F.conv2d(x, sd[pre + "final_layer.weight"], sd[pre + "final_layer.bias"])            # :569
    return F.interpolate(x, size=image_hw, mode=mode)                                        # :572


def convnext_block(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """ConvNeXtBlock.forward, mirage/output_adapter_utils.py:33-46 (gamma disabled, drop_path 0)."""
    C = x.shape[1]
    y = F.conv2d(x, sd[pre + "dwconv.weight"], sd[pre + "dwconv.bias"], padding=3, groups=C)  # :35
    y = y.permute(0, 2, 3, 1)                                                                # :36
    y = F.layer_norm(y, (C,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-6)          # :37
    y = F.gelu(y @ sd[pre + "pwconv1.weight"].t() + sd[pre + "pwconv1.bias"])                # :38-39
    y = y @ sd[pre + "pwconv2.weight"].t() + sd[pre + "pwconv2.bias"]                        # :40
    return x + y.permute(0, 3, 1, 2)                                                         # :43-45


def convnext_adapter(tokens: Tensor, sd: Dict[str, Tensor], pre: str, grid: Tuple[int, int],
                     image_hw: Tuple[int, int], preds_per_patch: int, depth: int, start: int = 0,
                     mode: str = "bilinear") -> Tensor:
    """ConvNeXtAdapter.forward, mirage/output_adapters.py:493-517."""
    nh, nw = grid
    B, _, D = tokens.shape
    x = tokens[:, start:start + nh * nw] @ sd[pre + "proj_dec.weight"].t() + sd[pre + "proj_dec.bias"]   # :503
    s = int(preds_per_patch ** 0.5)
    C = x.shape[-1] // preds_per_patch
    x = x.reshape(B, nh * nw * preds_per_patch, C)                                           # :504
    x = x.reshape(B, nh, nw, s, s, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, nh * s, nw * s)  # :505-508
    for i in range(depth):
        x = convnext_block(x, sd, f"{pre}blocks.{i}.")                                       # :509
    x = F.conv2d(x, sd[pre + "final_layer.weight"], sd[pre + "final_layer.bias"])            # :510
    return F.interpolate(x, size=image_hw, mode=mode)                                        # :513
