"""Device-side input pipeline for MultiMAE pretraining.

The reference augments every sample on CPU worker processes and ships fp32 tensors to the GPU
(mutils/datasets_pretrain.py:18-83 ``DataAugmentationForMIRAGE``, :172-185 loading).  Here the RAW
uint8 arrays of a batch travel over PCIe (0.79 MB instead of 2.23 MB per sample) and one gather kernel per
modality (csrc/augment.cu) does conversion, flip, intensity shift, the affine warp and the layer-map resize
on the GPU -- so an 8-GPU step is not bound by host cores or by the PCIe switch the GPUs share.

Only the per-sample random PARAMETERS are drawn on the host (a few floats per sample), with the reference's
distributions: flip ~ Bernoulli(hflip) shared by the modalities; per image modality a shift ~ +-|N(0, s)|;
one ``RandomAffine(degrees=10, translate=(0.1, 0.1), scale=(0.9, 1.1), shear=5)`` draw per sample, applied in
full to ``bscan`` / ``bscanlayermap`` and as x-translation + scale only to the other modalities (:52-58).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import ops

LABEL_TASKS = ('layermaps', 'bscanlayermap')
FULL_AFFINE_TASKS = ('bscan', 'bscanlayermap')


def inverse_affine_matrix(angle_deg, tx, ty, scale, shear_x_deg, shear_y_deg=0.0):
    """Output->input affine map in centred pixel coordinates for rotation, translation, isotropic scale and
    shear, in the convention of torchvision's tensor ``F.affine`` (centre at the image centre): vectorised
    over numpy arrays, returns [..., 6] = (m00, m01, m02, m10, m11, m12)."""
    rot = np.radians(np.asarray(angle_deg, dtype=np.float64))
    sx = np.radians(np.asarray(shear_x_deg, dtype=np.float64))
    sy = np.radians(np.asarray(shear_y_deg, dtype=np.float64)) + 0.0 * rot
    tx = np.asarray(tx, dtype=np.float64)
    ty = np.asarray(ty, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    # forward map = R(rot) * Shear(sx, sy) * scale, then translation; its linear part is [[a, b], [c, d]] * scale
    a = np.cos(rot - sy) / np.cos(sy)
    b = -np.cos(rot - sy) * np.tan(sx) / np.cos(sy) - np.sin(rot)
    c = np.sin(rot - sy) / np.cos(sy)
    d = -np.sin(rot - sy) * np.tan(sx) / np.cos(sy) + np.cos(rot)
    # inverse of the linear part (its determinant is 1 before scaling), then undo the translation
    i00, i01, i10, i11 = d / scale, -b / scale, -c / scale, a / scale
    m02 = i00 * (-tx) + i01 * (-ty)
    m12 = i10 * (-tx) + i11 * (-ty)
    return np.stack([i00, i01, m02, i10, i11, m12], axis=-1)


class DeviceAugmentationForMIRAGE:
    """``aug(batch_u8, generator=None) -> {task: network input}`` on the GPU.

    ``batch_u8``: {task: uint8 CUDA tensor [B, H, W]} as stored on disk (images 0..255, layer maps = class
    ids).  Returns fp32 ``[B, 1, H, W]`` in [0, 1] for image tasks and int64 ``[B, h, w]`` (``input_size[task]``,
    nearest resize) for label tasks -- the tensors ``MIRAGEModel.forward`` and the criteria take.
    ``args`` needs ``input_size`` (dict), ``hflip``, ``intensity_shift``, ``affine`` as in the reference's
    argparse namespace.
    """

    def __init__(self, args, degrees=10.0, translate=(0.1, 0.1), scale=(0.9, 1.1), shear=5.0):
        self.args = args
        self.input_size = args.input_size
        self.hflip = float(getattr(args, 'hflip', 0.0))
        self.intensity_shift = float(getattr(args, 'intensity_shift', 0.0))
        self.use_affine = bool(getattr(args, 'affine', False))
        self.degrees, self.translate, self.scale, self.shear = degrees, translate, scale, shear
        self._pinned = {}

    # -- host: parameters -----------------------------------------------------------------------------
    def sample_params(self, tasks, batch: int, size_hw=(512, 512), generator: Optional[torch.Generator] = None):
        """{task: float32 [B, 8]} = (flip, shift, m00, m01, m02, m10, m11, m12) per sample."""
        def uni(lo, hi):
            return (torch.rand(batch, generator=generator, dtype=torch.float64) * (hi - lo) + lo).numpy()
        flip = (torch.rand(batch, generator=generator, dtype=torch.float64).numpy() < self.hflip)
        H, W = size_hw
        if self.use_affine:
            angle = uni(-self.degrees, self.degrees)
            # torchvision's RandomAffine.get_params takes img_size = [width, height] and rounds translations
            tx = np.round(uni(-self.translate[0] * W, self.translate[0] * W))
            ty = np.round(uni(-self.translate[1] * H, self.translate[1] * H))
            sc = uni(self.scale[0], self.scale[1])
            shx = uni(-self.shear, self.shear)
        out = {}
        for task in tasks:
            p = np.zeros((batch, 8), dtype=np.float32)
            p[:, 0] = flip
            if self.intensity_shift > 0 and task not in LABEL_TASKS:
                mag = torch.randn(batch, generator=generator, dtype=torch.float64).numpy() * self.intensity_shift
                sign = np.where(torch.rand(batch, generator=generator, dtype=torch.float64).numpy() < 0.5, -1.0, 1.0)
                p[:, 1] = (mag * sign).astype(np.float32)
            if self.use_affine:
                if task in FULL_AFFINE_TASKS:
                    m = inverse_affine_matrix(angle, tx, ty, sc, shx)
                else:
                    m = inverse_affine_matrix(0.0 * angle, tx, 0.0 * ty, sc, 0.0 * shx)
            else:
                m = np.tile(np.array([1, 0, 0, 0, 1, 0], dtype=np.float64), (batch, 1))
            p[:, 2:] = m.astype(np.float32)
            out[task] = torch.from_numpy(p)
        return out

    # -- device ---------------------------------------------------------------------------------------
    def apply(self, batch_u8: Dict[str, torch.Tensor], params: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        out = {}
        for task, src in batch_u8.items():
            if src.dim() == 4:
                src = src[:, 0]
            src = src.contiguous()
            p = params[task]
            if not p.is_cuda:
                slot = self._pinned.get(task)
                if slot is None or slot.shape != p.shape:
                    slot = self._pinned[task] = torch.empty(p.shape, dtype=p.dtype).pin_memory()
                slot.copy_(p)
                p = slot.to(src.device, non_blocking=True)
            if task in LABEL_TASKS:
                oh, ow = self.input_size[task]
                out[task] = ops.augment_labels(src, p, (int(oh), int(ow)))
            else:
                oh, ow = self.input_size[task]
                if (int(oh), int(ow)) != tuple(src.shape[1:]):
                    raise NotImplementedError('image modalities are augmented at their stored size '
                                              f'({tuple(src.shape[1:])}), got input_size {self.input_size[task]}')
                out[task] = ops.augment_image(src, p)
        return out

    def __call__(self, batch_u8: Dict[str, torch.Tensor], generator: Optional[torch.Generator] = None):
        first = next(iter(batch_u8.values()))
        B = first.shape[0]
        hw = tuple(first.shape[-2:])
        return self.apply(batch_u8, self.sample_params(list(batch_u8.keys()), B, hw, generator))


def stage_uint8_batch(host_batch: Dict[str, torch.Tensor], device, stream: Optional[torch.cuda.Stream] = None):
    """Pinned uint8 host tensors -> device, asynchronously on ``stream`` (default: current)."""
    ctx = torch.cuda.stream(stream) if stream is not None else _Null()
    with ctx:
        return {k: v.to(device, non_blocking=True) for k, v in host_batch.items()}


class _Null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
