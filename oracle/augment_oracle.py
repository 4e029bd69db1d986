"""CPU restatement of the reference's pretraining augmentation for ONE sample (test infrastructure only).

Follows mutils/datasets_pretrain.py:36-83 (``DataAugmentationForMIRAGE.__call__``) and :172-185 (loading) op by
op with the same torchvision calls, but takes the random draws as ARGUMENTS (flip, shift per task, affine
parameters) so the CUDA path can be checked on identical parameters.  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torchvision.transforms.functional as TF
from torchvision import transforms

LABEL_TASKS = ('layermaps', 'bscanlayermap')


def load_like_reference(raw_u8: np.ndarray, task: str) -> np.ndarray:
    """datasets_pretrain.py:175-180: layer maps -> int, images -> float32 / 255."""
    if task in LABEL_TASKS:
        return raw_u8.astype(int)
    return raw_u8.astype(np.float32) / 255.0


def augment_sample(task_dict: dict, flip: bool, shifts: dict, affine_params, input_size: dict, use_affine=True):
    """task_dict: {task: array [H, W]} as load_like_reference returns them; ``affine_params`` =
    (angle, (tx, ty), scale, (shear_x, shear_y)) as RandomAffine.get_params returns them (:39-42)."""
    out = {}
    for task, arr in task_dict.items():
        if flip:
            arr = np.flip(arr, axis=-1)                                             # :43-44
        if task not in LABEL_TASKS and shifts.get(task) is not None:
            shift = np.asarray([shifts[task]], dtype=np.float32)
            arr = np.clip(arr + shift, 0, 1)                                        # :45-50
        img = torch.from_numpy(arr.copy()).contiguous().unsqueeze(0)                # :52
        if task in ('bscan', 'bscanlayermap'):
            c_params = affine_params                                                # :54-56
        else:
            c_params = 0, (affine_params[1][0], 0), affine_params[2], 0             # :57-58
        if use_affine:
            img = TF.affine(img, *c_params, interpolation=transforms.InterpolationMode.BILINEAR, fill=0,
                            center=None)                                            # :59-67
        interp = TF.InterpolationMode.NEAREST if task in LABEL_TASKS else TF.InterpolationMode.BILINEAR
        if img.shape[1:] != tuple(input_size[task]):
            img = TF.resize(img, list(input_size[task]), interpolation=interp)      # :72-78
        if task in LABEL_TASKS:
            img = img.squeeze(0)
        out[task] = img
    return out
