"""Checkpoint I/O with the reference's on-disk schema (mutils/checkpoint.py:9-72), so that files written by
either side load in the other:

    {'model': model.state_dict(), 'optimizer': optimizer.state_dict(), 'epoch': int,
     'scaler': loss_scaler.state_dict(), 'args': argparse.Namespace}          -> checkpoint-<epoch>.pth

``state_dict`` keys and shapes of every module in this package are the reference's, the
optimizer state is torch.optim.AdamW's (optim.FusedAdamW reads and writes that layout), and the pickled ``args``
Namespace is the model configuration ``mirage_wrapper.MIRAGEWrapper`` rebuilds the model from.  Host-side only.
"""
from __future__ import annotations

import glob
import os
from pathlib import Path

import torch


def save_model(args, epoch, model, optimizer, loss_scaler, loss_balancer=None):
    """mutils/checkpoint.py:9-25 (the DeepSpeed branch of the reference is dead code and not mirrored)."""
    output_dir = Path(args.output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    path = output_dir / ('checkpoint-%s.pth' % str(epoch))
    to_save = {
        'model': model.state_dict(),
        'optimizer': optimizer.state_dict(),
        'epoch': epoch,
        'scaler': loss_scaler.state_dict() if loss_scaler is not None else {},
        'args': args,
    }
    if loss_balancer is not None:
        to_save['loss_balancer'] = loss_balancer.state_dict()
    torch.save(to_save, path)
    return path


def latest_checkpoint(output_dir) -> str:
    """Newest numeric ``checkpoint-N.pth`` in ``output_dir`` ('' when there is none); :46-54."""
    latest = -1
    for ckpt in glob.glob(os.path.join(str(output_dir), 'checkpoint-*.pth')):
        t = ckpt.split('-')[-1].split('.')[0]
        if t.isdigit():
            latest = max(int(t), latest)
    return os.path.join(str(output_dir), 'checkpoint-%d.pth' % latest) if latest >= 0 else ''


def auto_load_model(args, model, optimizer, loss_scaler, best=False):
    """Resume from ``args.resume`` or, with ``args.auto_resume``, from the newest checkpoint (:35-72).
    Restores model, optimizer, scaler and sets ``args.start_epoch``."""
    output_dir = Path(args.output_dir)
    if getattr(args, 'auto_resume', False) and len(getattr(args, 'resume', '') or '') == 0:
        if best:
            args.resume = os.path.join(output_dir, 'checkpoint-best.pth')
            assert os.path.exists(args.resume), f"Best checkpoint not found at {args.resume}"
        else:
            args.resume = latest_checkpoint(output_dir)
    if not getattr(args, 'resume', ''):
        return None
    if args.resume.startswith('https'):
        checkpoint = torch.hub.load_state_dict_from_url(args.resume, map_location='cpu')
    else:
        checkpoint = torch.load(args.resume, map_location='cpu', weights_only=False)
    model.load_state_dict(checkpoint['model'])
    if not best and 'optimizer' in checkpoint and 'epoch' in checkpoint:
        optimizer.load_state_dict(checkpoint['optimizer'])
        args.start_epoch = checkpoint['epoch'] + 1
        if 'scaler' in checkpoint and loss_scaler is not None:
            loss_scaler.load_state_dict(checkpoint['scaler'])
    return checkpoint
