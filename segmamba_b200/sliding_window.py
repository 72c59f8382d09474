"""GPU-resident sliding-window inference with gaussian blending, optional mirror TTA and window sharding over ranks.

Restates the non-buffered path of MONAI's ``sliding_window_inference`` (monai/inferers/utils.py:138-321) with the
reference's settings (``SlidingWindowInferer(roi_size=[128]*3, sw_batch_size=2, overlap=0.5, mode="gaussian")``,
4_predict.py:55-59): window enumeration ``dense_patch_slices`` (monai/data/utils.py:171-211) with
``scan_interval = int(roi * (1 - overlap))`` (inferers/utils.py:363-384), gaussian importance map with
sigma = 0.125 * roi clamped from below (monai/data/utils.py:1088-1137), weighted accumulate, divide by the weight map.
The 8-pass mirror test-time augmentation of ``Predictor.maybe_mirror_and_predict`` (light_training/prediction.py:110-159)
is available as ``mirror_axes``.

Differences that stay behind the interface: everything (input, windows, accumulators) stays on the GPU -- the reference
moves every TTA result to the CPU and rebuilds the model per case (4_predict.py:44-54,73) -- and with a process group the
windows are sharded ``windows[rank::world]`` with no collective on the data path; the only exchange is the final
assembly (one reduce of the weighted accumulator to rank 0).
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Sequence

import torch
import torch.nn.functional as F


def scan_interval(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> tuple:
    out = []
    for img, roi in zip(image_size, roi_size):
        if roi == img:
            out.append(int(roi))
        else:
            iv = int(roi * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], overlap: float = 0.5) -> list:
    """start index of every window, in MONAI's order (first spatial axis slowest)."""
    interval = scan_interval(image_size, roi_size, overlap)
    per_dim = []
    for img, roi, iv in zip(image_size, roi_size, interval):
        num = 1 if img <= roi else int(math.ceil(float(img - roi) / iv)) + 1
        per_dim.append([min(i * iv, img - roi) for i in range(num)])
    return list(itertools.product(*per_dim))


def gaussian_importance_map(roi_size: Sequence[int], sigma_scale: float = 0.125, device=None, dtype=torch.float32):
    m = None
    for i, roi in enumerate(roi_size):
        sigma = sigma_scale * roi
        x = torch.arange(start=-(roi - 1) / 2.0, end=(roi - 1) / 2.0 + 1, dtype=torch.float32, device=device)
        gi = torch.exp(x ** 2 / (-2 * sigma ** 2))
        m = gi if m is None else m.unsqueeze(-1) * gi[(None,) * i]
    min_non_zero = max(float(m.min()), 1e-3)
    return torch.clamp_(m, min=min_non_zero).to(dtype)


def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int, predictor: Callable,
                             overlap: float = 0.5, mode: str = "gaussian", sigma_scale: float = 0.125,
                             mirror_axes: Sequence[int] | None = None, group=None, assemble_on: int | None = 0):
    """inputs: (B, C, D, H, W) -> (B, C_out, D, H, W), same device.

    ``mirror_axes``: spatial axes (0, 1, 2) to flip for test-time augmentation; all 2^k combinations are averaged, as
    prediction.py:125-155 does.  ``group``: torch.distributed process group (or True for the default group) to shard the
    windows over; rank ``assemble_on`` gets the result (None: every rank, via all_reduce), other ranks return None.
    """
    if mode not in ("gaussian", "constant"):
        raise ValueError("mode must be 'gaussian' or 'constant'")
    if not 0 <= overlap < 1:
        raise ValueError(f"overlap must be >= 0 and < 1, got {overlap}.")
    batch = inputs.shape[0]
    image_size_ = tuple(inputs.shape[2:])
    roi_size = tuple(int(r) for r in roi_size)
    # pad when the image is smaller than the window (inferers/utils.py:163-171)
    image_size = tuple(max(i, r) for i, r in zip(image_size_, roi_size))
    pad_size = []
    for k in range(len(inputs.shape) - 1, 1, -1):
        diff = max(roi_size[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    if any(pad_size):
        inputs = F.pad(inputs, pad=pad_size, mode="constant", value=0.0)

    starts = window_starts(image_size, roi_size, overlap)
    num_win = len(starts)
    dev, dtype = inputs.device, inputs.dtype
    w = (gaussian_importance_map(roi_size, sigma_scale, dev, dtype) if mode == "gaussian"
         else torch.ones(roi_size, device=dev, dtype=dtype))[None, None]

    rank, world = 0, 1
    dist = None
    if group is not None:
        import torch.distributed as dist_mod
        dist = dist_mod
        grp = None if group is True else group
        rank, world = dist.get_rank(grp), dist.get_world_size(grp)

    flips = [()]
    if mirror_axes:
        flips = [c for k in range(len(mirror_axes) + 1) for c in itertools.combinations(mirror_axes, k)]

    # work list = (flip, image index, window index), sharded round-robin over ranks.  Like the reference, a mirrored
    # pass flips the WHOLE volume, windows it at the same start indices, and flips the blended result back
    # (prediction.py:128-155) -- not the same as flipping each window, because the starts are not mirror-symmetric.
    work = [(fi, b, wi) for fi in range(len(flips)) for b in range(batch) for wi in range(num_win)]
    mine = work[rank::world]

    count = torch.zeros((1, 1) + image_size, device=dev, dtype=dtype)
    for st in starts:                                   # the weight map does not depend on the predictions
        count[(slice(None), slice(None)) + tuple(slice(s, s + r) for s, r in zip(st, roi_size))] += w
    out = None                                          # sum over flips of flip_back(acc_f / count), this rank's share
    acc, acc_flip, xin, xin_flip = None, None, None, None

    def flush():
        nonlocal out, acc
        if acc is None:
            return
        res = acc / count
        fl = flips[acc_flip]
        if fl:
            res = torch.flip(res, dims=[2 + a for a in fl])
        out = res if out is None else out + res
        acc = None

    for i0 in range(0, len(mine), sw_batch_size):
        # a predictor batch never mixes flips (one accumulator is live at a time): a batch that spans several flips --
        # few windows per flip, e.g. image <= roi with sw_batch_size >= 3, or a thin shard -- is split into one run per flip
        for fi, run in itertools.groupby(mine[i0:i0 + sw_batch_size], key=lambda c: c[0]):
            part = list(run)
            if xin_flip != fi:
                flush()
                fl = flips[fi]
                xin = torch.flip(inputs, dims=[2 + a for a in fl]) if fl else inputs
                xin_flip = fi
            wins = [xin[(slice(b, b + 1), slice(None)) + tuple(slice(s, s + r) for s, r in zip(starts[wi], roi_size))]
                    for _, b, wi in part]
            pred = predictor(torch.cat(wins) if len(wins) > 1 else wins[0])
            if acc is None:
                acc = torch.zeros((batch, pred.shape[1]) + image_size, device=dev, dtype=dtype)
                acc_flip = fi
            for k, (_, b, wi) in enumerate(part):
                sl = tuple(slice(s, s + r) for s, r in zip(starts[wi], roi_size))
                acc[(slice(b, b + 1), slice(None)) + sl] += pred[k:k + 1].to(dtype) * w
    flush()
    if out is None:                                      # a rank without work still takes part in the assembly
        probe = predictor(inputs[(slice(0, 1), slice(None)) + tuple(slice(0, r) for r in roi_size)])
        out = torch.zeros((batch, probe.shape[1]) + image_size, device=dev, dtype=dtype)

    if world > 1:
        grp = None if group is True else group
        if assemble_on is None:
            dist.all_reduce(out, group=grp)
        else:
            dist.reduce(out, dst=assemble_on, group=grp)
            if rank != assemble_on:
                return None
    out = out / len(flips)
    if any(pad_size):
        final = []
        nd = len(image_size)
        for sp in range(nd):
            si = nd - sp - 1
            final.insert(0, slice(pad_size[sp * 2], pad_size[sp * 2] + image_size_[si]))
        out = out[(slice(None), slice(None)) + tuple(final)]
    return out


class SlidingWindowInferer:
    """call-compatible subset of monai.inferers.SlidingWindowInferer (inferer.py:382-535)."""

    def __init__(self, roi_size, sw_batch_size=1, overlap=0.25, mode="constant", sigma_scale=0.125, mirror_axes=None,
                 group=None, assemble_on=0):
        self.roi_size, self.sw_batch_size, self.overlap, self.mode = roi_size, sw_batch_size, overlap, mode
        self.sigma_scale, self.mirror_axes, self.group, self.assemble_on = sigma_scale, mirror_axes, group, assemble_on

    def __call__(self, inputs, network, *args, **kwargs):
        fn = (lambda x: network(x, *args, **kwargs)) if (args or kwargs) else network
        return sliding_window_inference(inputs, self.roi_size, self.sw_batch_size, fn, self.overlap, self.mode, self.sigma_scale,
                                        self.mirror_axes, self.group, self.assemble_on)
