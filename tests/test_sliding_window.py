"""Sliding-window host logic (CPU): window enumeration, gaussian map and blending against vendored-MONAI golden outputs,
mirror TTA against the whole-volume-flip definition, and window sharding over a world_size-2 gloo group."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_inputs as gi
from segmamba_b200 import sliding_window as sw


def _predictor():
    w = torch.from_numpy(np.random.RandomState(51).standard_normal((3, 2, 3, 3, 3)).astype(np.float32))
    return lambda t: torch.nn.functional.conv3d(t, w, padding=1)


def test_window_starts_and_gaussian_match_monai():
    gold = gi.load("sliding_window")
    assert np.array_equal(np.array(sw.window_starts((155, 240, 240), (128, 128, 128), 0.5)), gold["starts_brats"])
    assert np.array_equal(np.array(sw.window_starts((40, 50, 33), (32, 32, 32), 0.5)), gold["starts_small"])
    assert len(sw.window_starts((155, 240, 240), (128, 128, 128), 0.5)) == 18
    g = sw.gaussian_importance_map((32, 32, 32))
    assert np.allclose(g.numpy(), gold["gauss32"], rtol=1e-6, atol=1e-7)


def test_blend_matches_monai_golden():
    gold = gi.load("sliding_window")
    x = gi.model_input(50, (1, 2, 40, 50, 33))
    out = sw.sliding_window_inference(x, (32, 32, 32), 2, _predictor(), overlap=0.5, mode="gaussian")
    assert np.allclose(out.numpy(), gold["blend_small"], rtol=1e-4, atol=1e-5)
    inf = sw.SlidingWindowInferer(roi_size=(32, 32, 32), sw_batch_size=2, overlap=0.5, mode="gaussian")
    assert torch.equal(inf(x, _predictor()), out)


def test_small_image_is_padded_like_monai():
    x = gi.model_input(52, (1, 2, 20, 40, 33))           # first axis smaller than the window
    out = sw.sliding_window_inference(x, (32, 32, 32), 1, _predictor())
    assert out.shape == (1, 3, 20, 40, 33)


def test_mirror_tta_is_whole_volume_flip():
    """prediction.py:128-155: flip the volume, window it, flip the result back, average."""
    x = gi.model_input(53, (1, 2, 40, 50, 33))
    pred = _predictor()
    base = lambda t: sw.sliding_window_inference(t, (32, 32, 32), 2, pred)
    ref = base(x)
    combos = [(2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4)]
    for dims in combos:
        ref = ref + torch.flip(base(torch.flip(x, dims)), dims)
    ref = ref / 8
    got = sw.sliding_window_inference(x, (32, 32, 32), 2, pred, mirror_axes=(0, 1, 2))
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_mirror_tta_batches_spanning_several_flips():
    """sw_batch_size larger than the work items of one flip (image <= roi: one window per flip): a predictor batch would
    span several flips; every run of equal flip must go to its own accumulator.  8 flips x 1 window, batch sizes 1..5."""
    x = gi.model_input(54, (1, 2, 8, 8, 8))
    pred = _predictor()
    ref = sw.sliding_window_inference(x, (8, 8, 8), 1, pred, mirror_axes=(0, 1, 2))
    manual = sum(torch.flip(pred(torch.flip(x, d) if d else x), d) if d else pred(x)
                 for d in [(), (2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4)]) / 8
    assert torch.allclose(ref, manual, rtol=1e-5, atol=1e-6)
    for bs in (2, 3, 4, 5):
        got = sw.sliding_window_inference(x, (8, 8, 8), bs, pred, mirror_axes=(0, 1, 2))
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6), f"sw_batch_size={bs}"
    # two windows per flip, batch of 3: runs of (2, 1), (1, 2), ...
    x2 = gi.model_input(55, (1, 2, 12, 8, 8))
    ref2 = sw.sliding_window_inference(x2, (8, 8, 8), 1, pred, mirror_axes=(0, 1, 2))
    for bs in (3, 5):
        assert torch.allclose(sw.sliding_window_inference(x2, (8, 8, 8), bs, pred, mirror_axes=(0, 1, 2)), ref2,
                              rtol=1e-5, atol=1e-6)


def _worker_thin(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = gi.model_input(54, (1, 2, 8, 8, 8))
        out = sw.sliding_window_inference(x, (8, 8, 8), 4, _predictor(), group=True, assemble_on=None, mirror_axes=(0, 1, 2))
        ret[f"thin{rank}"] = out.numpy()
    finally:
        dist.destroy_process_group()


def test_thin_shards_spanning_flips_two_ranks_gloo():
    """world 2, one window per flip, sw_batch_size 4: every rank's batches span four flips."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_thin, args=(2, port, ret), nprocs=2, join=True)
    x = gi.model_input(54, (1, 2, 8, 8, 8))
    single = sw.sliding_window_inference(x, (8, 8, 8), 1, _predictor(), mirror_axes=(0, 1, 2))
    assert np.allclose(ret["thin0"], single.numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(ret["thin1"], single.numpy(), rtol=1e-5, atol=1e-6)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = gi.model_input(50, (1, 2, 40, 50, 33))
        out = sw.sliding_window_inference(x, (32, 32, 32), 2, _predictor(), group=True, assemble_on=0, mirror_axes=(1,))
        allr = sw.sliding_window_inference(x, (32, 32, 32), 2, _predictor(), group=True, assemble_on=None)
        if rank == 0:
            ret["out"] = out.numpy()
        else:
            assert out is None
        ret[f"all{rank}"] = allr.numpy()
    finally:
        dist.destroy_process_group()


def test_sharded_over_two_ranks_gloo():
    """world_size 2 on CPU (gloo): windows[rank::2] per rank, one reduce at the end, same result as a single process."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    x = gi.model_input(50, (1, 2, 40, 50, 33))
    single = sw.sliding_window_inference(x, (32, 32, 32), 2, _predictor(), mirror_axes=(1,))
    assert np.allclose(ret["out"], single.numpy(), rtol=1e-5, atol=1e-6)
    gold = gi.load("sliding_window")
    assert np.allclose(ret["all0"], gold["blend_small"], rtol=1e-4, atol=1e-5)
    assert np.allclose(ret["all1"], gold["blend_small"], rtol=1e-4, atol=1e-5)
