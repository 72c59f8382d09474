"""Host-side training loop: PolyLRScheduler and the per-iteration sequence of the reference trainer
against vectors produced by the reference's own scheduler class (oracle/gen_golden.py: gen_train), and checkpoint/resume."""
import os

import numpy as np
import torch

from golden_inputs import train_toy_model
from segmamba_b200.train_step import PolyLRScheduler, TrainStep, load_checkpoint, reference_optimizer, save_checkpoint

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_loop.npz"))


def test_poly_lr_matches_reference_class():
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD(p, lr=1e-2)
    sch = PolyLRScheduler(opt, initial_lr=1e-2, max_steps=10)
    lrs = [opt.param_groups[0]["lr"]]
    for _ in range(9):
        opt.step()
        sch.step()
        lrs.append(opt.param_groups[0]["lr"])
    np.testing.assert_allclose(np.array(lrs), GOLD["poly_lrs"], rtol=1e-12, atol=0)
    sch.step(current_step=3)                                     # explicit position, as the reference's signature allows
    assert abs(opt.param_groups[0]["lr"] - 1e-2 * (1 - 3 / 10) ** 0.9) < 1e-15


def _make():
    model = train_toy_model()
    opt = reference_optimizer(model)
    sch = PolyLRScheduler(opt, initial_lr=1e-2, max_steps=20)
    return model, TrainStep(model, opt, torch.nn.CrossEntropyLoss(), autocast_dtype=None, scheduler=sch)


def test_train_loop_matches_reference_sequence():
    """bit-exact on CPU: same torch ops in the same order as trainer.py:444-477."""
    model, step = _make()
    xs, ys = torch.from_numpy(GOLD["train_x"]), torch.from_numpy(GOLD["train_y"])
    losses = [float(step(xs[i], ys[i])) for i in range(6)]
    np.testing.assert_array_equal(np.array(losses), GOLD["train_losses"])
    assert step.optimizer.param_groups[0]["lr"] == float(GOLD["train_final_lr"])
    for k, v in model.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), GOLD["train_param_" + k])
    assert step.global_step == 6


def test_checkpoint_resume_is_exact(tmp_path):
    xs, ys = torch.from_numpy(GOLD["train_x"]), torch.from_numpy(GOLD["train_y"])
    _, step = _make()
    for i in range(3):
        step(xs[i], ys[i])
    path = str(tmp_path / "ckpt" / "state.pt")
    save_checkpoint(path, step)
    model2, step2 = _make()                                      # fresh objects, as after a restart
    assert load_checkpoint(path, step2) == 3
    for i in range(3, 6):
        step2(xs[i], ys[i])
    for k, v in model2.state_dict().items():                     # identical to the uninterrupted run
        np.testing.assert_array_equal(v.numpy(), GOLD["train_param_" + k])
    assert step2.optimizer.param_groups[0]["lr"] == float(GOLD["train_final_lr"])
    # weights-only consumers (4_predict.py:52-53) read the "model" entry
    model3 = train_toy_model()
    model3.load_state_dict(torch.load(path, weights_only=True)["model"], strict=True)


def test_fp16_path_builds_a_scaler_and_bad_checkpoints_raise(tmp_path):
    model = train_toy_model()
    step = TrainStep(model, reference_optimizer(model), torch.nn.CrossEntropyLoss(), autocast_dtype=torch.float16)
    assert step.grad_scaler is not None                          # trainer.py:67
    try:
        step.load_state_dict({"format": "something else"})
    except ValueError:
        pass
    else:
        raise AssertionError("foreign checkpoint accepted")


def _ddp_worker(rank, world, port, ret, ckpt):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = train_toy_model()
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        opt = reference_optimizer(model)
        step = TrainStep(ddp, opt, torch.nn.CrossEntropyLoss(), autocast_dtype=None,
                         scheduler=PolyLRScheduler(opt, initial_lr=1e-2, max_steps=20))
        xs, ys = torch.from_numpy(GOLD["train_x"]), torch.from_numpy(GOLD["train_y"])
        for i in range(3):                                       # each rank gets half of the batch (weak scaling unit = 1 patch)
            step(xs[i, rank:rank + 1], ys[i, rank:rank + 1])
        ret[rank] = {k: v.numpy().copy() for k, v in model.state_dict().items()}
        if rank == 0:
            save_checkpoint(ckpt, step)                          # unwraps .module: keys carry no "module." prefix
    finally:
        dist.destroy_process_group()


def test_ddp_two_ranks_gloo_matches_single_process(tmp_path):
    """world_size 2 on CPU (gloo): per-rank half batches + DDP's gradient averaging == the single-process full batch
    (CE loss is a mean, so the averaged half-batch gradients are the full-batch gradient up to summation order)."""
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    ckpt = str(tmp_path / "ddp.pt")
    mp.spawn(_ddp_worker, args=(2, 29500 + (os.getpid() % 2000) + 7, ret, ckpt), nprocs=2, join=True)
    model, step = _make()
    xs, ys = torch.from_numpy(GOLD["train_x"]), torch.from_numpy(GOLD["train_y"])
    for i in range(3):
        step(xs[i], ys[i])
    for k, v in model.state_dict().items():
        np.testing.assert_allclose(ret[0][k], v.numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_array_equal(ret[0][k], ret[1][k])      # replicas stay identical
    saved = torch.load(ckpt, weights_only=True)
    assert set(saved["model"].keys()) == set(model.state_dict().keys())


def _toy_with_norm():
    torch.manual_seed(11)
    return torch.nn.Sequential(torch.nn.Conv3d(2, 8, 3, padding=1), torch.nn.GroupNorm(2, 8), torch.nn.ReLU(),
                               torch.nn.ConvTranspose3d(8, 8, 2, 2), torch.nn.Conv3d(8, 3, 1))


def test_master_weights_step_equals_autocast_step(tmp_path):
    """bf16 parameters + fp32 masters (MasterWeights) against the plain autocast step on CPU: the low-precision parameters hold
    exactly the values autocast's casts produce, so losses and fp32 weights agree bit for bit over several steps; parameters
    outside the matmul modules (here the GroupNorm affine) stay fp32 and are their own masters; checkpoints carry fp32."""
    from segmamba_b200.master_weights import MasterWeights
    g = torch.Generator().manual_seed(3)
    xs = torch.rand(5, 2, 2, 6, 6, 6, generator=g)
    ys = torch.randint(0, 3, (5, 2, 12, 12, 12), generator=g)
    ref_model = _toy_with_norm()
    ref = TrainStep(ref_model, reference_optimizer(ref_model), torch.nn.CrossEntropyLoss(), autocast_dtype=torch.bfloat16)
    model = _toy_with_norm()
    mw = MasterWeights(model)
    assert set(mw.converted_names()) == {"0.weight", "0.bias", "3.weight", "3.bias", "4.weight", "4.bias"}
    assert model[0].weight.dtype == torch.bfloat16 and model[1].weight.dtype == torch.float32
    opt = torch.optim.SGD(mw.optimizer_parameters(), lr=1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    step = TrainStep(model, opt, torch.nn.CrossEntropyLoss(), autocast_dtype=torch.bfloat16, master_weights=mw)
    for i in range(4):
        la, lb = float(ref(xs[i], ys[i])), float(step(xs[i], ys[i]))
        assert la == lb, (i, la, lb)
    sd = mw.state_dict()
    for k, v in ref_model.state_dict().items():
        assert sd[k].dtype == v.dtype
        np.testing.assert_array_equal(sd[k].numpy(), v.numpy())
        if k in mw.converted_names():                            # the model holds the rounded masters
            np.testing.assert_array_equal(model.state_dict()[k].float().numpy(), v.to(torch.bfloat16).float().numpy())
    # resume: a fresh pair restored from the checkpoint continues identically
    path = str(tmp_path / "mw.pt")
    save_checkpoint(path, step)
    model2 = _toy_with_norm()
    mw2 = MasterWeights(model2)
    opt2 = torch.optim.SGD(mw2.optimizer_parameters(), lr=1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    step2 = TrainStep(model2, opt2, torch.nn.CrossEntropyLoss(), autocast_dtype=torch.bfloat16, master_weights=mw2)
    assert load_checkpoint(path, step2) == 4
    assert float(ref(xs[4], ys[4])) == float(step2(xs[4], ys[4]))
    for k, v in ref_model.state_dict().items():
        np.testing.assert_array_equal(mw2.state_dict()[k].numpy(), v.numpy())
