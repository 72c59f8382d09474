"""Training-step harness around the native SegMamba: the inner loop of the reference trainer, plus full checkpoint/resume.

Mirrors, step for step, what ``Trainer.train_epoch`` does per iteration (light_training/trainer.py:445-478):

    grads = None -> autocast forward + loss -> (scaled) backward -> unscale -> clip_grad_norm_(12) -> optimizer.step
    -> scaler.update -> scheduler.step

with the reference's optimiser and schedule (``SGD(lr=1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)``,
3_train.py:51-52; ``PolyLRScheduler``, light_training/utils/lr_scheduler.py:22-38, built at trainer.py:399-403).  The
reference trains under fp16 autocast with a ``GradScaler`` (trainer.py:67,450); bf16 (BASELINE.json configs[2], no
scaler) is the default here, fp16 + scaler is selectable.

What the reference does not have: it saves model weights only (``save_new_model_and_delete_last``,
light_training/utils/files_helper.py:13-22), so a run cannot be resumed; ``save_checkpoint`` / ``load_checkpoint`` carry
model, optimiser (momentum buffers), schedule position, scaler state and the step counter.

Device-agnostic host logic (the model decides where the compute runs), so it is unit-tested on CPU with a stand-in
module; ``bench.py`` times the same sequence of operations.
"""
from __future__ import annotations

import os
from typing import Callable

import torch


class PolyLRScheduler:
    """lr(step) = initial_lr * (1 - step / max_steps) ** exponent, written into every param group.

    Same constructor arguments and ``step(current_step=None)`` behaviour as the reference class
    (light_training/utils/lr_scheduler.py:22-38).  The reference derives from ``_LRScheduler``, whose constructor issues one
    ``step()``: construction sets lr(0) and leaves the counter at 1, so the n-th training iteration's ``step()`` sets lr(n).
    That initial step is reproduced here (tests/golden/train_loop.npz holds the reference class's trajectory).
    ``current_step`` is accepted for signature compatibility; as in the reference it does not move the counter (resume goes
    through ``state_dict`` / ``load_state_dict``).
    """

    def __init__(self, optimizer: torch.optim.Optimizer, initial_lr: float, max_steps: int, exponent: float = 0.9,
                 current_step: int | None = None):
        self.optimizer, self.initial_lr, self.max_steps, self.exponent = optimizer, initial_lr, max_steps, exponent
        self.ctr = 0
        self.step()                                              # the base-class constructor's initial step

    def lr_at(self, step: int) -> float:
        return self.initial_lr * (1 - step / self.max_steps) ** self.exponent

    def step(self, current_step: int | None = None) -> None:
        if current_step is None or current_step == -1:
            current_step = self.ctr
            self.ctr += 1
        new_lr = self.lr_at(current_step)
        for group in self.optimizer.param_groups:
            group["lr"] = new_lr

    def state_dict(self) -> dict:
        return {"ctr": self.ctr, "initial_lr": self.initial_lr, "max_steps": self.max_steps, "exponent": self.exponent}

    def load_state_dict(self, state: dict) -> None:
        self.ctr = int(state["ctr"])
        self.initial_lr, self.max_steps, self.exponent = state["initial_lr"], state["max_steps"], state["exponent"]


def reference_optimizer(model: torch.nn.Module, lr: float = 1e-2) -> torch.optim.SGD:
    """the optimiser of 3_train.py:51-52."""
    return torch.optim.SGD(model.parameters(), lr=lr, weight_decay=3e-5, momentum=0.99, nesterov=True)


class TrainStep:
    """``loss = TrainStep(model, optimizer, loss_fn, ...)(image, label)``: one iteration of trainer.py:445-478.

    model          nn.Module or DistributedDataParallel wrapper (gradient all-reduce happens inside backward).
    loss_fn        ``loss_fn(logits, label) -> scalar``; the reference uses ``nn.CrossEntropyLoss()`` (3_train.py:54,60).
    autocast_dtype torch.bfloat16 (default), torch.float16 (then a GradScaler is created unless one is passed), or None.
    clip_grad_norm 12, as trainer.py:464,469; None disables.
    scheduler      anything with ``step()``; called after the optimiser step, as trainer.py:476-477.
    master_weights optional ``MasterWeights(model)`` (segmamba_b200/master_weights.py): the model's matmul / convolution
                   parameters live in the autocast dtype, ``optimizer`` must have been built on
                   ``master_weights.optimizer_parameters()``; same arithmetic, two multi-tensor copies per step instead of one
                   cast kernel per weight and per gradient.
    """

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable,
                 autocast_dtype: torch.dtype | None = torch.bfloat16, grad_scaler=None, clip_grad_norm: float | None = 12.0,
                 scheduler=None, master_weights=None):
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.autocast_dtype, self.clip, self.scheduler = autocast_dtype, clip_grad_norm, scheduler
        if grad_scaler is None and autocast_dtype is torch.float16:
            grad_scaler = torch.amp.GradScaler()
        self.grad_scaler = grad_scaler
        self.global_step = 0
        self.master_weights = master_weights
        self._model_params = [p for p in model.parameters() if p.requires_grad]
        self._params = master_weights.optimizer_parameters() if master_weights is not None else self._model_params

    def __call__(self, image: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        self.global_step += 1
        self.model.train()
        for p in self._params:                                   # trainer.py:444 (grad = None, not zeros)
            p.grad = None
        for p in self._model_params:
            p.grad = None
        with torch.autocast(image.device.type, dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            logits = self.model(image)
            loss = self.loss_fn(logits.float(), label)
        mw = self.master_weights
        if self.grad_scaler is not None:
            self.grad_scaler.scale(loss).backward()
            if mw is not None:
                mw.grads_to_master()
            self.grad_scaler.unscale_(self.optimizer)
            if self.clip is not None:
                torch.nn.utils.clip_grad_norm_(self._params, self.clip)
            self.grad_scaler.step(self.optimizer)
            self.grad_scaler.update()
        else:
            loss.backward()
            if mw is not None:
                mw.grads_to_master()
            if self.clip is not None:
                torch.nn.utils.clip_grad_norm_(self._params, self.clip)
            self.optimizer.step()
        if mw is not None:
            mw.master_to_model()
        if self.scheduler is not None:
            self.scheduler.step()
        return loss.detach()

    # ---- checkpoint / resume -------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        module = self.model.module if hasattr(self.model, "module") else self.model
        return {
            "format": "segmamba_b200.train_step/1",
            # the reference's 291 keys in fp32: loadable by 4_predict.py:52-53
            "model": self.master_weights.state_dict() if self.master_weights is not None else module.state_dict(),
            "optimizer": self.optimizer.state_dict(),
            "scheduler": self.scheduler.state_dict() if self.scheduler is not None and hasattr(self.scheduler, "state_dict") else None,
            "grad_scaler": self.grad_scaler.state_dict() if self.grad_scaler is not None else None,
            "global_step": self.global_step,
        }

    def load_state_dict(self, state: dict) -> None:
        if state.get("format") != "segmamba_b200.train_step/1":
            raise ValueError(f"not a train_step checkpoint (format={state.get('format')!r})")
        module = self.model.module if hasattr(self.model, "module") else self.model
        if self.master_weights is not None:
            self.master_weights.load_state_dict(state["model"], strict=True)
        else:
            module.load_state_dict(state["model"], strict=True)
        self.optimizer.load_state_dict(state["optimizer"])
        if self.scheduler is not None and state.get("scheduler") is not None:
            self.scheduler.load_state_dict(state["scheduler"])
        if self.grad_scaler is not None and state.get("grad_scaler") is not None:
            self.grad_scaler.load_state_dict(state["grad_scaler"])
        self.global_step = int(state["global_step"])


def save_checkpoint(path: str, step: TrainStep) -> None:
    """atomic write (tmp file + rename) of the full training state; call on rank 0 only."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp"
    torch.save(step.state_dict(), tmp)
    os.replace(tmp, path)


def load_checkpoint(path: str, step: TrainStep, map_location="cpu") -> int:
    """restore the full training state; returns the global step to continue from."""
    step.load_state_dict(torch.load(path, map_location=map_location, weights_only=True))
    return step.global_step
