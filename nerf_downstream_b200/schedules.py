"""Learning-rate schedules of the reference trainer as closed forms of the step count.

The reference wraps `torch.optim.lr_scheduler` classes (co3d_3d/src/modules/optim.py:72-233), picks one by name
(`get_scheduler`, optim.py:313-330), optionally puts `GradualWarmupScheduler(multiplier=1.0)` in front
(optim.py:236-310) and lets Lightning call `.step()` once per optimiser step (`scheduler_interval = "step"`,
co3d_cls.gin:20).  The fused SGD step of this framework (`spc_sgd_step`) takes the learning rate as a scalar, so a
schedule here is a pure function `lr(t)`, t = number of `.step()` calls so far (torch's `last_epoch`).
`tests/test_training_host.py` pins every schedule against vectors produced by the reference's own `optim.py`
(tests/golden/make_schedules.py) and against the torch classes directly.
"""
from __future__ import annotations

import bisect
import math
from typing import Callable, Optional, Sequence


class Schedule:
    """lr(t) for t = 0, 1, 2, ...; `momentum(t)` is not None only for CyclicLR (torch cycles SGD momentum too)."""

    def __init__(self, name: str, fn: Callable[[int], float], momentum_fn: Optional[Callable[[int], float]] = None,
                 desc: str = ""):
        self.name, self._fn, self._mfn, self.desc = name, fn, momentum_fn, desc

    def lr(self, t: int) -> float:
        return self._fn(int(t))

    def momentum(self, t: int) -> Optional[float]:
        return None if self._mfn is None else self._mfn(int(t))

    __call__ = lr

    def __repr__(self):
        return f"{self.name}({self.desc})"


def poly(base_lr: float, max_steps: int, poly_exp: float) -> Schedule:
    """PolyLR: LambdaLR with (1 - t / (max_steps + 1)) ** poly_exp  (optim.py:181-204)."""
    return Schedule("PolyLR", lambda t: base_lr * (1 - t / (max_steps + 1)) ** poly_exp,
                    desc=f"max_steps={max_steps}, poly_exp={poly_exp}")


def squared(base_lr: float, max_iter: int) -> Schedule:
    """SquaredLR: (1 - t / (max_iter + 1)) ** 2  (optim.py:207-216)."""
    return Schedule("SquaredLR", lambda t: base_lr * (1 - t / (max_iter + 1)) ** 2, desc=f"max_iter={max_iter}")


def cosine(base_lr: float, t_max: int, eta_min: float = 0) -> Schedule:
    """CosineAnnealingLR with T_max = train.max_steps (interval "step") or train.max_epochs (optim.py:103-124).
    torch evaluates the chained recurrence; it equals this closed form up to rounding (tested to 1e-9 relative)."""
    return Schedule("CosineAnnealingLR",
                    lambda t: eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t / t_max)) / 2,
                    desc=f"T_max={t_max}, eta_min={eta_min}")


def step(base_lr: float, step_size: int, gamma: float = 0.1) -> Schedule:
    """StepLR (optim.py:72-74)."""
    return Schedule("StepLR", lambda t: base_lr * gamma ** (t // step_size), desc=f"step_size={step_size}, gamma={gamma}")


def multistep(base_lr: float, milestones: Sequence[int] = (20000, 40000), gamma: float = 0.1) -> Schedule:
    """MultiStepLR, reference defaults milestones [20000, 40000], gamma 0.1 (optim.py:77-89)."""
    ms = sorted(milestones)
    return Schedule("MultiStepLR", lambda t: base_lr * gamma ** bisect.bisect_right(ms, t),
                    desc=f"milestones={list(ms)}, gamma={gamma}")


def exponential(base_lr: float, gamma: float = 0.99) -> Schedule:
    """ExponentialLR, reference default gamma 0.99 (optim.py:92-100)."""
    return Schedule("ExponentialLR", lambda t: base_lr * gamma ** t, desc=f"gamma={gamma}")


def cyclic(base_lr: float, max_lr: float, step_size_up: int = 2000, mode: str = "triangular", gamma: float = 1.0,
           max_steps: Optional[int] = None, base_momentum: float = 0.8, max_momentum: float = 0.9) -> Schedule:
    """CyclicLR as the reference configures it (optim.py:146-178): max_lr = train.lr, step_size_down = step_size_up,
    scale per CYCLE; modes "triangular", "triangular2", "exp_range" (gamma ** cycle) and "cosine"
    ((1 + cos(pi cycle / (max_steps / (2 step_size_up)))) / 2).  Momentum is cycled inversely, as torch does."""
    total = 2.0 * step_size_up
    ratio = step_size_up / total
    if mode == "triangular":
        scale = lambda c: 1.0                                    # noqa: E731
    elif mode == "triangular2":
        scale = lambda c: 1.0 / (2.0 ** (c - 1))                 # noqa: E731
    elif mode == "exp_range":
        scale = lambda c: gamma ** c                             # noqa: E731
    elif mode == "cosine":
        if max_steps is None:
            raise ValueError("mode 'cosine' needs max_steps (train.max_steps)")
        n_cycles = max_steps / (2 * step_size_up)
        scale = lambda c: (1 + math.cos(c / n_cycles * math.pi)) / 2   # noqa: E731
    else:
        raise ValueError(f"Invalid mode:{mode}")

    def factor(t):
        cycle = math.floor(1 + t / total)
        x = 1.0 + t / total - cycle
        f = x / ratio if x <= ratio else (x - 1) / (ratio - 1)
        return f * scale(cycle)

    return Schedule("CyclicLR", lambda t: base_lr + (max_lr - base_lr) * factor(t),
                    lambda t: max_momentum - (max_momentum - base_momentum) * factor(t),
                    desc=f"max_lr={max_lr}, base_lr={base_lr}, step_size_up={step_size_up}, mode={mode}, gamma={gamma}")


def warmup(after: Schedule, base_lr: float, warmup_steps: int) -> Schedule:
    """GradualWarmupScheduler(multiplier=1.0, total_epoch=warmup_steps) in front of `after` (optim.py:236-310,
    324-328): lr = base_lr * t / warmup_steps for t <= warmup_steps; step warmup_steps + 1 hands over to the wrapped
    scheduler at ITS step 0, which then advances once per step."""
    w = int(warmup_steps)

    def fn(t):
        return base_lr * (t / w) if t <= w else after.lr(t - w - 1)

    mfn = None
    if after._mfn is not None:
        mfn = lambda t: after.momentum(max(t - w - 1, 0))        # noqa: E731
    return Schedule("GradualWarmupScheduler", fn, mfn, desc=f"warmup_steps={w}, after={after!r}")


NAMES = ("StepLR", "MultiStepLR", "ExponentialLR", "CosineAnnealingLR", "CyclicLR", "PolyLR", "SquaredLR")


def get_schedule(scheduler_name: str, lr: float, max_steps: int, warmup_steps: Optional[int] = None,
                 max_epochs: int = -1, scheduler_interval: str = "step", **kw) -> Optional[Schedule]:
    """`get_scheduler(scheduler_name, optimizer, warmup_steps)` (optim.py:313-330).  Arguments the reference binds
    through gin (`PolyLR.poly_exp`, `StepLR.step_size`, `CyclicLR.base_lr` ...) are keyword arguments here; the
    ones without a default in the reference are required here too.  "none" (any case) -> None, as
    lightning_module_base.py:108-116."""
    if scheduler_name.lower() == "none":
        return None
    if scheduler_name not in NAMES:
        raise ValueError(f"optimizer {scheduler_name} not recognized in {list(NAMES)}.")
    if scheduler_name == "PolyLR":
        s = poly(lr, max_steps, kw.pop("poly_exp"))
    elif scheduler_name == "SquaredLR":
        s = squared(lr, kw.pop("max_iter"))
    elif scheduler_name == "CosineAnnealingLR":
        s = cosine(lr, max_steps if scheduler_interval == "step" else max_epochs, kw.pop("eta_min", 0))
    elif scheduler_name == "StepLR":
        s = step(lr, kw.pop("step_size"), kw.pop("gamma", 0.1))
    elif scheduler_name == "MultiStepLR":
        s = multistep(lr, kw.pop("milestones", (20000, 40000)), kw.pop("gamma", 0.1))
    elif scheduler_name == "ExponentialLR":
        s = exponential(lr, kw.pop("gamma", 0.99))
    else:
        s = cyclic(kw.pop("base_lr"), lr, kw.pop("step_size_up", 2000), kw.pop("mode", "trianglular"),
                   kw.pop("gamma", 1.0), max_steps)
    if kw:
        raise TypeError(f"{scheduler_name}: unexpected arguments {sorted(kw)}")
    if warmup_steps is not None and warmup_steps > 0:
        s = warmup(s, lr, warmup_steps)
    return s
