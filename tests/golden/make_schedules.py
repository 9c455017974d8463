"""Generate tests/golden/schedules_ref.json by running the REFERENCE's own scheduler factory,
    co3d_3d/src/modules/optim.py  (get_optimizer, get_scheduler, GradualWarmupScheduler),
imported unchanged from /root/reference on top of `nerf_downstream_b200.ginlite` standing in for gin-config (not
installed here).  For every case the optimiser is stepped like Lightning does with interval "step"
(optimizer.step(); scheduler.step()) and the learning rate (and SGD momentum, which CyclicLR cycles) in force for each
step is recorded.

Run from the repository root:  python tests/golden/make_schedules.py
"""
import importlib
import inspect
import json
import sys
import types
import warnings
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")

from nerf_downstream_b200 import ginlite  # noqa: E402

CASES = [
    # name, scheduler, warmup, steps recorded, gin bindings
    ("cosine_cls", "CosineAnnealingLR", -1, 400, ["train.max_steps=400", "train.lr=0.1"]),
    ("cosine_warm", "CosineAnnealingLR", 25, 325, ["train.max_steps=300", "train.lr=0.1"]),
    ("cosine_epoch", "CosineAnnealingLR", -1, 40, ["train.max_steps=300", "train.max_epochs=40",
                                                   "train.scheduler_interval='epoch'", "train.lr=0.05"]),
    ("poly", "PolyLR", -1, 300, ["train.max_steps=300", "train.lr=0.1", "PolyLR.poly_exp=0.9"]),
    ("poly_warm", "PolyLR", 10, 210, ["train.max_steps=200", "train.lr=0.02", "PolyLR.poly_exp=0.9"]),
    ("squared", "SquaredLR", -1, 120, ["train.max_steps=120", "train.lr=0.1", "SquaredLR.max_iter=120"]),
    ("step", "StepLR", -1, 100, ["train.max_steps=100", "train.lr=0.1", "StepLR.step_size=30", "StepLR.gamma=0.5"]),
    ("multistep", "MultiStepLR", -1, 80, ["train.max_steps=80", "train.lr=0.1", "MultiStepLR.milestones=[20, 50]"]),
    ("multistep_warm", "MultiStepLR", 5, 80, ["train.max_steps=80", "train.lr=0.1",
                                              "MultiStepLR.milestones=[20, 50]"]),
    ("exponential", "ExponentialLR", -1, 60, ["train.max_steps=60", "train.lr=0.1"]),
    ("cyclic_tri", "CyclicLR", -1, 100, ["train.max_steps=100", "train.lr=0.1", "CyclicLR.base_lr=0.01",
                                         "CyclicLR.step_size_up=13", "CyclicLR.mode='triangular'"]),
    ("cyclic_tri2", "CyclicLR", -1, 100, ["train.max_steps=100", "train.lr=0.1", "CyclicLR.base_lr=0.01",
                                          "CyclicLR.step_size_up=13", "CyclicLR.mode='triangular2'"]),
    ("cyclic_exp", "CyclicLR", -1, 100, ["train.max_steps=100", "train.lr=0.1", "CyclicLR.base_lr=0.01",
                                         "CyclicLR.step_size_up=13", "CyclicLR.mode='exp_range'",
                                         "CyclicLR.gamma=0.9"]),
    ("cyclic_cos", "CyclicLR", -1, 100, ["train.max_steps=100", "train.lr=0.1", "CyclicLR.base_lr=0.01",
                                         "CyclicLR.step_size_up=10", "CyclicLR.mode='cosine'"]),
]


def load_reference_optim():
    sys.modules["gin"] = ginlite
    for name in ("co3d_3d", "co3d_3d.src", "co3d_3d.src.modules"):
        mod = types.ModuleType(name)
        mod.__path__ = [str(REF / name.replace(".", "/"))]
        sys.modules[name] = mod
    return importlib.import_module("co3d_3d.src.modules.optim")


def accept_verbose():
    """The reference passes `verbose` through to torch's schedulers (optim.py:87-88,112-119,195-200); torch >= 2.7
    dropped that (purely cosmetic) argument.  Give torch's classes back a tolerant signature instead of touching the
    reference."""
    import torch.optim.lr_scheduler as L
    for cls in (L.CosineAnnealingLR, L.LambdaLR, L.MultiStepLR, L.StepLR, L.ExponentialLR):
        orig = cls.__init__
        n_pos = len(inspect.signature(orig).parameters)          # incl. self
        if "verbose" in inspect.signature(orig).parameters:
            continue

        def init(self, *args, _orig=orig, _n=n_pos, **kwargs):
            kwargs.pop("verbose", None)
            _orig(self, *args[:_n - 1], **kwargs)
        cls.__init__ = init


def main():
    accept_verbose()
    optim = load_reference_optim()
    out = {"torch": torch.__version__, "cases": {}}
    for name, sched, warm, steps, binds in CASES:
        ginlite.clear_config()
        ginlite.parse_config_files_and_bindings([], ["train.scheduler_interval='step'", "train.max_epochs=-1",
                                                     "SGD.momentum=0.9", *binds])
        lr = ginlite.query_parameter("train.lr")
        p = torch.nn.Parameter(torch.zeros(3))
        opt = optim.get_optimizer("SGD", [p], lr=lr, weight_decay=1e-4)
        assert opt.param_groups[0]["momentum"] == 0.9       # injected by the SGD.momentum binding
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sch = optim.get_scheduler(sched, opt, warm)
            lrs, moms = [], []
            for _ in range(steps):
                lrs.append(opt.param_groups[0]["lr"])
                moms.append(opt.param_groups[0]["momentum"])
                p.grad = torch.zeros(3)
                opt.step()
                sch.step()
        out["cases"][name] = {"scheduler": sched, "warmup_steps": warm, "bindings": binds, "repr": repr(sch),
                              "lr": lrs, "momentum": moms if sched == "CyclicLR" else None}
        print(name, repr(sch), lrs[:3], lrs[-1])
    path = Path(__file__).with_name("schedules_ref.json")
    path.write_text(json.dumps(out, indent=0))
    print("wrote", path)


if __name__ == "__main__":
    main()
