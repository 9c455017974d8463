"""Trainer + config parity for the hot path.

What the reference's Lightning glue computes around the sparse network, restated as a plain loop over the
data-parallel step of `trainer.DataParallelTrainer` — same configuration names (`train.*` gin bindings,
co3d_3d/train.py:50-96), same losses, metrics, schedules and checkpoint layout, so that a `*.gin` file of the
reference drives a run here and a Lightning checkpoint of the reference evaluates here:

  * losses   — `F.cross_entropy` (classification_training.py:33) and `SegLoss` (segmentation_training.py:27-44:
               class weights all one except an optional `void_weight` on the LAST class, `ignore_index`), with the
               `use_sync_grad` re-weighting by point counts (segmentation_training.py:112-120);
  * metrics  — top-1 / top-5 (classification_training.py:82-97 and torchmetrics `Accuracy`), overall accuracy and
               mIoU of a step (`_eval_metrics`, segmentation_training.py:230-239 with utils/__init__.py:103-130),
               the epoch `IoUMeter` (metrics.py:5-58; the counting kernel lives in `pipeline.py`);
  * schedule — `schedules.get_schedule`, stepped once per optimiser step, total steps = max_steps + warm-up
               (train.py:175);
  * checkpoints — Lightning layout `{"state_dict": {"model.<key>": ...}, "optimizer_states": [...],
               "lr_schedulers": [...], "global_step", "epoch"}` (eval.py:47-67, lightning_module_base.py:75-105),
               incl. `convert_self_supervised_checkpoint` (lightning_module_base.py:62-72).

Datasets, augmentations, loggers and the Lightning runtime itself stay out of scope: batches are
whatever iterable of `{"coordinates", "features", "labels"}` dicts the caller provides (`data/utils.py:25-50`).
Nothing here falls back to a CPU model: the network runs through the CUDA library or not at all.
"""
from __future__ import annotations

import json
import math
import os
from collections import OrderedDict
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ginlite, schedules

EPS = 1e-10     # utils/__init__.py:7


# ---- metrics -------------------------------------------------------------------------------------------------
@torch.no_grad()
def accuracy_topk(output: torch.Tensor, target: torch.Tensor, topk: Sequence[int] = (1,)) -> List[float]:
    """Percent of rows whose target is among the k largest logits (classification_training.py:82-97)."""
    maxk = min(max(topk), output.shape[1])          # (the reference raises with fewer than max(topk) classes)
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1))
    return [correct[:k].flatten().float().sum().mul(100.0 / batch_size).item() for k in topk]


@torch.no_grad()
def precision_at_one(pred: torch.Tensor, target: torch.Tensor, ignore_label: int = 255) -> float:
    """Percent of non-ignored points predicted correctly; NaN when every point is ignored (utils/__init__.py:103-114)."""
    keep = target.reshape(-1) != ignore_label
    n = int(keep.sum())
    if n == 0:
        return float("nan")
    return (pred.reshape(-1)[keep] == target.reshape(-1)[keep]).float().sum().mul(100.0 / n).item()


@torch.no_grad()
def fast_hist(pred: torch.Tensor, label: torch.Tensor, n: int) -> np.ndarray:
    """Confusion matrix hist[label, pred] over points with 0 <= label < n (utils/__init__.py:117-123) — note the
    reference filters on the label RANGE here, not on the ignore label."""
    k = (label >= 0) & (label < n)
    count = torch.bincount(n * label[k].long() + pred[k].long(), minlength=n * n)
    return count.cpu().numpy().reshape(n, n)


def per_class_iu(hist: np.ndarray) -> np.ndarray:
    """diag / (row sum + column sum - diag + EPS) (utils/__init__.py:126-128)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist) + EPS)


@torch.no_grad()
def eval_metrics(output: torch.Tensor, target: torch.Tensor, num_labels: int, ignore_label: int = 255) -> Dict[str, float]:
    """`SegmentationTraining._eval_metrics` (segmentation_training.py:230-239): {"OA", "mIoU"} in percent."""
    pred = output.argmax(1)
    hist = fast_hist(pred, target, num_labels)
    return {"OA": precision_at_one(pred, target, ignore_label), "mIoU": float((per_class_iu(hist) * 100).mean())}


def metrics_from_counts(counts: torch.Tensor) -> Dict[str, float]:
    """{"OA", "mIoU"} of `eval_metrics` from the [3, C] seen / correct / predicted counts the fused segmentation head
    accumulates (`pipeline.seg_head_loss(counts=...)`): the confusion-matrix row sums, diagonal and column sums."""
    seen, correct, positive = (counts[i].double().cpu().numpy() for i in range(3))
    total = seen.sum()
    oa = float(correct.sum() * 100.0 / total) if total > 0 else float("nan")
    with np.errstate(divide="ignore", invalid="ignore"):
        iu = correct / (seen + positive - correct + EPS)
    return {"OA": oa, "mIoU": float((iu * 100).mean())}


def count_parameters(model: torch.nn.Module) -> Dict[str, float]:
    """`count_parameters` (co3d_3d/src/utils/prune.py:11-23): trainable parameters and the number removed by
    `torch.nn.utils.prune` masks (`kernel_mask` / `weight_mask` buffers) — logged as val/total_params, val/pruned_params."""
    return {"total": float(sum(p.numel() for p in model.parameters() if p.requires_grad)),
            "pruned": float(sum((1 - b.int()).sum() for n, b in model.named_buffers()
                                if "kernel_mask" in n or "weight_mask" in n))}


def get_parameters_to_prune(model: torch.nn.Module):
    """`get_parameters_to_prune` (utils/prune.py:34-59): (module, parameter name) of every sparse convolution kernel and
    every MinkowskiLinear weight."""
    from . import me as ME
    out = []
    for _, module in model.named_modules():
        if isinstance(module, (ME.MinkowskiConvolution, ME.MinkowskiConvolutionTranspose)):
            out.append((module, "kernel"))
        elif isinstance(module, ME.MinkowskiLinear):
            out.append((module.linear, "weight"))
    return out


class AccuracyMeter:
    """torchmetrics `Accuracy(num_classes, top_k)` as the reference uses it (classification_training.py:12-18,59-60,
    67-68): accumulates #correct / #seen over `update(logits, labels)` calls, `compute()` returns the fraction."""

    def __init__(self, num_classes: int, top_k: int = 1):
        self.num_classes, self.top_k = num_classes, top_k
        self.reset()

    def reset(self):
        self.correct, self.total = 0, 0

    @torch.no_grad()
    def update(self, logits: torch.Tensor, labels: torch.Tensor):
        k = min(self.top_k, logits.shape[1])
        top = logits.topk(k, 1, True, True).indices
        self.correct += int((top == labels.view(-1, 1)).any(1).sum())
        self.total += int(labels.numel())

    __call__ = update

    def compute(self) -> float:
        return self.correct / self.total if self.total else float("nan")

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized():
            t = torch.tensor([self.correct, self.total], dtype=torch.int64)
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.all_reduce(t, group=group)
            self.correct, self.total = int(t[0]), int(t[1])


# ---- losses --------------------------------------------------------------------------------------------------
class SegLoss(torch.nn.Module):
    """segmentation_training.py:27-44.  weight = ones(num_labels), weight[-1] = void_weight when given and > 0.
    On CUDA logits with unit weights this is the fused `spc_ce_fwd/bwd` kernel; with a void weight the weighted
    kernel (`spc_ce_fwd_weighted`)."""

    def __init__(self, ignore_index: int, num_labels: int, void_weight: Optional[float] = None):
        super().__init__()
        self.ignore_index = ignore_index
        weight = torch.ones(num_labels)
        self.weighted = void_weight is not None and void_weight > 0
        if self.weighted:
            weight[-1] = void_weight
        self.register_buffer("weight", weight, persistent=False)

    def forward(self, output: torch.Tensor, batch) -> torch.Tensor:
        labels = (batch["labels"] if isinstance(batch, dict) else batch).long().to(output.device)
        if output.is_cuda:
            from . import ops
            return ops.cross_entropy(output, labels, self.ignore_index,
                                     self.weight.to(output.device) if self.weighted else None)
        return F.cross_entropy(output, labels, weight=self.weight.to(output.device), ignore_index=self.ignore_index)


def sync_grad_scale(num_points: int, group=None) -> float:
    """`training_step_end` with use_sync_grad (segmentation_training.py:112-120): the factor a rank's loss is multiplied
    with, num_points / sum(all ranks' num_points) * world — so that after DDP's gradient MEAN every point of the global
    batch weighs the same."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    t = torch.tensor([float(num_points)], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    total = sum(float(g) for g in gathered)
    return float(num_points) / total * world


# ---- checkpoints (Lightning layout) --------------------------------------------------------------------------
def convert_self_supervised_checkpoint(state_dict) -> "OrderedDict[str, torch.Tensor]":
    """lightning_module_base.py:62-72: drop `predictor` / `final` entries, `model.encoder` -> `model`."""
    out = OrderedDict()
    for k, v in state_dict.items():
        if "predictor" in k or "final" in k:
            continue
        out[k.replace("model.encoder", "model")] = v
    return out


def lightning_state_dict(model: torch.nn.Module) -> "OrderedDict[str, torch.Tensor]":
    """The module's state dict under the `model.` prefix Lightning gives it (`BaseModule.model`, eval.py:47-67)."""
    return OrderedDict((f"model.{k}", v.detach().clone()) for k, v in model.state_dict().items())


def load_lightning_state_dict(model: torch.nn.Module, state_dict, strict: bool = True):
    """Load `checkpoint["state_dict"]` of the reference (keys `model.<...>`; entries of the Lightning module itself,
    e.g. `criterion.weight`, are not the network's and are skipped)."""
    own = OrderedDict((k[len("model."):], v) for k, v in state_dict.items() if k.startswith("model."))
    if not own and len(state_dict):
        raise KeyError("no 'model.'-prefixed entries: not a checkpoint of the reference's training modules")
    return model.load_state_dict(own, strict=strict)


def optimizer_state(trainer) -> dict:
    """`torch.optim.SGD.state_dict()` layout (what Lightning stores in `optimizer_states`): parameters are numbered in
    `model.parameters()` order, each with its `momentum_buffer`."""
    if getattr(trainer, "torch_optimizer", None) is not None:
        # torch numbers parameters in the order they were handed over (the arena's); Lightning stores exactly this
        return trainer.torch_optimizer.state_dict()
    params = [p for p in trainer.model.parameters() if p.requires_grad]
    slot = {id(p): o for p, o in zip(trainer.arena.order, trainer.arena.offsets)}
    state = {}
    if trainer.steps > 0:
        for i, p in enumerate(params):
            o = slot[id(p)]
            state[i] = {"momentum_buffer": trainer.momentum_buf[o:o + p.numel()].view_as(p).detach().clone()}
    group = {"lr": trainer.lr, "momentum": trainer.momentum, "dampening": 0, "weight_decay": trainer.weight_decay,
             "nesterov": False, "maximize": False, "foreach": None, "differentiable": False, "fused": None,
             "initial_lr": getattr(trainer, "initial_lr", trainer.lr), "params": list(range(len(params)))}
    return {"state": state, "param_groups": [group]}


def load_optimizer_state(trainer, opt_state: dict, keep_lr: Optional[float] = None) -> None:
    """Inverse of `optimizer_state`.  `keep_lr`: start from the configured learning rate instead of the stored one
    (the intent of lightning_module_base.py:93-101)."""
    if getattr(trainer, "torch_optimizer", None) is not None:
        trainer.torch_optimizer.load_state_dict(opt_state)
        trainer.set_lr(trainer.torch_optimizer.param_groups[0]["lr"] if keep_lr is None else keep_lr)
        trainer.steps = max(trainer.steps, 1)
        return
    params = [p for p in trainer.model.parameters() if p.requires_grad]
    slot = {id(p): o for p, o in zip(trainer.arena.order, trainer.arena.offsets)}
    group = opt_state["param_groups"][0]
    if len(group["params"]) != len(params):
        raise ValueError(f"optimizer state holds {len(group['params'])} parameters, the model has {len(params)}")
    state = opt_state.get("state", {})
    trainer.momentum_buf.zero_()
    loaded = 0
    for i, p in enumerate(params):
        st = state.get(i, state.get(str(i)))
        if st is None or st.get("momentum_buffer") is None:
            continue
        o = slot[id(p)]
        trainer.momentum_buf[o:o + p.numel()].view_as(p).copy_(st["momentum_buffer"])
        loaded += 1
    trainer.momentum, trainer.weight_decay = group["momentum"], group["weight_decay"]
    trainer.lr = group["lr"] if keep_lr is None else keep_lr
    # spc_sgd_step initialises the buffer with the first gradient (torch: buf = grad on the first step)
    trainer.steps = max(trainer.steps, 1) if loaded else 0


def save_checkpoint(path: str, trainer, global_step: int, epoch: int = 0, extra: Optional[dict] = None) -> None:
    ckpt = {"epoch": epoch, "global_step": global_step, "pytorch-lightning_version": "1.5.10",
            "state_dict": lightning_state_dict(trainer.model),
            "optimizer_states": [optimizer_state(trainer)],
            "lr_schedulers": [{"last_epoch": global_step, "_step_count": global_step + 1,
                               "_last_lr": [trainer.lr]}]}
    ckpt.update(extra or {})
    tmp = f"{path}.tmp"
    torch.save(ckpt, tmp)
    os.replace(tmp, path)


def load_checkpoint(path: str, trainer=None, model: Optional[torch.nn.Module] = None, load_weights: bool = True,
                    load_optimizers: bool = False, transfer_self_supervised: bool = False, lr: Optional[float] = None,
                    map_location="cpu") -> dict:
    """`configure_optimizers` checkpoint handling (lightning_module_base.py:82-105): weights (optionally converted from
    a self-supervised run, non-strict) and / or optimizer state."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model = model if model is not None else trainer.model
    if load_weights:
        sd = ckpt["state_dict"]
        if transfer_self_supervised:
            load_lightning_state_dict(model, convert_self_supervised_checkpoint(sd), strict=False)
        else:
            load_lightning_state_dict(model, sd)
    if load_optimizers:
        if trainer is None:
            raise ValueError("load_optimizers needs the trainer")
        load_optimizer_state(trainer, ckpt["optimizer_states"][0], keep_lr=lr)
    return ckpt


# ---- batches ---------------------------------------------------------------------------------------------------
def collate_mink(list_data: Sequence[dict]) -> dict:
    """`collate_mink` (co3d_3d/src/data/utils.py:25-50): samples `{"coordinates" [n_i, 3], "features" [n_i, C],
    "labels" ...}` -> one batch with float32 `(b, x, y, z)` coordinates (`ME.utils.sparse_collate(dtype=float32)`),
    concatenated features, every `*label*` / `*instance*` / `*dists*` entry concatenated, `metadata` / `dataset` /
    `colors` kept as lists."""
    from .me import utils as me_utils
    coordinates_batch, features_batch = me_utils.sparse_collate([d["coordinates"] for d in list_data],
                                                                [d["features"] for d in list_data], dtype=torch.float32)
    package = {"coordinates": coordinates_batch, "features": features_batch}
    for key in list_data[0].keys():
        if "label" in key or "instance" in key or "dists" in key:
            package[key] = torch.from_numpy(np.concatenate([np.asarray(d[key]) for d in list_data]))
    for key in ("metadata", "dataset", "colors"):
        if key in list_data[0]:
            package[key] = [d[key] for d in list_data]
    return package


# ---- the run ---------------------------------------------------------------------------------------------------
class TrainConfig:
    """The `train.*` bindings of a gin file with the defaults of `train()` (co3d_3d/train.py:50-96)."""
    DEFAULTS = dict(max_steps=None, max_epochs=-1, warmup_steps=-1, training_module="SegmentationTraining",
                    optimizer_name="SGD", scheduler_name="PolyLR", scheduler_interval="step", lr=1e-3,
                    weight_decay=1e-4, batch_size=8, val_batch_size=6, val_every_n_steps=1000, log_every_n_steps=10,
                    resume_training=False, checkpoint_path=None, load_weights=False, load_optimizers=False,
                    transfer_self_supervised=False, use_sync_batchnorm=False, use_sync_grad=False, ignore_label=-100,
                    monitor_metric="val/mIoU", void_weight=None, gpus=1)

    def __init__(self, **overrides):
        bound = {k.split(".", 1)[1]: v for k, v in ginlite.config_dict().items() if k.startswith("train.")}
        unknown = set(overrides) - set(self.DEFAULTS)
        if unknown:
            raise TypeError(f"unknown train() arguments: {sorted(unknown)}")
        for k, v in self.DEFAULTS.items():
            setattr(self, k, overrides.get(k, bound.get(k, v)))
        if self.max_steps is None:
            raise ginlite.GinError("train.max_steps is not bound (required argument of train(), train.py:56)")
        self.momentum = _bound(f"{self.optimizer_name}.momentum", 0.0)
        self.optimizer_kwargs = {k.split(".", 1)[1]: v for k, v in ginlite.config_dict().items()
                                 if k.startswith(f"{self.optimizer_name}.")}
        self.scheduler_kwargs = {k.split(".", 1)[1]: v for k, v in ginlite.config_dict().items()
                                 if k.startswith(f"{self.scheduler_name}.")}

    @property
    def total_steps(self) -> int:
        """`max_steps + warmup_steps` optimiser steps (train.py:175)."""
        return self.max_steps + (self.warmup_steps if self.warmup_steps and self.warmup_steps > 0 else 0)

    def schedule(self) -> Optional[schedules.Schedule]:
        return schedules.get_schedule(self.scheduler_name, self.lr, self.max_steps, self.warmup_steps, self.max_epochs,
                                      self.scheduler_interval, **self.scheduler_kwargs)


def _bound(key: str, default):
    try:
        return ginlite.query_parameter(key)
    except ValueError:
        return default


def get_model(name: Optional[str] = None, in_channel: Optional[int] = None, out_channel: Optional[int] = None):
    """`get_model` (src/models/__init__.py:18-20) over the networks of `models.py`, arguments from `get_model.*`."""
    from . import models
    name = name if name is not None else ginlite.query_parameter("get_model.name")
    in_channel = in_channel if in_channel is not None else ginlite.query_parameter("get_model.in_channel")
    out_channel = out_channel if out_channel is not None else ginlite.query_parameter("get_model.out_channel")
    if name not in models.MODELS:
        raise KeyError(f"model {name!r} is not built here (have {sorted(models.MODELS)})")
    return models.MODELS[name](in_channel, out_channel)


class Run:
    """One training run: `fit()` is `trainer.fit` of train.py:163-186 for a single rank of the data-parallel job.

    model      — a network over the ME-compatible surface (takes a TensorField; `models.py` or the reference's files)
    make_input — batch dict -> network input (default: `ME.TensorField(coordinates=, features=)`, base_model.py:10-13)
    fused_head — segmentation only, for models with `forward_sparse` (models.SparseResUNet): slice, SegLoss and the
                 metric counts run as ONE kernel over the points (`pipeline.seg_head_loss`) instead of three passes
    evaluate_only — no optimiser state, no schedule: only `validate()` is usable (`evaluate()`, eval.py)
    exception_safe — `ExceptionSafeSegmentationTraining` (segmentation_training.py:233-302): a step or validation batch
                 that raises RuntimeError (the library reports every failure, incl. out-of-memory, that way) is
                 counted and skipped; the schedule advances regardless
    """

    def __init__(self, model: torch.nn.Module, cfg: TrainConfig, num_labels: Optional[int] = None,
                 void_label=None, save_path: Optional[str] = None, make_input: Optional[Callable] = None,
                 log: Optional[Callable[[dict], None]] = None, fused_head: bool = False, evaluate_only: bool = False,
                 exception_safe: bool = False):
        from . import trainer as T
        self.model, self.cfg, self.save_path, self.log = model, cfg, save_path, log or (lambda d: None)
        self.segmentation = cfg.training_module == "SegmentationTraining"
        if cfg.training_module not in ("SegmentationTraining", "ClassificationTraining"):
            raise AssertionError(f"{cfg.training_module} not in ['SegmentationTraining', 'ClassificationTraining']")
        self.num_labels = num_labels if num_labels is not None else ginlite.query_parameter("get_model.out_channel")
        self.trainer = None
        self.schedule = None
        if not evaluate_only:
            self.trainer = T.DataParallelTrainer(model, lr=cfg.lr, momentum=cfg.momentum, weight_decay=cfg.weight_decay,
                                                 optimizer_name=cfg.optimizer_name,
                                                 optimizer_kwargs=None if cfg.optimizer_name == "SGD" else cfg.optimizer_kwargs)
            self.trainer.initial_lr = cfg.lr
            self.schedule = cfg.schedule()
        self.global_step = 0
        self.best = -math.inf
        self.make_input = make_input or self._tensor_field
        self.fused_head = bool(fused_head)
        self.exception_safe, self.fail_count = bool(exception_safe), 0
        if self.fused_head and not (self.segmentation and hasattr(model, "forward_sparse")):
            raise ValueError("fused_head needs SegmentationTraining and a model with forward_sparse()")
        if self.segmentation:
            self.criterion = SegLoss(cfg.ignore_label, self.num_labels, cfg.void_weight)
            from .pipeline import IoUMeter
            self.iou_meter = IoUMeter(self.num_labels, cfg.ignore_label, void_label)
        else:
            self.acc1_meter = AccuracyMeter(self.num_labels, 1)
            self.acc5_meter = AccuracyMeter(self.num_labels, 5)
        if evaluate_only:
            return
        if cfg.checkpoint_path is not None and (cfg.load_weights or cfg.load_optimizers or cfg.resume_training):
            ckpt = load_checkpoint(cfg.checkpoint_path, self.trainer, load_weights=cfg.load_weights or cfg.resume_training,
                                   load_optimizers=cfg.load_optimizers or cfg.resume_training,
                                   transfer_self_supervised=cfg.transfer_self_supervised,
                                   lr=None if cfg.resume_training else cfg.lr)
            if cfg.resume_training:
                self.global_step = int(ckpt.get("global_step", 0))
        self._apply_schedule()

    @staticmethod
    def _tensor_field(batch):
        from . import me as ME
        return ME.TensorField(coordinates=batch["coordinates"], features=batch["features"])

    def _apply_schedule(self):
        if self.schedule is not None:
            self.trainer.set_lr(self.schedule.lr(self.global_step))
            m = self.schedule.momentum(self.global_step)
            if m is not None:
                self.trainer.momentum = m

    # -- one optimiser step (training_step + backward + optimizer.step + scheduler.step) -------------------------
    def training_step(self, batch) -> Optional[torch.Tensor]:
        if not self.exception_safe:
            return self._training_step(batch)
        try:
            return self._training_step(batch)
        except RuntimeError as e:               # segmentation_training.py:276-283
            self.fail_count += 1
            print(f"Failed with {e}. Failure rate: {float(self.fail_count) / (self.global_step + 1)}")
            self.trainer.abort_step()           # zero gradients; at world > 1 keep the collectives matched
            self.global_step += 1               # "regardless of the failure status, update" the scheduler
            self._apply_schedule()
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            return None

    def _to_device(self, batch):
        """Lightning moves every tensor of the collated batch to the module's device before training_step /
        validation_step; `collate_mink` (data/utils.py:25-50) returns host tensors."""
        dev = next(self.model.parameters()).device
        if dev.type != "cuda":
            return batch
        return {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) and v.device != dev else v)
                for k, v in batch.items()}

    def _training_step(self, batch) -> torch.Tensor:
        self.model.train()
        batch = self._to_device(batch)
        labels = batch["labels"].long()
        step_counts = None
        if self.fused_head:
            # slice + SegLoss + metric counts in one kernel over the points; logits stay on the voxel rows
            from . import pipeline
            field = self.make_input(batch)
            sparse_out = self.model.forward_sparse(field)
            logits = sparse_out.F
            step_counts = torch.zeros((3, logits.shape[1]), dtype=torch.int64, device=logits.device)
            crit = self.criterion
            loss = pipeline.seg_head_loss(sparse_out, field, labels.to(logits.device), crit.ignore_index,
                                          crit.weight.to(logits.device) if crit.weighted else None, step_counts)
        else:
            logits = self.model(self.make_input(batch))
            if self.segmentation:
                loss = self.criterion(logits, batch)
            else:
                from . import ops
                loss = ops.cross_entropy(logits, labels) if logits.is_cuda else F.cross_entropy(logits, labels)
        if self.segmentation and self.cfg.use_sync_grad:
            loss = loss * sync_grad_scale(batch["coordinates"].shape[0])
        step = self.global_step
        if step % self.cfg.log_every_n_steps == 0 and step > 0:
            loss_float = loss.detach().cpu().item()
            if not np.isfinite(loss_float):
                raise ValueError(f"Invalid loss: {loss_float}")
            from . import ops as _ops
            _ops.raise_on_bad_targets()         # a label outside [0, C) that is not ignore_index (F.cross_entropy asserts)
            out = {"train/loss": loss_float, "train/lr": self.trainer.lr, "global_step": step}
            if self.segmentation:
                metrics = metrics_from_counts(step_counts) if step_counts is not None else \
                    eval_metrics(logits.detach(), labels, logits.shape[1], self.cfg.ignore_label)
                for k, v in metrics.items():
                    out[f"train/{k}"] = v
                out["train/ignore_ratio"] = ((labels == self.cfg.ignore_label).sum() / labels.shape[0] * 100).item()
            else:
                out["train/acc1"], out["train/acc5"] = accuracy_topk(logits.detach(), labels, (1, 5))
            self.log(out)
        self.trainer.backward_and_step(loss)
        self.global_step += 1
        self._apply_schedule()
        return loss.detach()

    # -- validation epoch ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def validate(self, batches: Iterable[dict]) -> Dict[str, float]:
        self.model.eval()
        losses, oas = [], []
        if self.segmentation:
            self.iou_meter.counts = None
        else:
            self.acc1_meter.reset()
            self.acc5_meter.reset()
        for batch in batches:
            if self.exception_safe:
                try:
                    self._validate_batch(batch, losses, oas)
                except RuntimeError as e:       # segmentation_training.py:316-326
                    print(f"Validation step failed with {e}.")
                    if torch.cuda.is_available():
                        torch.cuda.synchronize()
            else:
                self._validate_batch(batch, losses, oas)
        return self._validation_epoch_end(losses, oas)

    def _validate_batch(self, batch, losses, oas) -> None:
        batch = self._to_device(batch)
        labels = batch["labels"].long()
        if self.fused_head:
            from . import pipeline
            field = self.make_input(batch)
            sparse_out = self.model.forward_sparse(field)
            dev = sparse_out.F.device
            step_counts = torch.zeros((3, self.num_labels), dtype=torch.int64, device=dev)
            crit = self.criterion
            losses.append(pipeline.seg_head_loss(sparse_out, field, labels.to(dev), crit.ignore_index,
                                                 crit.weight.to(dev) if crit.weighted else None, step_counts).item())
            oas.append(metrics_from_counts(step_counts)["OA"])
            self.iou_meter.counts_buffer(dev).add_(step_counts)
            return
        logits = self.model(self.make_input(batch))
        if self.segmentation:
            losses.append(self.criterion(logits, batch).item())
            oas.append(eval_metrics(logits, labels, logits.shape[1], self.cfg.ignore_label)["OA"])
            self.iou_meter.update(logits, labels)
        else:
            losses.append(F.cross_entropy(logits, labels).item())
            self.acc1_meter(logits, labels)
            self.acc5_meter(logits, labels)

    def _validation_epoch_end(self, losses, oas) -> Dict[str, float]:
        from . import ops as _ops
        _ops.raise_on_bad_targets()
        assert len(losses) > 0
        out = {"val/loss": float(np.mean(losses)), "global_step": self.global_step}
        # epoch metrics are sums over ALL ranks' validation shards (torchmetrics dist_reduce_fx="sum", metrics.py:17-28)
        if self.segmentation:
            self.iou_meter.all_reduce()
        else:
            self.acc1_meter.all_reduce()
            self.acc5_meter.all_reduce()
        if self.segmentation:
            miou, ious, macc, accs = self.iou_meter.compute()
            out.update({"val/OA": float(np.mean(oas)), "val/mIoU": float(miou) * 100, "val/mAcc": float(macc) * 100})
            if out["val/mIoU"] > self.best:
                self.best = out["val/mIoU"]
            out["val/best_mIoU"] = self.best
            if self.save_path:
                with open(os.path.join(self.save_path, "eval_results.json"), "w") as f:
                    json.dump({"iou": [*(ious * 100).cpu().tolist(), float(miou)],
                               "acc": [*(accs * 100).cpu().tolist(), float(macc)]}, f)
        else:
            out.update({"val/acc1": self.acc1_meter.compute(), "val/acc5": self.acc5_meter.compute()})
        for k, v in count_parameters(self.model).items():          # segmentation_training.py:204-206
            out[f"val/{k}_params"] = v
        self.log(out)
        return out

    # -- the loop --------------------------------------------------------------------------------------------------
    def fit(self, train_batches: Callable[[], Iterable[dict]], val_batches: Optional[Callable[[], Iterable[dict]]] = None):
        """`train_batches()` / `val_batches()` return a fresh iterable per epoch.  Stops after `total_steps` optimiser
        steps (or `max_epochs` epochs when > 0); validates every `val_every_n_steps`; keeps `last.ckpt` and the best
        checkpoint by `monitor_metric` (ModelCheckpoint(save_top_k=1, save_last=True, mode="max"), train.py:150-157)."""
        cfg = self.cfg
        rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
        if self.save_path and rank0:
            os.makedirs(self.save_path, exist_ok=True)
        epoch, best_monitor, last_val = 0, -math.inf, None
        while self.global_step < cfg.total_steps and (cfg.max_epochs <= 0 or epoch < cfg.max_epochs):
            seen = 0
            for batch in train_batches():
                if self.global_step >= cfg.total_steps:
                    break
                self.training_step(batch)
                seen += 1
                if val_batches is not None and self.global_step % cfg.val_every_n_steps == 0:
                    last_val = self.validate(val_batches())
                    monitor = last_val.get(cfg.monitor_metric)
                    if self.save_path and rank0:
                        save_checkpoint(os.path.join(self.save_path, "last.ckpt"), self.trainer, self.global_step, epoch)
                        if monitor is not None and monitor > best_monitor:
                            best_monitor = monitor
                            save_checkpoint(os.path.join(self.save_path, "best.ckpt"), self.trainer, self.global_step,
                                            epoch, {"monitor": {cfg.monitor_metric: monitor}})
            if seen == 0:
                raise RuntimeError("train_batches() produced no batch")
            epoch += 1
        if self.save_path and rank0:
            save_checkpoint(os.path.join(self.save_path, "last.ckpt"), self.trainer, self.global_step, epoch)
        return last_val


def _select_device(device) -> None:
    """The library launches on the CURRENT device's current stream (one process per GPU): make `device` current."""
    d = torch.device(device)
    if d.type == "cuda" and torch.cuda.is_available():
        torch.cuda.set_device(d.index if d.index is not None else torch.cuda.current_device())


def train(config_files: Sequence[str], bindings: Sequence[str], train_batches, val_batches=None, model=None,
          save_path: Optional[str] = None, device="cuda", log=None, fused_head: bool = False, **overrides) -> Run:
    """`python -m co3d_3d.train --ginc ... --ginb ...` for one rank (train.py:199-263): parse the gin files and
    bindings, build `get_model()` unless a model is passed, run `fit`."""
    ginlite.parse_config_files_and_bindings(config_files, bindings)
    cfg = TrainConfig(**overrides)
    _select_device(device)
    if model is None:
        model = get_model().to(device)
    run = Run(model, cfg, save_path=save_path, void_label=_bound("PlenoxelScannetDataset.void_label", None), log=log,
              fused_head=fused_head)
    run.fit(train_batches, val_batches)
    return run


def evaluate(load_path: str, val_batches, model=None, save_path: Optional[str] = None, tag: str = "default",
             ignore_label: Optional[int] = None, training_module: str = "SegmentationTraining", device="cuda",
             replace: bool = False, fused_head: bool = False, make_input: Optional[Callable] = None):
    """`python eval.py --ginc ... --load_path <ckpt>` (co3d_3d/eval.py:21-103) for one GPU: build `get_model()` in eval
    mode, load `checkpoint["state_dict"]` (Lightning layout), run one validation pass and write `<save_path>/<tag>.json`
    (a list with one result dict, the shape `Trainer.validate` returns).  An existing json is kept unless `replace`
    (eval.py:42-45).  Checkpoints of pruned networks (`*_mask` / `*_orig` entries) load through identity pruning and
    are made permanent, as eval.py:51-74 does."""
    save_path = save_path if save_path is not None else os.path.dirname(load_path)
    if save_path and not os.path.exists(save_path):
        os.makedirs(save_path, exist_ok=True)
    json_path = os.path.join(save_path, f"{tag}.json")
    if not replace and os.path.isfile(json_path):
        print("====== skip existing experiment =====")
        return None
    _select_device(device)
    if model is None:
        model = get_model().to(device)
    model.eval()
    ckpt = torch.load(load_path, map_location="cpu", weights_only=False)
    pruned = any("_mask" in k for k in ckpt["state_dict"])
    if pruned:
        # eval.py:51-58: identity pruning creates the `*_orig` / `*_mask` entries the checkpoint holds ...
        import torch.nn.utils.prune as torch_prune
        to_prune = get_parameters_to_prune(model)
        for module, name in to_prune:
            torch_prune.identity(module, name)
    load_lightning_state_dict(model, ckpt["state_dict"])
    n_params = count_parameters(model)
    if pruned:
        # ... and eval.py:71-74 makes the masks permanent.  (The reference then switches its convolutions to sparse
        # weight layouts, eval.py:76-79; here the dense tensor-core kernels run on the zero-filled kernels: same result.)
        for module, name in to_prune:
            torch_prune.remove(module, name)
    if ignore_label is None:
        ignore_label = _bound(f"{_bound('get_dataset.dataset_name', 'train')}.ignore_label",
                              _bound("train.ignore_label", -100))
    cfg = TrainConfig(max_steps=0, training_module=training_module, ignore_label=ignore_label)
    run = Run(model, cfg, save_path=save_path, void_label=_bound("PlenoxelScannetDataset.void_label", None),
              make_input=make_input, fused_head=fused_head, evaluate_only=True)
    run.global_step = int(ckpt.get("global_step", 0))
    results = run.validate(val_batches() if callable(val_batches) else val_batches)
    results["val/total_params"], results["val/pruned_params"] = n_params["total"], n_params["pruned"]
    results = {k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in results.items()}
    with open(json_path, "w") as f:
        f.write(json.dumps([results], indent=4))
    return results
