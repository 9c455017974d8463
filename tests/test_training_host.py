"""Trainer + config parity, host side — runs without a GPU.

Pinned against the reference itself where it can be imported here: `schedules_ref.json` comes from the reference's
`optim.py` running unchanged on `ginlite` (tests/golden/make_schedules.py), `metrics_ref.json` from its
`utils/__init__.py` (tests/golden/make_metrics.py).  The fused SGD / counting kernels are CUDA-only, so — as in
test_trainer_gloo.py — they are replaced by their torch restatements here; what is under test is the host logic.
"""
import json
import math
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from nerf_downstream_b200 import ginlite, schedules, training

GOLDEN = Path(__file__).resolve().parent / "golden"
REF_CONFIGS = Path("/root/reference/co3d_3d/configs")


@pytest.fixture(autouse=True)
def _clean_gin():
    ginlite.clear_config()
    yield
    ginlite.clear_config()


# ---- ginlite ----------------------------------------------------------------------------------------------------
CONFIG = '''
# Dataset
get_dataset.dataset_name = "PlenoxelScannetDataset"   # trailing comment
PlenoxelScannetDataset.train_transformations = [
    "RandomRotation",
    "RandomCrop",   # comment inside a list
    "Elastic#Distortion",
]
PlenoxelScannetDataset.void_label = None
ElasticDistortion.distortion_params = [(4, 16)]
get_model.name = 'Res16UNet34C'
get_model.in_channel = 27
get_model.out_channel = 20
train.max_steps = 60000
train.scheduler_interval = 'step'
train.lr = 1e-1
train.ignore_label = -255
train.loggers = ["csv", "wandb"]
SGD.momentum = 0.9
'''


def test_ginlite_parses_bindings():
    ginlite.parse_config(CONFIG)
    q = ginlite.query_parameter
    assert q("get_dataset.dataset_name") == "PlenoxelScannetDataset"
    assert q("PlenoxelScannetDataset.train_transformations") == ["RandomRotation", "RandomCrop", "Elastic#Distortion"]
    assert q("PlenoxelScannetDataset.void_label") is None
    assert q("ElasticDistortion.distortion_params") == [(4, 16)]
    assert q("train.lr") == 0.1 and q("train.ignore_label") == -255 and q("train.max_steps") == 60000
    # --ginb bindings override files, later wins (train.py:245-252)
    ginlite.parse_config_files_and_bindings([], ["train.gpus=8", "train.lr = 0.05"])
    assert q("train.gpus") == 8 and q("train.lr") == 0.05
    with pytest.raises(ValueError, match="no bound value"):
        q("train.nonexistent")
    assert "train.lr = 0.05" in ginlite.config_str()


@pytest.mark.parametrize("bad", ["train.lr", "train.lr = foo(", "include 'x.gin'", "a/b.c = 1", "x = 3",
                                 "train.lr = [1, 2"])
def test_ginlite_rejects(bad):
    with pytest.raises(ginlite.GinError):
        ginlite.parse_config(bad)


def test_ginlite_configurable_injection():
    @ginlite.configurable
    def get_model(name, in_channel, out_channel, sparse=None):
        return name, in_channel, out_channel, sparse

    @ginlite.configurable()
    class Opt:
        def __init__(self, params, lr=1.0, momentum=0.0):
            self.params, self.lr, self.momentum = params, lr, momentum

    @ginlite.configurable
    class SubOpt(Opt):          # the reference's `class SGD(optim.SGD): pass` pattern (optim.py:12-14)
        pass

    ginlite.parse_config("get_model.name='ResNet14'\nget_model.in_channel=27\nget_model.out_channel=51\n"
                         "SubOpt.momentum = 0.9\nOpt.lr = 0.5")
    assert get_model() == ("ResNet14", 27, 51, None)
    assert get_model(out_channel=7, sparse=[0]) == ("ResNet14", 27, 7, [0])      # the caller wins
    assert get_model("X") == ("X", 27, 51, None)
    o = SubOpt([1], lr=0.1)
    assert (o.lr, o.momentum) == (0.1, 0.9)
    assert Opt([1]).lr == 0.5 and Opt([1]).momentum == 0.0
    ginlite.bind_parameter("get_model.bogus", 1)
    with pytest.raises(ginlite.GinError, match="no parameter 'bogus'"):
        get_model()


@pytest.mark.skipif(not REF_CONFIGS.exists(), reason="reference tree not mounted")
def test_ginlite_reads_every_reference_config():
    files = sorted(REF_CONFIGS.glob("*.gin"))
    assert len(files) >= 25
    for f in files:
        ginlite.clear_config()
        ginlite.parse_config_file(str(f))
        assert ginlite.config_dict(), f
    ginlite.clear_config()
    ginlite.parse_config_files_and_bindings([str(REF_CONFIGS / "scannet_plenoxel.gin"), str(REF_CONFIGS / "resunet34.gin")],
                                            ["train.gpus=8"])
    cfg = training.TrainConfig()
    assert (cfg.max_steps, cfg.lr, cfg.weight_decay, cfg.momentum) == (60000, 0.1, 1e-4, 0.9)
    assert (cfg.scheduler_name, cfg.ignore_label, cfg.val_every_n_steps, cfg.batch_size) == ("CosineAnnealingLR", -255, 400, 8)
    assert ginlite.query_parameter("get_model.name") == "Res16UNet34C" and cfg.gpus == 8
    assert abs(cfg.schedule().lr(30000) - 0.05) < 1e-12


# ---- schedules ----------------------------------------------------------------------------------------------------
def _schedule_for(case):
    ginlite.clear_config()
    ginlite.parse_config_files_and_bindings([], ["train.scheduler_interval='step'", "train.max_epochs=-1",
                                                 *case["bindings"]])
    q = ginlite.query_parameter
    kw = {k.split(".")[1]: v for k, v in ginlite.config_dict().items() if k.startswith(case["scheduler"] + ".")}
    return schedules.get_schedule(case["scheduler"], q("train.lr"), q("train.max_steps"), case["warmup_steps"],
                                  q("train.max_epochs"), q("train.scheduler_interval"), **kw)


def test_schedules_match_reference_optim():
    golden = json.loads((GOLDEN / "schedules_ref.json").read_text())["cases"]
    assert len(golden) >= 14
    for name, case in golden.items():
        s = _schedule_for(case)
        for t, want in enumerate(case["lr"]):
            assert abs(s.lr(t) - want) <= 1e-12 + 1e-9 * abs(want), (name, t, s.lr(t), want)
        if case["momentum"] is not None:
            for t, want in enumerate(case["momentum"]):
                assert abs(s.momentum(t) - want) <= 1e-12, (name, t)
        else:
            assert s.momentum(0) is None
        if "object at" not in case["repr"] and case["warmup_steps"] <= 0 and case["scheduler"] != "CyclicLR":
            assert repr(s) == case["repr"]


def test_schedules_match_torch_directly():
    p = torch.nn.Parameter(torch.zeros(1))

    def run(make, steps):
        opt = torch.optim.SGD([p], lr=0.1, momentum=0.9)
        sch = make(opt)
        out = []
        for _ in range(steps):
            out.append(opt.param_groups[0]["lr"])
            opt.step()
            sch.step()
        return out
    L = torch.optim.lr_scheduler
    for got, want in [
        (schedules.cosine(0.1, 60000), run(lambda o: L.CosineAnnealingLR(o, 60000), 2000)),
        (schedules.cosine(0.1, 50, 0.001), run(lambda o: L.CosineAnnealingLR(o, 50, 0.001), 51)),
        (schedules.step(0.1, 7, 0.3), run(lambda o: L.StepLR(o, 7, 0.3), 40)),
        (schedules.multistep(0.1), run(lambda o: L.MultiStepLR(o, [20000, 40000], 0.1), 10)),
        (schedules.exponential(0.1), run(lambda o: L.ExponentialLR(o, 0.99), 100)),
    ]:
        for t, w in enumerate(want):
            assert abs(got.lr(t) - w) <= 1e-9 * abs(w) + 1e-15, (got, t)


def test_schedule_errors_follow_reference():
    with pytest.raises(ValueError, match="not recognized"):
        schedules.get_schedule("LinearLR", 0.1, 100)
    with pytest.raises(ValueError, match="Invalid mode"):                       # the reference's default mode is a
        schedules.get_schedule("CyclicLR", 0.1, 100, base_lr=0.01)              # typo ("trianglular", optim.py:148)
    with pytest.raises(KeyError):
        schedules.get_schedule("PolyLR", 0.1, 100)                              # poly_exp has no default (optim.py:192)
    assert schedules.get_schedule("None", 0.1, 100) is None and schedules.get_schedule("none", 0.1, 100) is None
    # old helpers of trainer.py now agree with the reference's PolyFunctor
    from nerf_downstream_b200 import trainer
    assert trainer.poly_lr(0.1, 5, 100) == schedules.poly(0.1, 100, 0.9).lr(5)
    assert trainer.cosine_lr(0.1, 5, 100) == schedules.cosine(0.1, 100).lr(5)


# ---- metrics -----------------------------------------------------------------------------------------------------
def test_metrics_match_reference_utils():
    golden = json.loads((GOLDEN / "metrics_ref.json").read_text())
    for name, c in golden.items():
        t, p = torch.tensor(c["target"]), torch.tensor(c["pred"])
        oa = training.precision_at_one(p, t, c["ignore"])
        if c["precision_at_one"] is None:
            assert math.isnan(oa)
        else:
            assert abs(oa - c["precision_at_one"]) < 1e-4, name
        hist = training.fast_hist(p, t, c["C"])
        assert hist.tolist() == c["hist"], name
        iu = training.per_class_iu(hist)
        for a, b in zip(iu, c["per_class_iu"]):
            assert (b is None and np.isnan(a)) or abs(a - b) < 1e-12
        # eval_metrics on one-hot logits reproduces both numbers
        logits = torch.nn.functional.one_hot(p, c["C"]).float()
        m = training.eval_metrics(logits, t, c["C"], c["ignore"])
        assert (math.isnan(m["OA"]) and c["precision_at_one"] is None) or abs(m["OA"] - c["precision_at_one"]) < 1e-4
        assert abs(m["mIoU"] - float((np.array(iu) * 100).mean())) < 1e-9 or np.isnan(m["mIoU"])


def test_topk_accuracy():
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(64, 51, generator=g)
    labels = torch.randint(0, 51, (64,), generator=g)
    a1, a5 = training.accuracy_topk(logits, labels, (1, 5))
    rank = (logits > logits.gather(1, labels.view(-1, 1))).sum(1)        # number of strictly larger logits
    assert abs(a1 - (rank < 1).float().mean().item() * 100) < 1e-4
    assert abs(a5 - (rank < 5).float().mean().item() * 100) < 1e-4
    m1, m5 = training.AccuracyMeter(51, 1), training.AccuracyMeter(51, 5)
    for lo in range(0, 64, 16):
        m1(logits[lo:lo + 16], labels[lo:lo + 16])
        m5(logits[lo:lo + 16], labels[lo:lo + 16])
    assert abs(m1.compute() * 100 - a1) < 1e-4 and abs(m5.compute() * 100 - a5) < 1e-4
    m1.reset()
    assert math.isnan(m1.compute())


def test_seg_loss_weights():
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(200, 6, generator=g)
    labels = torch.randint(0, 6, (200,), generator=g)
    labels[::7] = -255
    plain = training.SegLoss(-255, 6)
    assert not plain.weighted and torch.equal(plain.weight, torch.ones(6))
    want = torch.nn.functional.cross_entropy(logits, labels, ignore_index=-255)
    assert torch.allclose(plain(logits, {"labels": labels.int()}), want)
    void = training.SegLoss(-255, 6, void_weight=0.1)
    w = torch.ones(6)
    w[-1] = 0.1
    assert void.weighted and torch.equal(void.weight, w)
    assert torch.allclose(void(logits, {"labels": labels}),
                          torch.nn.functional.cross_entropy(logits, labels, weight=w, ignore_index=-255))
    assert not training.SegLoss(-255, 6, void_weight=0.0).weighted and "weight" not in plain.state_dict()


# ---- checkpoints + the loop ------------------------------------------------------------------------------------
def _cpu_kernels(monkeypatch):
    """torch restatements of the two CUDA-only kernels the loop calls (spc_sgd_step, spc_seg_metrics)."""
    from nerf_downstream_b200 import ops, pipeline

    def sgd_cpu(param, grad, buf, lr, momentum, weight_decay, grad_scale, first_step):
        d = grad * grad_scale + weight_decay * param
        buf.copy_(d if first_step else momentum * buf + d)
        param.sub_(lr * buf)

    def seg_counts_cpu(logits, target, ignore_label, out=None):
        C = logits.shape[1]
        out = torch.zeros((3, C), dtype=torch.int64) if out is None else out
        keep = target != ignore_label
        pred, tgt = logits.argmax(1)[keep], target[keep]
        for c in range(C):
            out[0, c] += (tgt == c).sum()
            out[1, c] += ((tgt == c) & (pred == tgt)).sum()
            out[2, c] += (pred == c).sum()
        return out

    monkeypatch.setattr(ops, "sgd_step", sgd_cpu)
    monkeypatch.setattr(pipeline, "seg_counts", seg_counts_cpu)


class TinyNet(torch.nn.Module):
    def __init__(self, cin=5, cout=4):
        super().__init__()
        self.conv = torch.nn.Linear(cin, 8)
        self.bn = torch.nn.BatchNorm1d(8)
        self.final = torch.nn.Linear(8, cout)

    def forward(self, x):
        return self.final(torch.relu(self.bn(self.conv(x))))


def _batches(seed, n_batches, n=32, cin=5, cout=4, ignore=None):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        x = torch.randn(n, cin, generator=g)
        y = (x[:, :cout].argmax(1)).long()
        if ignore is not None:
            y[::5] = ignore
        out.append({"coordinates": torch.zeros(n, 4), "features": x, "labels": y})
    return out


def test_optimizer_state_is_torch_sgd_layout(monkeypatch):
    """Our arena state loads into a real torch.optim.SGD (what a Lightning checkpoint of the reference holds) and
    back, and both continue identically."""
    _cpu_kernels(monkeypatch)
    from nerf_downstream_b200 import trainer as T
    torch.manual_seed(0)
    a, b = TinyNet(), TinyNet()
    b.load_state_dict(a.state_dict())
    tr = T.DataParallelTrainer(a, lr=0.05, momentum=0.9, weight_decay=1e-4)
    opt = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    data = _batches(1, 6)

    def torch_step(batch):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(b(batch["features"]), batch["labels"]).backward()
        opt.step()
    for batch in data[:3]:
        tr.backward_and_step(torch.nn.functional.cross_entropy(a(batch["features"]), batch["labels"]))
        torch_step(batch)
    mine, theirs = training.optimizer_state(tr), opt.state_dict()
    assert mine["param_groups"][0]["params"] == theirs["param_groups"][0]["params"]
    for i in theirs["state"]:
        assert torch.allclose(mine["state"][i]["momentum_buffer"], theirs["state"][i]["momentum_buffer"], atol=1e-7)
    # cross-load: torch's state into a fresh arena trainer, ours into a fresh torch optimiser
    c, d = TinyNet(), TinyNet()
    c.load_state_dict(a.state_dict())
    d.load_state_dict(a.state_dict())
    tr2 = T.DataParallelTrainer(c, lr=0.01)
    training.load_optimizer_state(tr2, theirs)
    assert (tr2.lr, tr2.momentum, tr2.weight_decay, tr2.steps) == (0.05, 0.9, 1e-4, 1)
    opt2 = torch.optim.SGD(d.parameters(), lr=0.01)
    opt2.load_state_dict(mine)
    for batch in data[3:]:
        tr2.backward_and_step(torch.nn.functional.cross_entropy(c(batch["features"]), batch["labels"]))
        opt2.zero_grad()
        torch.nn.functional.cross_entropy(d(batch["features"]), batch["labels"]).backward()
        opt2.step()
        tr.backward_and_step(torch.nn.functional.cross_entropy(a(batch["features"]), batch["labels"]))
    for pa, pc, pd in zip(a.parameters(), c.parameters(), d.parameters()):
        assert torch.allclose(pa, pc, atol=1e-6) and torch.allclose(pa, pd, atol=1e-6)


def test_lightning_checkpoint_layout(tmp_path, monkeypatch):
    _cpu_kernels(monkeypatch)
    from nerf_downstream_b200 import trainer as T
    torch.manual_seed(1)
    net = TinyNet()
    tr = T.DataParallelTrainer(net, lr=0.1, momentum=0.9)
    path = str(tmp_path / "last.ckpt")
    training.save_checkpoint(path, tr, global_step=17, epoch=2)
    ckpt = torch.load(path, weights_only=False)
    assert {"state_dict", "optimizer_states", "lr_schedulers", "global_step", "epoch"} <= set(ckpt)
    assert all(k.startswith("model.") for k in ckpt["state_dict"])                       # eval.py:47-67
    assert set(k[6:] for k in ckpt["state_dict"]) == set(net.state_dict())
    # a Lightning module's own entries (e.g. the criterion's buffers) are ignored on load
    ckpt["state_dict"]["criterion.weight"] = torch.ones(4)
    other = TinyNet()
    training.load_lightning_state_dict(other, ckpt["state_dict"])
    for k, v in net.state_dict().items():
        assert torch.equal(other.state_dict()[k], v)
    with pytest.raises(KeyError):
        training.load_lightning_state_dict(other, {"conv.weight": torch.zeros(1)})
    # self-supervised transfer (lightning_module_base.py:62-72)
    ssl = {"model.encoder.conv.weight": torch.full((8, 5), 2.0), "model.predictor.0.weight": torch.zeros(1),
           "model.encoder.final.weight": torch.zeros(4, 8), "model.encoder.bn.weight": torch.full((8,), 3.0)}
    conv = training.convert_self_supervised_checkpoint(ssl)
    assert list(conv) == ["model.conv.weight", "model.bn.weight"]
    res = training.load_lightning_state_dict(other, conv, strict=False)
    assert "final.weight" in res.missing_keys and float(other.conv.weight.detach()[0, 0]) == 2.0


def test_fit_classification(tmp_path, monkeypatch):
    _cpu_kernels(monkeypatch)
    ginlite.parse_config("""
train.training_module = "ClassificationTraining"
train.max_steps = 40
train.warmup_steps = 5
train.scheduler_name = "CosineAnnealingLR"
train.lr = 0.2
train.weight_decay = 0.0
train.val_every_n_steps = 15
train.log_every_n_steps = 10
train.monitor_metric = "val/acc1"
SGD.momentum = 0.9
get_model.out_channel = 4
""")
    torch.manual_seed(2)
    net = TinyNet()
    cfg = training.TrainConfig()
    assert cfg.total_steps == 45 and cfg.momentum == 0.9
    logs = []
    run = training.Run(net, cfg, save_path=str(tmp_path), make_input=lambda b: b["features"], log=logs.append)
    assert run.trainer.lr == 0.0                                                     # warm-up starts from 0
    train_data, val_data = _batches(10, 8), _batches(11, 3)
    last = run.fit(lambda: train_data, lambda: val_data)
    assert run.global_step == 45                                                     # max_steps + warmup (train.py:175)
    sched = cfg.schedule()
    tr_logs = [d for d in logs if "train/loss" in d]
    assert [d["global_step"] for d in tr_logs] == [10, 20, 30, 40]
    for d in tr_logs:
        assert abs(d["train/lr"] - sched.lr(d["global_step"])) < 1e-12
        assert 0 <= d["train/acc1"] <= d["train/acc5"] <= 100
    assert abs(run.trainer.lr - sched.lr(45)) < 1e-12
    val_logs = [d for d in logs if "val/loss" in d]
    assert [d["global_step"] for d in val_logs] == [15, 30, 45]
    assert last["val/acc1"] > 0.6 and last["val/acc5"] == 1.0                         # 4 classes: top-5 is everything
    assert val_logs[-1]["val/loss"] < val_logs[0]["val/loss"]
    assert (tmp_path / "last.ckpt").exists() and (tmp_path / "best.ckpt").exists()
    # resume: weights, momentum, step counter and learning rate continue
    ginlite.bind_parameter("train.checkpoint_path", str(tmp_path / "last.ckpt"))
    ginlite.bind_parameter("train.resume_training", True)
    net2 = TinyNet()
    run2 = training.Run(net2, training.TrainConfig(), make_input=lambda b: b["features"])
    assert run2.global_step == 45 and abs(run2.trainer.lr - sched.lr(45)) < 1e-12
    for k, v in net.state_dict().items():
        assert torch.equal(net2.state_dict()[k], v), k
    assert torch.allclose(training.optimizer_state(run2.trainer)["state"][0]["momentum_buffer"],
                          training.optimizer_state(run.trainer)["state"][0]["momentum_buffer"])


def test_fit_segmentation_step_and_validation(monkeypatch):
    _cpu_kernels(monkeypatch)
    ginlite.parse_config("""
train.max_steps = 12
train.scheduler_name = "PolyLR"
PolyLR.poly_exp = 0.9
train.lr = 0.1
train.ignore_label = -255
train.val_every_n_steps = 6
train.log_every_n_steps = 4
train.void_weight = 0.5
SGD.momentum = 0.9
get_model.out_channel = 4
""")
    torch.manual_seed(5)
    logs = []
    run = training.Run(TinyNet(), training.TrainConfig(), void_label=3, make_input=lambda b: b["features"],
                       log=logs.append)
    assert run.segmentation and run.criterion.weighted and float(run.criterion.weight[-1]) == 0.5
    data, val = _batches(20, 4, ignore=-255), _batches(21, 2, ignore=-255)
    last = run.fit(lambda: data, lambda: val)
    assert run.global_step == 12 and abs(run.trainer.lr - 0.1 * (1 - 12 / 13) ** 0.9) < 1e-12
    t = [d for d in logs if "train/loss" in d]
    assert [d["global_step"] for d in t] == [4, 8] and abs(t[0]["train/ignore_ratio"] - 21.875) < 1e-4
    assert {"train/OA", "train/mIoU", "train/lr"} <= set(t[0])
    assert {"val/mIoU", "val/mAcc", "val/OA", "val/best_mIoU", "val/loss"} <= set(last)
    # void_label set -> the last class is left out of the means (metrics.py:52-57)
    miou, ious, macc, accs = run.iou_meter.compute()
    assert abs(float(miou) - float(ious[:-1].mean())) < 1e-7 and abs(last["val/mIoU"] - float(miou) * 100) < 1e-4


def test_train_config_errors():
    with pytest.raises(ginlite.GinError, match="max_steps"):
        training.TrainConfig()
    with pytest.raises(TypeError, match="unknown"):
        training.TrainConfig(max_steps=1, bogus=2)
    ginlite.parse_config("get_model.name='PointNet'\nget_model.in_channel=3\nget_model.out_channel=4")
    with pytest.raises(KeyError, match="not built"):
        training.get_model()
    with pytest.raises(AssertionError, match="not in"):
        training.Run(TinyNet(), training.TrainConfig(max_steps=1, training_module="Foo"), num_labels=4)


def test_fused_head_run_equals_three_pass_run(monkeypatch):
    """`fused_head=True` (one kernel for slice + SegLoss + counts) drives the loop to the same losses, metrics and
    weights as the three-pass form.  The kernel itself is replaced by its torch restatement here (GPU parity:
    tests/test_gpu_widen.py); under test are the loop, the counts -> OA / mIoU conversion and the meters."""
    _cpu_kernels(monkeypatch)
    from types import SimpleNamespace

    from nerf_downstream_b200 import pipeline

    class SparseTiny(TinyNet):
        def forward_sparse(self, x):
            return SimpleNamespace(F=TinyNet.forward(self, x))

    def head_cpu(out, field, labels, ignore_index=-100, weight=None, counts=None):
        if counts is not None:
            pipeline.seg_counts(out.F.detach(), labels, ignore_index, out=counts)
        return torch.nn.functional.cross_entropy(out.F, labels, weight=weight, ignore_index=ignore_index)

    monkeypatch.setattr(pipeline, "seg_head_loss", head_cpu)
    ginlite.parse_config("train.max_steps = 9\ntrain.scheduler_name = 'PolyLR'\nPolyLR.poly_exp = 0.9\ntrain.lr = 0.1\n"
                         "train.ignore_label = -255\ntrain.val_every_n_steps = 3\ntrain.log_every_n_steps = 2\n"
                         "train.void_weight = 0.5\nSGD.momentum = 0.9\nget_model.out_channel = 4")
    data, val = _batches(30, 3, ignore=-255), _batches(31, 2, ignore=-255)
    results = []
    for fused in (False, True):
        torch.manual_seed(7)
        logs = []
        run = training.Run(SparseTiny(), training.TrainConfig(), make_input=lambda b: b["features"], log=logs.append,
                           fused_head=fused)
        last = run.fit(lambda: data, lambda: val)
        results.append((logs, last, [p.detach().clone() for p in run.model.parameters()]))
    (la, va, pa), (lb, vb, pb) = results
    assert len(la) == len(lb) and len(la) >= 6
    for a, b in zip(la, lb):
        assert a.keys() == b.keys()
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-5 * (1 + abs(a[k])), (k, a[k], b[k])
    for k in va:
        assert abs(va[k] - vb[k]) <= 1e-5 * (1 + abs(va[k])), k
    for x, y in zip(pa, pb):
        assert torch.allclose(x, y, atol=1e-6)
    with pytest.raises(ValueError, match="forward_sparse"):
        training.Run(TinyNet(), training.TrainConfig(), make_input=lambda b: b["features"], fused_head=True)


def test_metrics_from_counts_equals_eval_metrics():
    golden = json.loads((GOLDEN / "metrics_ref.json").read_text())
    for name, c in golden.items():
        counts = torch.tensor([c["seen"], c["correct"], c["positive"]]).long()
        m = training.metrics_from_counts(counts)
        if c["precision_at_one"] is None:
            assert math.isnan(m["OA"])
            continue
        assert abs(m["OA"] - c["precision_at_one"]) < 1e-4
        want = float(np.mean([0.0 if v is None else v for v in c["per_class_iu"]]) * 100)
        assert abs(m["mIoU"] - want) < 1e-9, name


def test_evaluate_from_lightning_checkpoint(tmp_path, monkeypatch):
    """eval.py: a checkpoint written by the loop evaluates in a fresh model to the numbers of the run's own last
    validation, the result lands in <tag>.json, and an existing json is kept unless replace=True."""
    _cpu_kernels(monkeypatch)
    ginlite.parse_config("train.max_steps = 6\ntrain.scheduler_name = 'none'\ntrain.lr = 0.05\ntrain.ignore_label = -255\n"
                         "train.val_every_n_steps = 6\nSGD.momentum = 0.9\nget_model.out_channel = 4\n"
                         "get_dataset.dataset_name = 'PlenoxelScannetDataset'\nPlenoxelScannetDataset.ignore_label = -255")
    torch.manual_seed(3)
    data, val = _batches(40, 3, ignore=-255), _batches(41, 2, ignore=-255)
    run = training.Run(TinyNet(), training.TrainConfig(), save_path=str(tmp_path), make_input=lambda b: b["features"])
    assert run.schedule is None and run.trainer.lr == 0.05
    last = run.fit(lambda: data, lambda: val)
    res = training.evaluate(str(tmp_path / "last.ckpt"), lambda: val, model=TinyNet(), tag="t1",
                            make_input=lambda b: b["features"])
    for k in ("val/mIoU", "val/mAcc", "val/OA", "val/loss"):
        assert abs(res[k] - last[k]) < 1e-6, k
    stored = json.loads((tmp_path / "t1.json").read_text())
    assert isinstance(stored, list) and abs(stored[0]["val/mIoU"] - res["val/mIoU"]) < 1e-9
    assert training.evaluate(str(tmp_path / "last.ckpt"), lambda: val, model=TinyNet(), tag="t1",
                             make_input=lambda b: b["features"]) is None                       # kept
    assert training.evaluate(str(tmp_path / "last.ckpt"), val, model=TinyNet(), tag="t1", replace=True,
                             make_input=lambda b: b["features"]) is not None


def test_exception_safe_run_survives_failing_batches(monkeypatch):
    """ExceptionSafeSegmentationTraining (segmentation_training.py:233-326): a step that raises RuntimeError (how the
    library reports e.g. out-of-memory) is counted and skipped, the schedule still advances; a failing validation
    batch is skipped."""
    _cpu_kernels(monkeypatch)

    class Flaky(TinyNet):
        calls = 0

        def forward(self, x):
            Flaky.calls += 1
            if Flaky.calls in (2, 5):
                raise RuntimeError("CUDA out of memory (simulated)")
            return super().forward(x)

    ginlite.parse_config("train.max_steps = 6\ntrain.scheduler_name = 'PolyLR'\nPolyLR.poly_exp = 0.9\ntrain.lr = 0.1\n"
                         "train.ignore_label = -255\nSGD.momentum = 0.9\nget_model.out_channel = 4")
    data = _batches(60, 6, ignore=-255)
    run = training.Run(Flaky(), training.TrainConfig(), make_input=lambda b: b["features"], exception_safe=True)
    before = [p.detach().clone() for p in run.model.parameters()]
    assert run.training_step(data[0]) is not None
    after1 = [p.detach().clone() for p in run.model.parameters()]
    assert run.training_step(data[1]) is None and run.fail_count == 1 and run.global_step == 2
    assert all(torch.equal(a, b) for a, b in zip(after1, run.model.parameters()))      # the failed step changed nothing
    assert any(not torch.equal(a, b) for a, b in zip(before, after1))
    assert abs(run.trainer.lr - 0.1 * (1 - 2 / 7) ** 0.9) < 1e-12                       # the scheduler stepped anyway
    run.fit(lambda: data[2:])
    assert run.global_step == 6 and run.fail_count == 2
    Flaky.calls = 0
    res = run.validate(data[:3])                                                        # batch 2 of 3 fails, is skipped
    assert np.isfinite(res["val/loss"])
    strict = training.Run(Flaky(), training.TrainConfig(), make_input=lambda b: b["features"])
    Flaky.calls = 1
    with pytest.raises(RuntimeError, match="out of memory"):
        strict.training_step(data[0])


def test_collate_mink():
    rng = np.random.default_rng(0)
    samples = [{"coordinates": rng.uniform(0, 9, (n, 3)).astype(np.float32), "features": rng.standard_normal((n, 5)).astype(np.float32),
                "labels": rng.integers(0, 4, n), "instance_ids": np.arange(n), "metadata": {"file": f"s{n}"}} for n in (7, 3, 5)]
    b = training.collate_mink(samples)
    assert b["coordinates"].dtype == torch.float32 and tuple(b["coordinates"].shape) == (15, 4)
    assert b["coordinates"][:, 0].tolist() == [0.0] * 7 + [1.0] * 3 + [2.0] * 5                # batch column first
    assert np.array_equal(b["coordinates"][7:10, 1:].numpy(), samples[1]["coordinates"])
    assert tuple(b["features"].shape) == (15, 5) and b["features"].dtype == torch.float32
    assert b["labels"].tolist() == np.concatenate([s["labels"] for s in samples]).tolist()
    assert b["instance_ids"].shape[0] == 15 and b["metadata"] == [{"file": "s7"}, {"file": "s3"}, {"file": "s5"}]
    assert "dataset" not in b


@pytest.mark.parametrize("name,kw", [("Adam", {}), ("AdamW", {"betas": (0.8, 0.99)}), ("RMSprop", {"momentum": 0.5}),
                                     ("Adagrad", {})])
def test_other_optimizers_of_get_optimizer(monkeypatch, name, kw):
    """`get_optimizer` (optim.py:57-69) knows nine torch optimizers; the ones besides SGD run as torch optimizers on
    the arena-backed parameters with their gin bindings, and the loop / schedule / checkpoint code is the same."""
    _cpu_kernels(monkeypatch)
    binds = "".join(f"{name}.{k} = {v!r}\n" for k, v in kw.items())
    ginlite.parse_config(f"train.max_steps = 5\ntrain.optimizer_name = '{name}'\ntrain.scheduler_name = 'ExponentialLR'\n"
                         f"ExponentialLR.gamma = 0.9\ntrain.lr = 0.01\ntrain.weight_decay = 1e-3\n"
                         f"train.training_module = 'ClassificationTraining'\nget_model.out_channel = 4\n{binds}")
    torch.manual_seed(0)
    a, b = TinyNet(), TinyNet()
    b.load_state_dict(a.state_dict())
    data = _batches(70, 5)
    run = training.Run(a, training.TrainConfig(), make_input=lambda d: d["features"])
    run.fit(lambda: data)
    opt = getattr(torch.optim, name)(b.parameters(), lr=0.01, weight_decay=1e-3, **kw)
    for t, batch in enumerate(data):
        for g in opt.param_groups:
            g["lr"] = 0.01 * 0.9 ** t
        opt.zero_grad()
        torch.nn.functional.cross_entropy(b(batch["features"]), batch["labels"]).backward()
        opt.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, atol=1e-6), name
    st = training.optimizer_state(run.trainer)
    assert set(st) == {"state", "param_groups"} and len(st["state"]) == len(list(a.parameters()))
    with pytest.raises(ValueError, match="not recognized"):
        from nerf_downstream_b200 import trainer as T
        T.DataParallelTrainer(TinyNet(), optimizer_name="LAMB")
