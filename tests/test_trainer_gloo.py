"""World-size-2 check of the data-parallel step on CPU / gloo: flat arenas, bucketed all-reduce from
post-accumulate hooks, 1/world scaling.  The fused SGD kernel is CUDA-only, so the update itself is
replaced by its torch restatement here; the exchange logic is what is under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_downstream_b200 import ops, trainer

    def sgd_cpu(param, grad, buf, lr, momentum, weight_decay, grad_scale, first_step):
        d = grad * grad_scale + weight_decay * param
        buf.copy_(d if first_step else momentum * buf + d)
        param.sub_(lr * buf)

    ops.sgd_step = sgd_cpu
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ref = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ref.load_state_dict(model.state_dict())
    tr = trainer.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4, bucket_mb=1e-5)
    assert tr.world == world and len(tr._buckets) >= 2
    g = torch.Generator().manual_seed(100)
    xs = [torch.randn(4, 7, generator=g) for _ in range(world)]
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        tr.backward_and_step(model(xs[rank]).pow(2).mean())
        # reference: gradient of the mean over ranks of the per-rank losses
        opt.zero_grad()
        (sum(ref(x).pow(2).mean() for x in xs) / world).backward()
        opt.step()
    err = max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
    # use_sync_grad loss re-weighting (segmentation_training.py:112-120): n_r / sum(n) * world
    from nerf_downstream_b200 import training
    scale = training.sync_grad_scale(100 if rank == 0 else 300)
    assert abs(scale - (0.5 if rank == 0 else 1.5)) < 1e-12, scale
    meter = training.AccuracyMeter(4)
    meter.correct, meter.total = 3 + rank, 10
    meter.all_reduce()
    assert (meter.correct, meter.total) == (7, 20)
    q.put((rank, err, [p.data_ptr() for p in model.parameters()][0] == tr.arena.data.data_ptr() + 4 * tr.arena.offsets[-1]))
    dist.destroy_process_group()


def test_data_parallel_step_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, in_arena in res:
        assert err < 1e-6, (rank, err)
        assert in_arena


def test_schedules():
    from nerf_downstream_b200 import trainer
    assert abs(trainer.cosine_lr(0.1, 0, 100) - 0.1) < 1e-12 and abs(trainer.cosine_lr(0.1, 100, 100)) < 1e-12
    assert abs(trainer.cosine_lr(0.1, 50, 100) - 0.05) < 1e-12
    # PolyFunctor: (1 - step / (max_steps + 1)) ** poly_exp  (optim.py:181-188)
    assert abs(trainer.poly_lr(0.1, 0, 100) - 0.1) < 1e-12
    assert abs(trainer.poly_lr(0.1, 99, 100) - 0.1 * (1 - 99 / 101) ** 0.9) < 1e-15


def _fit_worker(rank, world, port, q):
    """Two ranks run `training.Run.fit` (segmentation, use_sync_grad) on batches of DIFFERENT sizes; with the
    point-count re-weighting (segmentation_training.py:112-120) and the gradient mean of the data-parallel step, the
    result must equal single-process SGD on the concatenated batches with the same schedule."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_downstream_b200 import ginlite, ops, schedules, training

    def sgd_cpu(param, grad, buf, lr, momentum, weight_decay, grad_scale, first_step):
        d = grad * grad_scale + weight_decay * param
        buf.copy_(d if first_step else momentum * buf + d)
        param.sub_(lr * buf)

    ops.sgd_step = sgd_cpu
    ginlite.parse_config("train.max_steps = 5\ntrain.scheduler_name = 'PolyLR'\nPolyLR.poly_exp = 0.9\ntrain.lr = 0.1\n"
                         "train.weight_decay = 1e-4\ntrain.use_sync_grad = True\ntrain.log_every_n_steps = 100\n"
                         "SGD.momentum = 0.9\nget_model.out_channel = 3")

    def make():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))

    def batches(r):
        g = torch.Generator().manual_seed(50 + r)
        n = 20 if r == 0 else 44
        out = []
        for _ in range(5):
            x = torch.randn(n, 6, generator=g)
            out.append({"coordinates": torch.zeros(n, 4), "features": x, "labels": x[:, :3].argmax(1)})
        return out

    run = training.Run(make(), training.TrainConfig(), make_input=lambda b: b["features"])
    assert run.trainer.world == world
    run.fit(lambda: batches(rank))
    ref = make()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    sched = schedules.poly(0.1, 5, 0.9)
    all_b = [batches(r) for r in range(world)]
    for step in range(5):
        for gr in opt.param_groups:
            gr["lr"] = sched.lr(step)
        x = torch.cat([all_b[r][step]["features"] for r in range(world)])
        y = torch.cat([all_b[r][step]["labels"] for r in range(world)])
        opt.zero_grad()
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        opt.step()
    err = max((a - b).abs().max().item() for a, b in zip(run.model.parameters(), ref.parameters()))
    # epoch metrics: every rank validates ITS shard, the IoU counts are summed over ranks (metrics.py:17-28), so both
    # ranks report the mIoU of the union — equal to a single process validating all shards
    from nerf_downstream_b200 import pipeline

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
    pipeline.seg_counts = seg_counts_cpu
    mine = run.validate(batches(10 + rank)[:2])
    meter = pipeline.IoUMeter(3, run.cfg.ignore_label)
    run.model.eval()
    with torch.no_grad():
        for r in range(world):
            for b in batches(10 + r)[:2]:
                meter.update(run.model(b["features"]), b["labels"])
    want = float(meter.compute()[0]) * 100
    q.put((rank, err, run.global_step, abs(mine["val/mIoU"] - want)))
    dist.destroy_process_group()


def test_fit_world2_sync_grad_equals_global_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fit_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, steps, miou_err in res:
        assert steps == 5 and err < 1e-5, (rank, err, steps)
        assert miou_err < 1e-4, (rank, miou_err)


class _Patch:
    """monkeypatch stand-in for a spawned worker (tests/host_harness.install only needs setattr)."""

    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _syncbn_worker(rank, world, port, q):
    """MinkowskiSyncBatchNorm over two ranks with different row counts == nn.BatchNorm1d on the concatenated rows:
    outputs, input gradients, local weight / bias gradient sums, running statistics.  (Kernels answered by the host
    harness; the collective is a real gloo all-reduce.)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np

    from nerf_downstream_b200 import me as ME
    from tests import host_harness
    host_harness.install(_Patch(), "fp32")
    g = torch.Generator().manual_seed(7)
    rows = [50, 70]
    xs = [torch.randn(n, 6, generator=g) * 2 + 1 for n in rows]
    ws = [torch.randn(n, 6, generator=g) for n in rows]
    coords = torch.zeros(rows[rank], 4)
    coords[:, 1] = torch.arange(rows[rank])
    bn = ME.MinkowskiSyncBatchNorm(6, momentum=0.1).train()
    with torch.no_grad():
        bn.bn.weight.copy_(torch.linspace(0.5, 1.5, 6))
        bn.bn.bias.copy_(torch.linspace(-0.2, 0.3, 6))
    x = xs[rank].clone().requires_grad_(True)
    field = ME.TensorField(coordinates=coords, features=x)
    out = bn(field)
    (out.F * ws[rank]).sum().backward()

    ref = torch.nn.BatchNorm1d(6, momentum=0.1).train()
    with torch.no_grad():
        ref.weight.copy_(bn.bn.weight)
        ref.bias.copy_(bn.bn.bias)
    xa = torch.cat(xs).clone().requires_grad_(True)
    oa = ref(xa)
    (oa * torch.cat(ws)).sum().backward()
    lo, hi = (0, rows[0]) if rank == 0 else (rows[0], rows[0] + rows[1])
    xhat = (xa.detach() - xa.detach().mean(0)) / torch.sqrt(xa.detach().var(0, unbiased=False) + 1e-5)
    errs = {
        "out": (out.F.detach() - oa.detach()[lo:hi]).abs().max().item(),
        "dx": (x.grad - xa.grad[lo:hi]).abs().max().item(),
        "dgamma": (bn.bn.weight.grad - (torch.cat(ws)[lo:hi] * xhat[lo:hi]).sum(0)).abs().max().item(),
        "dbeta": (bn.bn.bias.grad - torch.cat(ws)[lo:hi].sum(0)).abs().max().item(),
        "run_mean": (bn.bn.running_mean - ref.running_mean).abs().max().item(),
        "run_var": (bn.bn.running_var - ref.running_var).abs().max().item(),
    }
    # eval mode: running statistics, no collective
    bn.eval()
    ev = bn(ME.TensorField(coordinates=coords, features=xs[rank])).F
    ref.eval()
    errs["eval"] = (ev - ref(xs[rank])).abs().max().item()
    q.put((rank, errs, int(bn.bn.num_batches_tracked)))
    dist.destroy_process_group()


def test_sync_batchnorm_world2_equals_batchnorm_on_all_rows():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_syncbn_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs, tracked in res:
        assert tracked == 1
        for k, v in errs.items():
            assert v < 2e-5, (rank, k, v)


def _worker_divergent_seeds(rank, world, port, q):
    """Ranks build their replicas from DIFFERENT seeds (per-rank seeding before model construction is common for
    augmentation): the trainer must broadcast rank 0's parameters and buffers like DistributedDataParallel, and a
    step that fails on ONE rank (exception-safe training) must not leave the others blocked in their all-reduces."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_downstream_b200 import ops, trainer

    def sgd_cpu(param, grad, buf, lr, momentum, weight_decay, grad_scale, first_step):
        d = grad * grad_scale + weight_decay * param
        buf.copy_(d if first_step else momentum * buf + d)
        param.sub_(lr * buf)

    ops.sgd_step = sgd_cpu
    torch.manual_seed(1234 + rank)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    with torch.no_grad():
        model[1].running_mean.add_(float(rank + 1))
    tr = trainer.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4, bucket_mb=1e-5)

    def flat():
        return torch.cat([p.detach().flatten() for p in model.parameters()] + [b.detach().flatten().float() for b in model.buffers()])

    gathered = [torch.zeros_like(flat()) for _ in range(world)]
    dist.all_gather(gathered, flat())
    equal_at_start = all(torch.equal(gathered[0], g) for g in gathered)
    rm_is_rank0 = bool((model[1].running_mean - 1.0).abs().max() < 1e-6)
    g = torch.Generator().manual_seed(100)
    xs = [torch.randn(6, 7, generator=g) for _ in range(world)]
    tr.backward_and_step(model(xs[rank]).pow(2).mean())              # a complete step records the bucket order
    # step 2: rank 1 "fails" after part of its backward ran; rank 0 completes normally
    if rank == 1:
        tr.arena.zero_grad()
        tr._pending = [n for (_, _, n) in tr._buckets]
        tr._handles, tr._launched = [], []
        tr.abort_step()
    else:
        tr.backward_and_step(model(xs[rank]).pow(2).mean())
    gathered = [torch.zeros_like(flat()) for _ in range(world)]
    params = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(params) for _ in range(world)]
    dist.all_gather(gathered, params)
    equal_after_abort = all(torch.allclose(gathered[0], g, atol=1e-7) for g in gathered)
    q.put((rank, equal_at_start, rm_is_rank0, equal_after_abort, tr.steps))
    dist.destroy_process_group()


def test_replicas_start_equal_and_abort_keeps_collectives_matched():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_divergent_seeds, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, equal_at_start, rm_is_rank0, equal_after_abort, steps in res:
        assert equal_at_start, rank
        assert rm_is_rank0, rank
        assert equal_after_abort, rank
        assert steps == 2
