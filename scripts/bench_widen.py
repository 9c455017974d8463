"""Device timings of the widened rows at the headline size (1 M voxels / points, 20 classes): the fused segmentation
head against the three passes it replaces, instance norm against the HBM roofline, trilinear interpolation.
CUDA events on the current stream, 5 warm-ups, 20 timed repetitions, L2 flushed between repetitions.
    python scripts/bench_widen.py  ->  one JSON line per kernel group"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import ops, pipeline  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = {}
try:
    peaks = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())
except Exception:
    pass


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    n, C = 1_000_000, 20
    g = torch.Generator(device=dev).manual_seed(0)
    if only in ("", "head"):
        bench_head(n, C, g)
    if only in ("", "inst"):
        bench_inst(n, g)
    if only in ("", "interp"):
        bench_interp(n, g)
    print(json.dumps({"launches": L.launch_count()}))


def bench_head(n, C, g):
    logits = torch.randn(n, C, device=dev, generator=g, requires_grad=True)
    inverse = torch.arange(n, dtype=torch.int32, device=dev)
    target = torch.randint(0, C, (n,), device=dev, generator=g)
    target[::10] = -255
    counts = torch.zeros(3, C, dtype=torch.int64, device=dev)

    def three_pass():
        pts = ops.GatherRowsFn.apply(logits, inverse)
        loss = ops.cross_entropy(pts, target, -255)
        pipeline.seg_counts(pts.detach(), target, -255, out=counts)
        loss.backward()
        logits.grad = None

    def fused():
        loss = ops.seg_head(logits, inverse, target, -255, None, counts)
        loss.backward()
        logits.grad = None

    t3, t1 = timed(three_pass), timed(fused)
    alg = (4.0 * C + 4 + 8) * n + 4.0 * C * n * 2          # fwd: logits + inverse + target; grad_raw write; bwd read+write
    print(json.dumps({"kernel": "seg_head fwd+bwd", "n": n, "C": C, "three_pass_ms": round(t3, 4), "fused_ms": round(t1, 4),
                      "speedup": round(t3 / t1, 2), "fused_GBps_algorithmic": round(alg / t1 / 1e6, 1)}))



def bench_inst(n, g):
    for Cn in (32, 96):
        x = torch.randn(n, Cn, device=dev, generator=g, requires_grad=True)
        coords = torch.zeros(n, 4, dtype=torch.int32, device=dev)
        coords[n // 2:, 0] = 1
        gam = torch.ones(1, Cn, device=dev, requires_grad=True)
        bet = torch.zeros(1, Cn, device=dev, requires_grad=True)
        gy = torch.randn(n, Cn, device=dev, generator=g)
        tf = timed(lambda: ops.InstanceNormFn.apply(x.detach(), coords, 2, gam.detach(), bet.detach(), 1e-8))

        def fb():
            y = ops.InstanceNormFn.apply(x, coords, 2, gam, bet, 1e-8)
            y.backward(gy)
            x.grad = None
        tfb = timed(fb)
        bytes_f, bytes_b = 12.0 * n * Cn, 20.0 * n * Cn
        print(json.dumps({"kernel": "instance_norm", "m": n, "C": Cn, "fwd_ms": round(tf, 4), "fwd_GBps": round(bytes_f / tf / 1e6, 1),
                          "bwd_ms": round(tfb - tf, 4), "bwd_GBps": round(bytes_b / max(tfb - tf, 1e-6) / 1e6, 1),
                          "hbm_peak_GBps": peaks.get("hbm_gbs")}))



def bench_interp(n, g):
    # interpolation of a 1 M-voxel map at 1 M points, 32 channels
    from nerf_downstream_b200 import synth
    c, _, _ = synth.room_batch(777, 1, n, channels=1)
    cmap, _, _, _ = ops.coords_insert(torch.from_numpy(c).to(dev), L.SRC_FLOAT, (1, 1, 1))
    q = torch.from_numpy(c).to(dev)
    feats = torch.randn(cmap.size, 32, device=dev, generator=g)
    tm = timed(lambda: ops.interp_map(cmap, q))
    idx, w = ops.interp_map(cmap, q)
    tg = timed(lambda: ops.InterpolateFn.apply(feats, idx, w))
    hit = float((idx >= 0).float().mean())
    print(json.dumps({"kernel": "interpolate", "m": cmap.size, "n": q.shape[0], "C": 32, "map_ms": round(tm, 4),
                      "gather_ms": round(tg, 4), "corner_hit_rate": round(hit, 3),
                      "gather_GBps_algorithmic": round((4.0 * 32 * (cmap.size + q.shape[0]) + 64.0 * q.shape[0]) / tg / 1e6, 1)}))


if __name__ == "__main__":
    main()
