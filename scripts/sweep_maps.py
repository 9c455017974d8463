"""BASELINE.json configs[3]: coordinate hash + kernel-map build sweep (100 K - 10 M voxels, stride 1/2/4),
plus the bandwidth-bound row kernels (BN, ReLU, pooling) at the same sizes.  Prints ms and the achieved
ALGORITHMIC GB/s (formulas: DESIGN.md §3) against the measured HBM peak."""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
peaks = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
HBM = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def line(name, m, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:34s} M={m:9d}  {ms:8.3f} ms  {gbs:8.0f} GB/s alg  {100 * gbs / HBM:5.1f}% of measured HBM peak", flush=True)


sizes = [int(s) for s in sys.argv[1:]] or [100_000, 300_000, 1_000_000, 3_000_000, 10_000_000]
for n in sizes:
    c, _, _ = synth.room_batch(777, 1, n, channels=1, shuffle=bool(os.environ.get("SWEEP_SHUFFLE")))   # raster order unless asked
    cg = torch.from_numpy(c).to(dev)
    ms = timed(lambda: ops.coords_insert(cg, L.SRC_FLOAT, (1, 1, 1)))       # includes the host read of M
    cmap, first, inv, cnt = ops.coords_insert(cg, L.SRC_FLOAT, (1, 1, 1))
    m1 = cmap.size
    line("quantize+hash+unique (ts1)", n, ms, 20.0 * n + 36.0 * m1)
    maps = {1: cmap}
    for s in (2, 4):
        src = maps[s // 2]
        ms = timed(lambda: ops.coords_insert(src.coords, L.SRC_STRIDE, (s, s, s)))
        maps[s], _, _, _ = ops.coords_insert(src.coords, L.SRC_STRIDE, (s, s, s))
        line(f"stride map {s // 2}->{s}", src.size, ms, 20.0 * src.size + 36.0 * maps[s].size)
    for ts in (1, 2, 4):
        mp = maps[ts]
        offs = ops.kernel_offsets((3, 3, 3), (ts,) * 3, (1, 1, 1))
        ms = timed(lambda: ops.build_kernel_map(mp, mp, offs))
        km = ops.build_kernel_map(mp, mp, offs)
        line(f"kernel map 3^3 s1 @ts{ts} (P/M={km.n_pairs / max(mp.size, 1):.1f})", mp.size, ms, (16 + 8 * 27 + 4 * 27) * mp.size)
    for ts in (1, 2):
        a, b = maps[ts], maps[ts * 2]
        for ks in (3, 2):
            offs = ops.kernel_offsets((ks,) * 3, (ts,) * 3, (1, 1, 1))
            ms = timed(lambda: ops.build_kernel_map(a, b, offs))
            line(f"kernel map {ks}^3 s2 {ts}->{ts * 2}", b.size, ms, (16 + 12 * ks ** 3) * b.size)
    km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
    ms = timed(lambda: km.__setattr__("_nbr_t", None) or km.nbr_t)
    line("transpose map 3^3", m1, ms, 8.0 * 27 * m1)
    ms = timed(lambda: ops.tile_mask(km.nbr, m1, 27))
    line("tile mask", m1, ms, 4.0 * 27 * m1)
    if n <= 3_000_000:
        for C in (32, 96):
            x = torch.randn(m1, C, device=dev)
            g = torch.randn(m1, C, device=dev)
            gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
            xr = x.clone().requires_grad_()
            ms = timed(lambda: ops.BatchNormFn.apply(x, gamma, beta, rm, rv, True, 0.1, 1e-5, True, None))
            line(f"BN+ReLU fwd C={C}", m1, ms, 12.0 * m1 * C)
            y = ops.BatchNormFn.apply(xr, gamma, beta, rm, rv, True, 0.1, 1e-5, True, None)
            ms = timed(lambda: torch.autograd.grad(y, xr, g, retain_graph=True))
            line(f"BN+ReLU bwd C={C}", m1, ms, 24.0 * m1 * C)
            ms = timed(lambda: ops.ReLUFn.apply(x))
            line(f"ReLU fwd C={C}", m1, ms, 8.0 * m1 * C)
    del maps, km
    torch.cuda.empty_cache()
