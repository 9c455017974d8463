"""Engine-side row order on the faithful geometry: per level executed / useful (tile, offset) volume before / after,
and the 96->96 convolution kernels on that level's self map in both orders.
usage: sort_diag.py SCENE_SCALE [VOXELS_PER_SCENE] [SCENES]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import me as ME  # noqa: E402
from nerf_downstream_b200 import ops, synth  # noqa: E402

scale = float(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
scenes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
c, f, _ = synth.faithful_room_batch(777, scenes, n, scene_scale=scale, channels=1)
c_d, f_d = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
prec = ops.PRECISIONS["bf16"]


def popc(mask):
    v = mask.to(torch.int64) & 0xFFFFFFFF
    s = torch.zeros_like(v)
    for k in range(32):
        s += (v >> k) & 1
    return int(s.sum())


def timed(fn):
    fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[2]


for sort in (False, True):
    ops.sort_rows = sort
    ops.sort_stats.update(considered=0, reordered=0)
    x = ME.TensorField(coordinates=c_d, features=f_d).sparse()
    mgr = x.coordinate_manager
    k1 = x.coordinate_map_key
    keys = {1: k1}
    keys[2] = mgr.stride(k1, 2)
    for ts in (4, 8):
        keys[ts] = ME.CoordinateMapKey([ts] * 3, "")
    print(f"== sort_rows {sort}: {ops.sort_stats}", flush=True)
    for ts in (1, 2, 4):
        key = keys[ts]
        km = mgr.get_kernel_map(key, key, ME.KernelGenerator(kernel_size=3, stride=1, dilation=1, dimension=3))
        m = mgr.size(key)
        ex, P = popc(km.mask) * 128, km.n_pairs
        line = f"ts{ts}: M={m} P/M={P / m:.2f} executed/useful {ex / max(P, 1):.2f} (executed offsets per row {ex / m:.2f})"
        if m >= 200_000:
            g = torch.Generator().manual_seed(0)
            xb = ops.to_bf16(torch.randn(m, 96, generator=g).to(dev))
            gb = ops.to_bf16(torch.randn(m, 96, generator=g).to(dev))
            w = (torch.randn(27, 96, 96, generator=g) / 51.0).to(dev)
            t = [timed(lambda: ops.conv_fwd_raw(xb, w, None, km, prec)), timed(lambda: ops.conv_dgrad_raw(gb, w, km, prec)),
                 timed(lambda: ops.conv_wgrad_raw(xb, gb, km, 27, 96, 96, prec))]
            line += f"  96->96 fwd {t[0]:.3f} dgrad {t[1]:.3f} wgrad {t[2]:.3f} ms"
        print(line, flush=True)
