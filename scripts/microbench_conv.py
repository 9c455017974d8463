"""Layer microbenchmark (BASELINE.json configs[4]): one sparse conv fwd / dgrad / wgrad on a
ScanNet-shaped map.  usage: microbench_conv.py VOXELS CIN COUT [KSIZE STRIDE] [--reps N] [--prec tf32|fp32]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("voxels", type=int)
ap.add_argument("cin", type=int)
ap.add_argument("cout", type=int)
ap.add_argument("ksize", type=int, nargs="?", default=3)
ap.add_argument("stride", type=int, nargs="?", default=1)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--prec", default="tf32")
ap.add_argument("--only", default="fwd,dgrad,wgrad")
ap.add_argument("--shuffle", action="store_true")
ap.add_argument("--dbg", default="", help="idx=val,... passed to spc_debug_set")
ap.add_argument("--sort-window", type=int, default=0,
                help="re-insert the voxels sorted by their 27-bit neighbour mask inside windows of W rows (0 = raster)")
args = ap.parse_args()
dev = torch.device("cuda:0")
for kv in filter(None, args.dbg.split(",")):
    i, v = kv.split("=")
    L.load().spc_debug_set(int(i), int(v))
prec = ops.PRECISIONS[args.prec]
c, _, _ = synth.room_batch(777, 1, args.voxels, channels=1, shuffle=args.shuffle)
cmap, _, _, _ = ops.coords_insert(torch.from_numpy(c).to(dev), L.SRC_FLOAT, (1, 1, 1))
if args.sort_window > 0:
    # what an engine-side row order would do to the (tile, offset) skipping: same voxels, rows grouped by mask
    km0 = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
    bits = (km0.nbr >= 0).to(torch.int64) << torch.arange(27, device=dev).view(27, 1)
    key = (torch.arange(cmap.size, device=dev) // args.sort_window << 27) | bits.sum(0)
    perm = torch.sort(key, stable=True)[1]
    cmap, _, _, _ = ops.coords_insert(cmap.coords[perm].contiguous(), L.SRC_INT, (1, 1, 1))
    del km0, bits, key
out_map = cmap
if args.stride > 1:
    out_map, _, _, _ = ops.coords_insert(cmap.coords, L.SRC_STRIDE, (args.stride,) * 3)
km = ops.build_kernel_map(cmap, out_map, ops.kernel_offsets((args.ksize,) * 3, (1, 1, 1), (1, 1, 1)))
K = args.ksize ** 3
g = torch.Generator().manual_seed(0)
x = torch.randn(km.m_in, args.cin, generator=g).to(dev)
w = (torch.randn(K, args.cin, args.cout, generator=g) / (K * args.cin) ** 0.5).to(dev)
go = torch.randn(km.m_out, args.cout, generator=g).to(dev)
_ = km.mask, km.nbr_t, km.mask_t
P = km.n_pairs
flops = 2.0 * P * args.cin * args.cout
xin, gin = (ops.to_bf16(x), ops.to_bf16(go)) if args.prec == "bf16" else (x, go)
fns = {"fwd": lambda: ops.conv_fwd_raw(xin, w, None, km, prec),
       "dgrad": lambda: ops.conv_dgrad_raw(gin, w, km, prec),
       "wgrad": lambda: ops.conv_wgrad_raw(xin, gin, km, K, args.cin, args.cout, prec),
       "cvt": lambda: ops.to_bf16(x)}
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MB > L2
for name in args.only.split(","):
    fn = fns[name]
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:6s} M_in={km.m_in} M_out={km.m_out} K={K} {args.cin}->{args.cout} P={P} {args.prec}: "
          f"{t:.3f} ms  {flops / t / 1e9:.1f} TFLOP/s  gather {P * args.cin * 4 / t / 1e6:.0f} GB/s", flush=True)
