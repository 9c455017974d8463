"""CTA-pair forward / dgrad kernel (conv_umma_pair.cu) against the single-CTA kernel: outputs must be bit-identical
(same MMA sequence per output row), epilogue statistics equal to 1e-12 relative; then timings.
usage: pair_check.py VOXELS CIN COUT [--time]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import ops, synth  # noqa: E402

n, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
lib = L.load()
c, _, _ = synth.room_batch(777, 1, n, channels=1)
cmap, _, _, _ = ops.coords_insert(torch.from_numpy(c).to(dev), L.SRC_FLOAT, (1, 1, 1))
km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
g = torch.Generator().manual_seed(0)
x = torch.randn(km.m_in, cin, generator=g).to(dev)
w = (torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5).to(dev)
go = torch.randn(km.m_out, cout, generator=g).to(dev)
xb, gb = ops.to_bf16(x), ops.to_bf16(go)
_ = km.mask
prec = ops.PRECISIONS["bf16"]


def run(knob):
    lib.spc_debug_set(8, knob)
    out = ops.conv_fwd_raw(xb, w, None, km, prec)
    sums = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    out_s, fused = ops.conv_fwd_raw(xb, w, None, km, prec, bn_sums=sums)
    din = ops.conv_dgrad_raw(gb, w, km, prec)
    torch.cuda.synchronize()
    return out, out_s, sums if fused else None, din


a = run(1)
print("single-CTA kernel done", flush=True)
b = run(2)
print("pair kernel done", flush=True)
ok = True
for name, u, v in (("fwd", a[0], b[0]), ("fwd+stats", a[1], b[1]), ("dgrad", a[3], b[3])):
    same = torch.equal(u, v)
    err = float((u - v).abs().max())
    print(f"{name:10s} bit-identical {same}  max |diff| {err:.3e}  max |ref| {float(u.abs().max()):.3e}", flush=True)
    ok &= same
if a[2] is not None and b[2] is not None:
    rel = float(((a[2] - b[2]).abs() / (a[2].abs() + 1e-30)).max())
    print(f"stats      max relative difference {rel:.3e}", flush=True)
    ok &= rel < 1e-9
else:
    print("stats      fused:", a[2] is not None, b[2] is not None)
print("PAIR KERNEL", "OK" if ok else "MISMATCH", flush=True)
if "--time" in sys.argv:
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for knob, tag in ((1, "single"), (2, "pair")):
        lib.spc_debug_set(8, knob)
        for name, fn in (("fwd", lambda: ops.conv_fwd_raw(xb, w, None, km, prec)),
                         ("dgrad", lambda: ops.conv_dgrad_raw(gb, w, km, prec))):
            fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[3]
            print(f"{tag:6s} {name:5s} {cin}->{cout} M={km.m_out}: {t:.3f} ms  {2.0 * km.n_pairs * cin * cout / t / 1e9:.1f} TFLOP/s", flush=True)
lib.spc_debug_set(8, 0)

# ---- wgrad: pair kernel (knob 9) against the single-CTA kernel: fp32 red.adds in a free order -> tolerance ----
def wgrad(knob):
    lib.spc_debug_set(9, knob)
    dw = ops.conv_wgrad_raw(xb, gb, km, 27, cin, cout, prec)
    torch.cuda.synchronize()
    return dw


wa = wgrad(1)
print("single-CTA wgrad done", flush=True)
wb = wgrad(2)
print("pair wgrad done", flush=True)
werr = float((wa - wb).abs().max() / wa.abs().max())
print(f"wgrad      max |diff| / max |ref| {werr:.3e}", flush=True)
print("PAIR WGRAD", "OK" if werr < 1e-4 else "MISMATCH", flush=True)
if "--time" in sys.argv:
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for knob, tag in ((1, "single"), (2, "pair")):
        lib.spc_debug_set(9, knob)
        fn = lambda: ops.conv_wgrad_raw(xb, gb, km, 27, cin, cout, prec)
        fn()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[3]
        print(f"{tag:6s} wgrad {cin}->{cout} M={km.m_out}: {t:.3f} ms  {2.0 * km.n_pairs * cin * cout / t / 1e9:.1f} TFLOP/s", flush=True)
lib.spc_debug_set(9, 0)
