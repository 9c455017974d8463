"""One process, one map, many knob settings of the tcgen05 convolution kernels (experiment library).
usage: SPARSECONV_B200_LIB=.../libsparseconv_b200_exp.so python scripts/exp/conv_knobs.py VOXELS "idx=val,idx=val;idx=val" [fwd,dgrad,wgrad]
Each ';'-separated group is one setting of spc_debug_set knobs (unset knobs are 0)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import ops, synth  # noqa: E402

voxels = int(sys.argv[1])
settings = sys.argv[2].split(";")
only = (sys.argv[3] if len(sys.argv) > 3 else "fwd").split(",")
shapes = [(96, 96), (32, 32), (64, 64), (128, 128), (256, 256)]
dev = torch.device("cuda:0")
lib = L.load()
c, _, _ = synth.room_batch(777, 1, voxels, channels=1)
cmap, _, _, _ = ops.coords_insert(torch.from_numpy(c).to(dev), L.SRC_FLOAT, (1, 1, 1))
km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
_ = km.mask, km.nbr_t, km.mask_t
P = km.n_pairs
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
prec = L.PREC_BF16
for cin, cout in shapes:
    g = torch.Generator().manual_seed(0)
    x = ops.to_bf16(torch.randn(km.m_in, cin, generator=g).to(dev))
    go = ops.to_bf16(torch.randn(km.m_out, cout, generator=g).to(dev))
    w = (torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5).to(dev)
    fns = {"fwd": lambda: ops.conv_fwd_raw(x, w, None, km, prec),
           "dgrad": lambda: ops.conv_dgrad_raw(go, w, km, prec),
           "wgrad": lambda: ops.conv_wgrad_raw(x, go, km, 27, cin, cout, prec)}
    for st in settings:
        for i in range(8):
            lib.spc_debug_set(i, 0)
        for kv in filter(None, st.split(",")):
            i, v = kv.split("=")
            lib.spc_debug_set(int(i), int(v))
        for name in only:
            fn = fns[name]
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            if hasattr(lib._handle, "__class__") and name == "fwd":
                try:
                    import ctypes
                    raw = ctypes.CDLL(None)  # noqa: F841
                    fn_rc = getattr(ctypes.CDLL(str(L.os.environ["SPARSECONV_B200_LIB"])), "spc_debug_role_cycles")
                    buf = (ctypes.c_longlong * 16)()
                    fn_rc(buf, 1)
                    fn()
                    torch.cuda.synchronize()
                    fn_rc(buf, 1)
                    v = list(buf)
                    nst = max(v[7], 1)
                    print(f"      per CTA-stage (cycles): producer warp0 loop {v[0] * 8 / nst:.0f}/8 = own-stage period; "
                          f"wait-empty {v[1] * 8 / nst:.0f} copy-issue {v[2] * 8 / nst:.0f} bookkeeping {v[3] * 8 / nst:.0f} | "
                          f"MMA loop {v[4] / nst:.0f} wait-full {v[5] / nst:.0f} wait-acc {v[6] / nst:.0f} | "
                          f"epilogue wait {v[8] / nst:.0f} body {v[9] / nst:.0f}  (stages {nst})")
                except Exception as e:  # product library: no counters
                    print("      (no role counters:", e, ")")
            print(f"{name:5s} {cin:3d}->{cout:3d} M={km.m_out} knobs[{st:12s}]: {t:.3f} ms  {2.0 * P * cin * cout / t / 1e9:7.1f} TFLOP/s",
                  flush=True)
