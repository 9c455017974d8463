"""Try every plausible MN-major operand layout for the tcgen05 wgrad kernel and report which one
reproduces the CUDA-core fp32 result (one GPU run instead of one per guess)."""
import itertools
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
c, _ = synth.random_cloud(5, 20000, extent=14, n_batch=2)
cmap, _, _, _ = ops.coords_insert(torch.from_numpy(c).to(dev), L.SRC_FLOAT, (1, 1, 1))
km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
CB = 64 * 128
for cin, cout in [(32, 32), (64, 96)]:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(cmap.size, cin, generator=g).to(dev)
    go = torch.randn(cmap.size, cout, generator=g).to(dev)
    ref = ops.conv_wgrad_raw(x, go, km, 27, cin, cout, L.PREC_FP32)
    for layout, swz, (lbo, sbo) in itertools.product([1, 2], [0, 1], [(CB, 512), (512, CB), (CB, 1024), (1024, CB)]):
        lib.spc_debug_set(0, layout)
        lib.spc_debug_set(1, swz + 1)
        lib.spc_debug_set(2, lbo)
        lib.spc_debug_set(3, sbo)
        out = ops.conv_wgrad_raw(x, go, km, 27, cin, cout, L.PREC_TF32)
        torch.cuda.synchronize()
        rel = float((out - ref).norm() / ref.norm())
        print(f"C {cin}->{cout} layout={layout} swz={swz} lbo={lbo} sbo={sbo} rel_err={rel:.3e} {'<== OK' if rel < 5e-3 else ''}",
              flush=True)
    for i in range(4):
        lib.spc_debug_set(i, 0)
