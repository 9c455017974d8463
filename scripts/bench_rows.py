"""Row-kernel timing targets for ncu (BN statistics / apply / backward, hash insert, kernel map, decode) on a
1 M-voxel scene — one pass of each, so `ncu -k regex:...` captures exactly these launches.
usage: bench_rows.py [voxels] [C]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import ops, pipeline, synth  # noqa: E402

voxels = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 96
dev = torch.device("cuda:0")
ops.set_default_precision("bf16")
coords, feats, labels = synth.room_batch(777, 1, voxels)
c = torch.from_numpy(coords).to(dev)
for rep in range(2):
    cmap, first, inverse, count = ops.coords_insert(c, L.SRC_FLOAT, (1, 1, 1))
    km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
    _ = km.mask, km.nbr_t
    m = cmap.size
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(m, C, generator=g).to(dev).requires_grad_()
    gamma, beta = torch.ones(C, device=dev, requires_grad=True), torch.zeros(C, device=dev, requires_grad=True)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    y = ops.BatchNormFn.apply(x, gamma, beta, rm, rv, True, 0.1, 1e-5, True, None)
    y.backward(torch.randn(m, C, generator=g).to(dev))
    recs = synth.compact_records(coords, feats, labels)
    for b, links, sh, lab, reso in recs:
        pipeline.plenoxel_decode(torch.from_numpy(links).to(dev), torch.from_numpy(sh).to(dev), 2 / 255, -1.0, reso, b,
                                 affine=(1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, 0.5, 0.5))
    torch.cuda.synchronize()
print("rows", m, "C", C)
