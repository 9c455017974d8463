"""How much (tile, offset) work a row re-ordering would remove from the convolution kernels, on the
CPU oracle: executed / useful = sum over 128-row tiles of popcount(OR of the rows' 27-bit neighbour masks) * 128 / P,
for raster order and for rows sorted by mask inside windows of W rows.
usage: sim_mask_sort.py [voxels]   (config-2 generator; results in DESIGN.md section 3)"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import synth  # noqa: E402
from oracle import ref_ops as R  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
c, _, _ = synth.room_batch(777, 1, n)
uc, _, _ = R.unique_first_c(R.quantize_c(c))
nbr = R.kernel_map_c(uc, uc, R.kernel_offsets((3, 3, 3), (1, 1, 1)))
M, P = uc.shape[0], int((nbr >= 0).sum())
mask = np.zeros(M, np.uint32)
for k in range(27):
    mask |= (nbr[k] >= 0).astype(np.uint32) << np.uint32(k)


def cost(order, T=128):
    m = mask[order]
    m = np.concatenate([m, np.zeros((-len(m)) % T, np.uint32)]).reshape(-1, T)
    o = np.bitwise_or.reduce(m, axis=1)
    return sum(bin(int(x)).count("1") for x in o) * T / P


print(f"M = {M}, P / M = {P / M:.2f}, no skipping: {27 * M / P:.3f}, distinct masks: {len(np.unique(mask))}")
ident = np.arange(M)
print(f"raster order (per-tile offset masks): executed / useful = {cost(ident):.3f}")
for W in (1024, 4096, 16384, 65536, M):
    order = ident.copy()
    for s in range(0, M, W):
        seg = order[s:s + W]
        order[s:s + W] = seg[np.argsort(mask[seg], kind="stable")]
    print(f"rows sorted by mask inside windows of {W:7d}: {cost(order):.3f}")
