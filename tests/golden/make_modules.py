"""Generate tests/golden/modules_ref.npz: the REFERENCE's own module code, imported unchanged from /root/reference,
driven on this repository's MinkowskiEngine surface with the C ABI answered by the CPU oracle (tests/host_harness.py):

  * `WeightSparseConvolution` / `WeightSparseConvolutionTranspose` (co3d_3d/src/models/mink/modules/sparse_conv.py:267-452;
    `sparsify` :346-379, `forward` :381-425) with pruned kernels incl. a fully pruned offset, 3^3 s1, 2^3 s2 and the
    transposed 2^3 s2: inputs, kernels, biases -> output rows AND output coordinates;
  * `ResNet14(27, 51)` (resnet.py) and `Res16UNet14A(27, 20)` (res16unet.py), train mode, forward + backward of a
    cross-entropy loss: logits, per-parameter gradient norms, and the full gradient of a few small parameters.

The network weights are NOT stored (57 MB): both this script and the replay (tests/test_gpu_golden_modules.py) fill
every state-dict entry from `deterministic_state(...)`, a counter-based generator keyed on the entry's name.
tests/test_gpu_golden_modules.py (-m gpu) replays the fixture on the real libsparseconv_b200.so.

Run from the repository root (needs /root/reference):  python tests/golden/make_modules.py
"""
import sys
import zlib
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from nerf_downstream_b200 import synth  # noqa: E402
from tests import host_harness, ref_harness  # noqa: E402


def deterministic_state(model: torch.nn.Module, seed: int) -> None:
    """Fill every floating-point state-dict entry from a generator keyed on (seed, entry name): conv kernels / linear
    weights ~ N(0, 2 / fan_in), BatchNorm weights in [0.5, 1.5], biases and running means small, running variances
    in [0.5, 1.5]."""
    sd = model.state_dict()
    with torch.no_grad():
        for name, t in sd.items():
            if not t.is_floating_point():
                continue
            rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
            shape = tuple(t.shape)
            if name.endswith("running_var"):
                v = rng.uniform(0.5, 1.5, shape)
            elif name.endswith("running_mean"):
                v = rng.normal(0, 0.1, shape)
            elif name.endswith("bn.weight") or name.endswith("norm.weight"):
                v = rng.uniform(0.5, 1.5, shape)
            elif name.endswith("bias"):
                v = rng.normal(0, 0.05, shape)
            else:   # kernel [K, Cin, Cout] / [Cin, Cout], linear.weight [out, in]
                fan_in = int(np.prod(shape[:-1])) if name.endswith("kernel") else int(shape[-1])
                v = rng.normal(0, (2.0 / max(fan_in, 1)) ** 0.5, shape)
            t.copy_(torch.from_numpy(np.asarray(v, np.float32)))


def _param_report(model, out, prefix):
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(float(p.grad.double().norm()))
        if p.numel() <= 64 * 27 * 4 and (n.endswith("bn.weight") or n.endswith("bias") or "conv0" in n or "conv1.kernel" == n):
            out[f"{prefix}.grad.{n}"] = p.grad.detach().numpy().copy()
    out[f"{prefix}.grad_names"] = np.array(names)
    out[f"{prefix}.grad_norms"] = np.array(norms, np.float64)


def main():
    from _pytest.monkeypatch import MonkeyPatch
    import MinkowskiEngine as ME
    mp = MonkeyPatch()
    host_harness.install(mp, "fp32")
    out = {}
    try:
        sc = ref_harness.load("co3d_3d.src.models.mink.modules.sparse_conv")
        rn = ref_harness.load("co3d_3d.src.models.mink.resnet")
        un = ref_harness.load("co3d_3d.src.models.mink.res16unet")

        # ---- the pruned-weight inference convolution ---------------------------------------------------------------
        coords, feats = synth.random_cloud(11, 1500, extent=7, n_batch=2, channels=32)
        out["wsc.coords"], out["wsc.feats"] = coords, feats
        x = ME.TensorField(coordinates=torch.from_numpy(coords), features=torch.from_numpy(feats)).sparse()
        g = torch.Generator().manual_seed(5)
        down = ME.MinkowskiConvolution(32, 32, kernel_size=2, stride=2, dimension=3)
        with torch.no_grad():
            down.kernel.copy_(torch.randn(down.kernel.shape, generator=g) * 0.1)
        xd = down(x)
        out["wsc.down_kernel"] = down.kernel.detach().numpy().copy()
        cases = [("k3s1", sc.WeightSparseConvolution, dict(kernel_size=3, stride=1), x),
                 ("k2s2", sc.WeightSparseConvolution, dict(kernel_size=2, stride=2), x),
                 ("k2s2_tr", sc.WeightSparseConvolutionTranspose, dict(kernel_size=2, stride=2), xd)]
        for name, cls, kw, inp in cases:
            m = cls(32, 64, dilation=1, bias=True, dimension=3, **kw)
            with torch.no_grad():
                k = torch.randn(m.kernel.shape, generator=g) * (torch.rand(m.kernel.shape, generator=g) > 0.8)
                k[1] = 0                                            # a fully pruned offset
                m.kernel.copy_(k)
                m.bias.copy_(torch.randn(1, 64, generator=g))
                out[f"wsc.{name}.kernel"], out[f"wsc.{name}.bias"] = m.kernel.numpy().copy(), m.bias.numpy().copy()
                m.sparsify("strided")
                y = m(inp)
            out[f"wsc.{name}.out_F"], out[f"wsc.{name}.out_C"] = y.F.numpy().copy(), y.C.numpy().copy()

        # ---- whole networks from the reference's unchanged files -------------------------------------------------
        coords, feats, labels = synth.co3d_batch(5, 3, lattice=24)
        model = rn.ResNet14(in_channel=27, out_channel=51).train()
        deterministic_state(model, 101)
        logits = model(ME.TensorField(coordinates=torch.from_numpy(coords), features=torch.from_numpy(feats)))
        logits = logits if torch.is_tensor(logits) else logits.F
        torch.nn.functional.cross_entropy(logits, torch.from_numpy(labels)).backward()
        out["resnet14.coords"], out["resnet14.feats"], out["resnet14.labels"] = coords, feats, labels
        out["resnet14.logits"] = logits.detach().numpy().copy()
        _param_report(model, out, "resnet14")

        coords, feats, labels = synth.room_batch(3, 2, 1500)
        model = un.Res16UNet14A(in_channel=27, out_channel=20).train()
        deterministic_state(model, 202)
        logits = model(ME.TensorField(coordinates=torch.from_numpy(coords), features=torch.from_numpy(feats)))
        logits = logits if torch.is_tensor(logits) else logits.F
        torch.nn.functional.cross_entropy(logits, torch.from_numpy(labels), ignore_index=255).backward()
        out["unet14a.coords"], out["unet14a.feats"], out["unet14a.labels"] = coords, feats, labels
        out["unet14a.logits"] = logits.detach().numpy().copy()
        _param_report(model, out, "unet14a")
    finally:
        mp.undo()
    path = Path(__file__).resolve().parent / "modules_ref.npz"
    np.savez_compressed(path, **out)
    print(path, f"{path.stat().st_size / 1024:.0f} KiB", len(out), "arrays")


if __name__ == "__main__":
    main()
