"""cProfile of the host side of ResNet14 steps (B = 16): where the ~5.5 ms of Python per step go."""
import cProfile
import pstats
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from nerf_downstream_b200 import me as ME  # noqa: E402
from nerf_downstream_b200 import models, ops, synth, trainer  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_precision("bf16")
torch.manual_seed(0)
model = models.ResNet14(27, 51).to(dev).train()
tr = trainer.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4)
coords, feats, labels = synth.co3d_batch(777, 16)
c, f, y = (torch.from_numpy(a).to(dev) for a in (coords, feats, labels))


def step():
    field = ME.TensorField(coordinates=c, features=f)
    loss = ops.cross_entropy(model(field), y)
    tr.backward_and_step(loss)


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(40)
