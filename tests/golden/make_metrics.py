"""Generate tests/golden/metrics_ref.json from the REFERENCE's own metric helpers,
    co3d_3d/src/utils/__init__.py : precision_at_one (:103-114), fast_hist (:117-123), per_class_iu (:126-128),
                                    IoUAccumulator (:159-197),
imported unchanged from /root/reference (the file needs numpy + torch only), on small seeded label / prediction
vectors that include ignored points, out-of-range labels and a class that never occurs.

Run from the repository root:  python tests/golden/make_metrics.py
"""
import importlib.util
import json
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/co3d_3d/src/utils/__init__.py")


def main():
    spec = importlib.util.spec_from_file_location("ref_utils", REF)
    U = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(U)
    rng = np.random.default_rng(20261017)
    cases = {}
    for name, n, C, ignore in [("scannet20", 400, 20, -255), ("small5", 120, 5, 255), ("all_ignored", 16, 4, -100)]:
        target = rng.integers(0, C, n)
        if name == "small5":
            target[target == 3] = 2                      # class 3 never occurs
        pred = np.where(rng.random(n) < 0.6, target, rng.integers(0, C, n))
        target[rng.random(n) < 0.15] = ignore
        if name == "all_ignored":
            target[:] = ignore
        t, p = torch.from_numpy(target), torch.from_numpy(pred)
        hist = U.fast_hist(p, t, C)
        acc = U.IoUAccumulator(C, ignore)
        acc.accumulate(p, t)
        miou, ious = acc.report()
        oa = U.precision_at_one(p, t, ignore)
        cases[name] = {"n": n, "C": C, "ignore": ignore, "target": target.tolist(), "pred": pred.tolist(),
                       "precision_at_one": None if np.isnan(oa) else oa, "hist": hist.tolist(),
                       "per_class_iu": [None if np.isnan(v) else float(v) for v in U.per_class_iu(hist)],
                       "acc_miou": float(miou), "acc_ious": [float(v) for v in ious],
                       "seen": acc.total_seen.tolist(), "correct": acc.total_correct.tolist(),
                       "positive": acc.total_positive.tolist()}
        print(name, oa, float(miou))
    out = Path(__file__).with_name("metrics_ref.json")
    out.write_text(json.dumps(cases))
    print("wrote", out)


if __name__ == "__main__":
    main()
