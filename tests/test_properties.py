"""Property tests (hypothesis) of the host-side pieces whose inputs are open-ended: the gin-lite parser, the affine
chain algebra, the schedule / warm-up composition, the counts -> metrics conversion."""
import math

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from nerf_downstream_b200 import augment, ginlite, schedules, training

literals = st.recursive(
    st.one_of(st.none(), st.booleans(), st.integers(-10**6, 10**6), st.floats(allow_nan=False, allow_infinity=False, width=32),
              st.text(alphabet="abc#=[]() _-.,'\"", max_size=12)),
    lambda inner: st.one_of(st.lists(inner, max_size=4), st.tuples(inner, inner)), max_leaves=8)
names = st.from_regex(r"[A-Za-z][A-Za-z0-9_]{0,8}", fullmatch=True)


@settings(max_examples=150, deadline=None, derandomize=True)
@given(st.dictionaries(st.tuples(names, names), literals, max_size=6))
def test_ginlite_round_trip(bindings):
    """config_str() of any set of literal bindings parses back to the same bindings (strings with '#', '=', brackets
    and quotes included)."""
    ginlite.clear_config()
    try:
        for (n, p), v in bindings.items():
            ginlite.bind_parameter(f"{n}.{p}", v)
        text = ginlite.config_str()
        want = ginlite.config_dict()
        ginlite.clear_config()
        ginlite.parse_config(text)
        got = ginlite.config_dict()
        assert got.keys() == want.keys()
        for k in want:
            assert repr(got[k]) == repr(want[k]), (k, got[k], want[k])
    finally:
        ginlite.clear_config()


ops_strategy = st.lists(st.one_of(
    st.tuples(st.just("rot"), st.floats(-3, 3), st.floats(0.2, 1), st.floats(-1, 1)),
    st.tuples(st.just("scale"), st.floats(0.2, 3)),
    st.tuples(st.just("move"), st.floats(-50, 50), st.floats(-50, 50), st.floats(-50, 50)),
    st.tuples(st.just("flip"), st.integers(0, 2)),
    st.tuples(st.just("div"), st.floats(0.01, 2))), min_size=1, max_size=8)


@settings(max_examples=100, deadline=None, derandomize=True)
@given(ops_strategy, st.integers(0, 2**31 - 1))
def test_affine_chain_equals_sequential_application(ops, seed):
    pts = np.random.default_rng(seed).uniform(-30, 30, (25, 3))
    chain, cur = augment.AffineChain(), pts.copy()
    for op in ops:
        if op[0] == "rot":
            M = augment.rotation_matrix([op[3], op[2], 0.3], op[1])
            chain.right_multiply(M)
            cur = cur @ M
        elif op[0] == "scale":
            chain.scale(op[1])
            cur = cur * op[1]
        elif op[0] == "move":
            chain.translate(op[1:])
            cur = cur + np.array(op[1:])
        elif op[0] == "flip":
            mx = cur[:, op[1]].max()
            chain.flip(op[1], chain.axis_max(pts, op[1]))
            cur[:, op[1]] = mx - cur[:, op[1]]
        else:
            chain.divide(op[1])
            cur = cur / op[1]
    scale = max(1.0, np.abs(cur).max())
    assert np.abs(chain.apply(pts) - cur).max() <= 1e-9 * scale
    a = np.array(chain.as_affine12())
    assert np.abs(pts @ a[:9].reshape(3, 3).T + a[9:] - cur).max() <= 1e-9 * scale


@settings(max_examples=100, deadline=None, derandomize=True)
@given(st.sampled_from(["PolyLR", "CosineAnnealingLR", "StepLR", "ExponentialLR"]), st.integers(1, 50), st.integers(10, 400),
       st.floats(1e-4, 1.0))
def test_warmup_hands_over_at_the_wrapped_schedules_step_zero(name, warm, max_steps, lr):
    kw = {"PolyLR": {"poly_exp": 0.9}, "StepLR": {"step_size": 7}}.get(name, {})
    plain = schedules.get_schedule(name, lr, max_steps, -1, **kw)
    warmed = schedules.get_schedule(name, lr, max_steps, warm, **kw)
    assert warmed.lr(0) == 0.0 and abs(warmed.lr(warm) - lr) <= 1e-12
    for t in (0, 1, max_steps // 2, max_steps - 1):
        assert warmed.lr(warm + 1 + t) == plain.lr(t)
    assert all(warmed.lr(t) <= warmed.lr(t + 1) + 1e-15 for t in range(warm))          # the ramp is monotone
    assert all(0.0 <= plain.lr(t) <= lr * (1 + 1e-12) for t in range(0, max_steps, max(1, max_steps // 17)))


@settings(max_examples=100, deadline=None, derandomize=True)
@given(st.integers(2, 12), st.integers(1, 300), st.integers(0, 2**31 - 1))
def test_metrics_from_counts_matches_eval_metrics(C, n, seed):
    import torch
    rng = np.random.default_rng(seed)
    target = rng.integers(0, C, n)
    pred = np.where(rng.random(n) < 0.5, target, rng.integers(0, C, n))
    target[rng.random(n) < 0.2] = -255
    counts = np.zeros((3, C), np.int64)
    keep = target != -255
    for c in range(C):
        counts[0, c] = (target[keep] == c).sum()
        counts[1, c] = ((target[keep] == c) & (pred[keep] == c)).sum()
        counts[2, c] = (pred[keep] == c).sum()
    a = training.metrics_from_counts(torch.from_numpy(counts))
    b = training.eval_metrics(torch.nn.functional.one_hot(torch.from_numpy(pred), C).float(), torch.from_numpy(target), C, -255)
    if keep.sum() == 0:
        assert math.isnan(a["OA"]) and math.isnan(b["OA"])
    else:
        assert abs(a["OA"] - b["OA"]) < 1e-3
    assert abs(a["mIoU"] - b["mIoU"]) < 1e-6
