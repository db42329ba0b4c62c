"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-4   # BASELINE.json north_star: "within 1e-4 rel fp32 on states/rewards"
ATOL = 2e-5   # floor for entries near zero (quaternion components, noise-dominated obs, cancelled reward terms)


def golden_cases():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n.split("_")[0] in ("hovering", "tracking", "balloon", "avoid", "planning")]  # env trajectories


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    task, mode = name.split("_")[0], name.split("_")[1]
    N, T, seed, A, max_len = [int(x) for x in g["meta"]]
    return g, task, mode, N, T, A, max_len


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu()
    return np.asarray(x, dtype=np.float64)


def assert_close(got, ref, what, rtol=RTOL, atol=ATOL):
    got = _np(got)
    ref = _np(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert (nan_g == nan_r).all(), f"{what}: NaN pattern differs at {np.argwhere(nan_g != nan_r)[:5]}"
    err = np.abs(np.nan_to_num(got) - np.nan_to_num(ref))
    bound = atol + rtol * np.abs(np.nan_to_num(ref))
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} entries off; worst at {i}: got {got[i]!r} ref {ref[i]!r} err {err[i]:.3e}")
    return float(err.max()) if err.size else 0.0


def task_tols(task):
    """(rtol, atol) for reward-like outputs: Balloon's guidance term is 30 x (difference of two norms), i.e. it amplifies
    the last-ulp differences of positions ~30x / |value|; everything else uses the north_star bar."""
    return (1e-4, 3e-3) if task == "balloon" else (RTOL, ATOL)


def pokes_at(g, t):
    """[(env, state row [13])] recorded state edits to apply BEFORE step t (tests/golden/make_golden.py run_case pokes)."""
    if "poke_t" not in g:
        return []
    return [(int(e), st) for pt, e, st in zip(g["poke_t"], g["poke_env"], g["poke_state"]) if int(pt) == t]
