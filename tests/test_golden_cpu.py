"""The oracle must reproduce the committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py
from the fp64 oracle): this freezes the checker against drift.  CPU only, a few seconds."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG  # noqa: E402


def _cmp(got, exp, name):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, (name, got.shape, exp.shape)
    if exp.dtype.kind in "iuSU":
        assert (got == exp).all(), name
    else:
        scale = max(float(np.abs(exp).max()), 1e-30)
        assert float(np.abs(got - exp).max()) <= 1e-10 * scale, name


def test_ops_vectors_are_reproduced():
    with np.load(os.path.join(HERE, "golden", "ops_small.npz")) as f:
        exp = {k: f[k] for k in f.files}
    got = MG.ops_vectors()
    assert sorted(got) == sorted(exp)
    for k in exp:
        _cmp(got[k], exp[k], k)


@pytest.mark.parametrize("fname,args", [("train_step_b2_l2_hinge.npz", (2, 2, 2)), ("train_step_b2_l3x1_hinge.npz", (2, 3, 1))])
def test_train_step_vectors_are_reproduced(fname, args):
    with np.load(os.path.join(HERE, "golden", fname)) as f:
        exp = {k: f[k] for k in f.files}
    got = MG.train_step_vectors(*args)
    assert sorted(got) == sorted(exp)
    for k in exp:
        _cmp(got[k], exp[k], k)


def test_golden_filter_bank_index_map_is_bit_exact():
    with np.load(os.path.join(HERE, "golden", "ops_small.npz")) as f:
        m = f["fb_index_map"]
    for l, k, h, w, c in m:
        assert (h, w, c) == (k % 4, 4 * l + k // 2048, (k % 2048) // 4)
