"""Linear == Tree differential property (src/model/tests.rs:50-93), seeded."""
import numpy as np
import pytest

import oracle_lib as o

CASES = [(4, 10, 16, 3000), (4, 14, 16, 2000), (8, 10, 16, 3000), (8, 14, 16, 2000), (8, 22, 24, 2000),
         (8, 24, 30, 1500), (8, 30, 32, 1500), (12, 14, 16, 1500), (12, 22, 24, 800)]


@pytest.mark.parametrize("s,f,c,iters", CASES)
def test_models_encode_equivalent(s, f, c, iters):
    rng = np.random.default_rng(1000 + s * 100 + f)
    lin, tree = o.Model(o.LINEAR, s, f, c), o.Model(o.TREE, s, f, c)
    for it in range(iters):
        assert lin.total_frequency() == tree.total_frequency()
        if it % 500 == 0:
            assert (lin.get_freq_table() == tree.get_freq_table()).all()
        sym = int(rng.integers(0, (1 << s) + 1))
        assert lin.get_frequency(sym) == tree.get_frequency(sym)
    bad = (1 << s) + 1
    assert lin.get_frequency(bad)[0] == o.INVALID_INPUT and lin.get_frequency(bad + 1)[0] == o.INVALID_INPUT
    assert tree.get_frequency(bad)[0] == o.INVALID_INPUT
    if (1 << f) - 1 - ((1 << s) + 1) < iters:
        assert lin.total_frequency() == (1 << f) - 1, "frozen regime reached"


@pytest.mark.parametrize("s,f,c,iters", CASES)
def test_models_decode_equivalent(s, f, c, iters):
    rng = np.random.default_rng(2000 + s * 100 + f)
    lin, tree = o.Model(o.LINEAR, s, f, c), o.Model(o.TREE, s, f, c)
    for it in range(iters):
        assert lin.total_frequency() == tree.total_frequency()
        value = int(rng.integers(0, lin.total_frequency()))
        a, b = lin.get_symbol(value), tree.get_symbol(value)
        assert a == b and a[0] == o.OK
        assert a[2] <= value < a[3]
    total = lin.total_frequency()
    for m in (lin, tree):
        assert m.get_symbol(total)[0] == o.INVALID_INPUT
        assert m.get_symbol(total + 1)[0] == o.INVALID_INPUT


def test_count_is_position_determined():
    """total at step t = min(NSYM + t, FMAX)."""
    m = o.Model(o.TREE, 8, 10, 16)
    rng = np.random.default_rng(7)
    for t in range(1200):
        assert m.total_frequency() == min(257 + t, 1023)
        m.get_frequency(int(rng.integers(0, 256)))
