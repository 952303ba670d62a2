"""Static adjacency generators against the golden copies dumped from the reference
(`tests/golden/adjacency.npz`, written by `tests/golden/make_golden.py` from
`/root/reference/model/layers/graph.py:341-348` and `model/layers/time.py:37-41`)."""
import os

import numpy as np
import pytest

from dstd_gcn_b200.model.layers.graph import Graph
from dstd_gcn_b200.model.layers.time import Time

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "adjacency.npz"))


@pytest.mark.parametrize("layout,v", [("h36m", 22), ("cmu", 25), ("3dpw", 23)])
def test_graph_matches_reference(layout, v):
    g = Graph(layout)
    a = g.get_all_adjacency()
    ref = GOLD[f"graph_{layout}_all"]
    assert a.shape == ref.shape == (2, v, v) and a.dtype == ref.dtype == np.float64
    np.testing.assert_array_equal(a, ref)
    np.testing.assert_array_equal(g.get_adjacency(), GOLD[f"graph_{layout}_full"])
    # properties SURVEY.md section 8(a) lists: both symmetric, `connect` = I + bones, `part` has a zero diagonal
    for k in range(2):
        np.testing.assert_array_equal(a[k], a[k].T)
    np.testing.assert_array_equal(np.diag(a[0]), np.ones(v))
    np.testing.assert_array_equal(np.diag(a[1]), np.zeros(v))


def test_graph_unknown_layout_raises():
    with pytest.raises(NotImplementedError):       # model/layers/graph.py:296-297
        Graph("nope")


@pytest.mark.parametrize("t", [2, 3, 8, 12, 35, 40, 125])
def test_time_matches_reference(t):
    a = Time(t).get_all_adjacency()
    ref = GOLD[f"time_{t}_all"]
    assert a.shape == ref.shape == (1, t, t)
    np.testing.assert_array_equal(a, ref)


def test_time_quirk_is_not_tridiagonal():
    """sub-diagonal shift plus ones at [0,0] [0,1] [T-2,T-1] [T-1,T-1]."""
    t = 8
    a = Time(t).get_all_adjacency()[0]
    want = np.zeros((t, t))
    for i in range(1, t):
        want[i, i - 1] = 1
    for i, j in ((0, 0), (0, 1), (t - 2, t - 1), (t - 1, t - 1)):
        want[i, j] = 1
    np.testing.assert_array_equal(a, want)
