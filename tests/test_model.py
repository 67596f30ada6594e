"""Model builder checks that need no GPU: the hull support maps the engine scans instead of all hull vertices."""
import numpy as np
import pytest

from cosim_b200.model import load_robot, _support_map_cached, support_buckets, SUP_G


@pytest.mark.parametrize("robot", ["flamingo_p_v3", "flamingo_light_v1", "w4_p_v2", "humanoid_p_v0"])
def test_support_map_is_complete(robot):
    """For random directions (and directions next to the axes, where bucket borders and flat mesh faces sit) the best
    candidate of the direction's bucket supports as far as the best hull vertex overall."""
    rb = load_robot(robot)
    rng = np.random.default_rng(3)
    D = rng.standard_normal((60000, 3)).astype(np.float32)
    D[:3000] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 3000)] * rng.choice([-1, 1], (3000, 1)).astype(np.float32)
    D[:3000] += (1e-3 * rng.standard_normal((3000, 3))).astype(np.float32)
    bk = support_buckets(D, SUP_G)
    nmesh = 0
    for g in range(len(rb["geom_type"])):
        if rb["geom_type"][g] != 7:
            continue
        v = np.ascontiguousarray(rb["hull_verts"][rb["geom_vadr"][g]:rb["geom_vadr"][g] + rb["geom_vnum"][g]])
        off, idx = _support_map_cached(v)
        P = D.astype(np.float64) @ v.astype(np.float64).T
        best = P.max(axis=1)
        # group directions by bucket
        order = np.argsort(bk, kind="stable")
        starts = np.searchsorted(bk[order], np.arange(len(off)))
        for b in range(len(off) - 1):
            rows = order[starts[b]:starts[b + 1]]
            if len(rows):
                got = P[np.ix_(rows, idx[off[b]:off[b + 1]])].max(axis=1)
                assert (got >= best[rows] - 1e-9).all(), f"{robot} geom {g} bucket {b}: a support vertex is missing from the bucket"
        nmesh += 1
        assert len(idx) / (len(off) - 1) < 40
    assert nmesh > 0
