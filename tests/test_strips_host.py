"""Host-side logic of the row-strip decomposition, CPU only: partitioning, neighbour
ranks, and (gloo, world_size 2 and 3) the halo exchange order and statistic reduction."""
import pytest

from helpers import launch_ranks
from spgg_b200 import strips


def test_partition_covers_the_lattice():
    for L, world in ((32768, 8), (32768, 2), (4096, 4), (200, 3), (100, 8), (48, 5)):
        spans = [strips.strip_rows(L, world, r) for r in range(world)]
        assert spans[0][0] == 0
        for (a, n), (b, _m) in zip(spans[:-1], spans[1:]):
            assert a + n == b
        assert spans[-1][0] + spans[-1][1] == L
        if L % 16 == 0 and L // 16 >= world:
            assert all(n % 16 == 0 for _a, n in spans)      # fast-kernel tile height
    assert strips.strip_rows(32768, 8, 3) == (3 * 4096, 4096)
    with pytest.raises(ValueError):
        strips.strip_rows(8, 4, 0)                           # strips thinner than the halos
    with pytest.raises(ValueError):
        strips.strip_rows(64, 2, 2)


def test_neighbours_are_periodic():
    assert strips.neighbours(8, 0) == (7, 1)
    assert strips.neighbours(8, 7) == (6, 0)
    assert strips.neighbours(2, 0) == (1, 1)
    assert strips.neighbours(1, 0) == (0, 0)


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_stat_reduction_over_gloo(world):
    res = launch_ranks(["host", "gloo"], world, timeout=180)
    for rc, out in res:
        assert rc == 0, out


def test_sweep_plan_groups_and_deals_round_robin():
    from spgg_b200 import sweep
    plist = ([dict(L=200, use_second_order=False, r=r) for r in range(10)] +
             [dict(L=200, use_second_order=True, r=r) for r in range(6)] +
             [dict(L=100, state_representation="action")])
    for world in (1, 2, 8):
        pl = sweep.plan(plist, world, max_batch=4)
        seen = sorted(i for _r, b in pl for i in b)
        assert seen == list(range(len(plist)))                    # every replica exactly once
        for _r, b in pl:
            assert len(b) <= 4
            assert len({sweep.group_key(plist[i]) for i in b}) == 1   # a batch shares its geometry
        assert {r for r, _b in pl} <= set(range(world))
        if world == 8:
            assert len({r for r, _b in pl}) >= 6                   # the work is spread over the GPUs
    assert sweep.plan(plist, 2, 4) == sweep.plan(plist, 2, 4)      # deterministic
