"""Worker processes of the strip tests (launched by test_strips_host.py / test_gpu_strips.py
with RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def host_logic(rank, world):
    """gloo, CPU tensors: halo exchange reproduces the periodic neighbours of a global
    lattice; statistic rows reduce to whole-lattice rows."""
    import torch
    import torch.distributed as dist
    from spgg_b200 import strips, _lib
    L, GH = 48, 2
    rs = np.random.RandomState(3)
    plane = rs.randint(0, 255, (L, 7)).astype(np.uint8)          # same on every rank
    row0, rows = strips.strip_rows(L, world, rank, align=1)
    mine = plane[row0:row0 + rows]
    to_up = torch.from_numpy(mine[:GH].copy().reshape(-1))
    to_down = torch.from_numpy(mine[-GH:].copy().reshape(-1))
    from_up, from_down = torch.empty_like(to_up), torch.empty_like(to_down)
    for _ in range(3):                                            # repeated exchanges keep matching
        strips.exchange_halos(dist, to_up, to_down, from_up, from_down, rank, world)
        want_up = plane[[(row0 - GH + k) % L for k in range(GH)]].reshape(-1)
        want_down = plane[[(row0 + rows + k) % L for k in range(GH)]].reshape(-1)
        assert np.array_equal(from_up.numpy(), want_up), "top ghosts"
        assert np.array_equal(from_down.numpy(), want_down), "bottom ghosts"
    rows_local = np.zeros((5, _lib.NSTAT))
    rows_local[:, _lib.ST_NC_OLD] = rank + 1
    rows_local[:, _lib.ST_SUM_Q] = 0.5 * (rank + 1)
    rows_local[:, _lib.ST_GMAX] = [0.1 * (rank + 1)] * 5
    red = strips.reduce_stat_rows(dist, rows_local, world)
    tot = world * (world + 1) / 2
    assert np.all(red[:, _lib.ST_NC_OLD] == tot) and np.allclose(red[:, _lib.ST_SUM_Q], 0.5 * tot)
    assert np.allclose(red[:, _lib.ST_GMAX], 0.1 * world)
    # the partition covers the lattice exactly once
    parts = [None] * world
    dist.all_gather_object(parts, (row0, rows))
    assert parts[0][0] == 0 and sum(r for _, r in parts) == L
    for (a, n), (b, _m) in zip(parts[:-1], parts[1:]):
        assert a + n == b


def gpu_strips(rank, world, precision, second, state, L, n_steps):
    """Real kernels: N strips (each rank one strip; ranks may share cuda:0 under gloo) ==
    the single-handle run of the same seed, bit for bit."""
    import torch
    import torch.distributed as dist
    import spgg_b200
    from spgg_b200 import strips
    from helpers import C1, full_params
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    p = full_params(dict(C1, L=L, use_second_order=second, state_representation=state))
    rs = np.random.RandomState(21)
    Q0 = rs.uniform(-0.01, 0.01, (L, L, 2, 2))
    S0 = rs.randint(0, 2, (L, L))
    R0 = np.zeros((L, L))
    se = strips.StripEngine(p, seed=77, precision=precision, device=dev)
    se.set_state_global(S0, R0, Q0)
    se.step(n_steps // 2)
    rows_a = se.stats()
    se.step(n_steps - n_steps // 2)
    rows_b = se.stats()
    S, R, Q = se.gather_state()
    if os.environ.get("SPGG_SPEC_TEST_POISON") and se.spec_mode:
        assert se.reruns > 0, "the spoiled guesses were not re-run"
    se.close()
    if rank == 0:
        eng = spgg_b200.Engine(p, seeds=77, precision=precision, device=dev)
        eng.set_state(S0, R0, Q0)
        eng.step(n_steps // 2)
        ra = eng.stats()
        eng.step(n_steps - n_steps // 2)
        rb = eng.stats()
        S1, R1, Q1 = eng.get_state()
        eng.close()
        assert np.array_equal(S, S1), f"{(S != S1).sum()} strategy mismatches"
        assert np.array_equal(R, R1)
        assert np.array_equal(Q, Q1)
        ints = [0, 1, 2, 3, 11, 12, 13, 14, 15, 16, 31, 32, 33]
        for got, want in ((rows_a, ra), (rows_b, rb)):
            assert np.array_equal(got[1:, ints], want[1:, ints])
            assert np.allclose(got[:, 17], want[:, 17])
            np.testing.assert_allclose(got[1:, 4:11], want[1:, 4:11], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(got[1:, 18:31], want[1:, 18:31], rtol=1e-4, atol=1e-3)
    dist.barrier()


def gpu_exit(rank, world, L):
    """spgg.py:405 over strips: every site prefers one action, epsilon = 0 -> the lattice is uniform
    after iteration 1 and iteration 2 breaks before acting - on every rank, although no rank sees
    more than its strip.  Same final state and iteration count as the single handle."""
    import torch
    import torch.distributed as dist
    import spgg_b200
    from spgg_b200 import strips
    from helpers import C1, full_params
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    p = full_params(dict(C1, L=L, epsilon=0.0, epsilon_min=0.0))
    rs = np.random.RandomState(8)
    S0 = rs.randint(0, 2, (L, L))
    R0 = np.zeros((L, L))
    for prefer in (0, 1):                                   # all cooperate / all defect
        Q0 = np.zeros((L, L, 2, 2))
        Q0[..., prefer] = 1.0
        se = strips.StripEngine(p, seed=3, precision="fp32", device=dev)
        assert se.eng.describe().startswith("fast")
        se.set_state_global(S0, R0, Q0)
        se.step(6)
        se.step(4)                                           # a chunk after the stop does nothing
        assert se.stopped_at() == 1 and se.iteration == 1, (se.stopped_at(), se.iteration)
        S, R, Q = se.gather_state()
        se.close()
        assert (S == prefer).all()
        if rank == 0:
            eng = spgg_b200.Engine(p, seeds=3, precision="fp32", device=dev)
            eng.set_state(S0, R0, Q0)
            eng.step(10)
            st = eng.status()
            assert st.stopped_at == 1 and st.iteration == 1
            S1, R1, Q1 = eng.get_state()
            eng.close()
            assert np.array_equal(S, S1) and np.array_equal(R, R1) and np.array_equal(Q, Q1)
    # uniform from the start: nothing runs at all
    se = strips.StripEngine(p, seed=3, precision="fp32", device=dev)
    se.set_state_global(np.ones((L, L), np.uint8), R0, np.zeros((L, L, 2, 2)))
    se.step(5)
    assert se.stopped_at() == 0 and se.iteration == 0
    se.close()
    dist.barrier()


def gpu_sweep(rank, world):
    """Replica sweep sharded over ranks == the same replicas run one by one."""
    import torch
    import spgg_b200
    from spgg_b200 import sweep
    from helpers import C1, C2, full_params
    torch.cuda.set_device(rank % torch.cuda.device_count())
    plist = [full_params(dict(C1, L=64, r=r, influence_factor=k)) for r in (3.0, 4.0) for k in (0.0, 1.0)]
    plist += [full_params(dict(C2, L=64, r=r)) for r in (3.6, 5.0)]
    plist += [full_params(dict(C1, L=128, r=3.0))]
    seeds = [100 + i for i in range(len(plist))]
    res = sweep.run_sweep(plist, seeds, iterations=25, max_batch=3, chunk=10)
    assert len(res) == len(plist)
    if rank == 0:
        for p, sd, o in zip(plist, seeds, res):
            single = sweep.run_batch([p], [sd], 25, chunk=25, device=torch.cuda.current_device())[0]
            assert o["iterations"] == single["iterations"] == 25
            assert np.array_equal(o["S"], single["S"]) and np.array_equal(o["R"], single["R"])
            for k in ("coop_rate_history", "switch_C_to_D", "group_comp_d2_history"):
                assert np.array_equal(o["series"][k], single["series"][k]), k
            np.testing.assert_allclose(o["series"]["avg_q_s0_c_history"],
                                       single["series"]["avg_q_s0_c_history"], rtol=1e-6, atol=1e-9)


def main():
    import torch.distributed as dist
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend = sys.argv[2]
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        if mode == "host":
            host_logic(rank, world)
        elif mode == "sweep":
            gpu_sweep(rank, world)
        elif mode == "exit":
            gpu_exit(rank, world, int(sys.argv[3]))
        else:
            precision, second, state, L, n = sys.argv[3:8]
            gpu_strips(rank, world, precision, second == "1", state, int(L), int(n))
    finally:
        dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
