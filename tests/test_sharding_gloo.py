"""Host-side multi-GPU logic on CPU: the channel-shard planner, and a world_size-2 gloo run in which each rank processes
its own shard with the CPU checker (standing in for its GPU) — the assembled result must equal the single-process run,
with no collective on the data path (only the optional gather and the max-over-ranks timing reduction)."""
import os
import socket

import numpy as np
import pytest


def test_plan_covers_every_channel_once(msdr):
    for total in (0, 1, 31, 32, 33, 4096, 4097, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            shards = msdr.shard.plan(total, world)
            assert len(shards) == world and shards[0].ch0 == 0
            assert sum(s.n for s in shards) == total
            for a, b in zip(shards, shards[1:]):
                assert a.ch0 + a.n == b.ch0
            sizes = [s.n for s in shards]
            assert max(sizes) - min(sizes) < 64 or total < 32 * world
            assert all(s.ch0 % 32 == 0 for s in shards if s.n)
            assert msdr.shard.plan(total, world, rank=world - 1) == shards[-1]


def test_weak_scaling_shards(msdr):
    s = [msdr.shard.weak_scaling_shard(4096, 8, r) for r in range(8)]
    assert [x.ch0 for x in s] == [r * 4096 for r in range(8)] and all(x.n == 4096 for x in s)


def _worker(rank, world, port, total, nb, tmp):
    import torch
    import torch.distributed as dist
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import minimal_sdr_b200 as m
    import oracle_lib as ol
    from chain_helpers import tables_for
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K = m.load_ref_constants()
    sh = m.shard.plan(total, world, rank)
    modes = m.synth.mixed_modes(sh.n, sh.ch0)
    x = m.synth.batch(modes, nb * 128, ch0=sh.ch0)
    orc = ol.CheckerLib("orc")
    o = orc.chain(sh.n)
    for c, md in enumerate(modes):
        o.set_mode(c, 1, md)
        o.fir_init(c, 1, *tables_for(K, md))
    o.biquad_set_coefficients(0, 0, sh.n, 0, K["biquad1_lowpass_coef"])
    o.biquad_set_coefficients(1, 0, sh.n, 0, K["biquad2_notch_coef"])
    y = o.run(x)[0]
    full = m.shard.gather_rows(torch.from_numpy(y), sh, total)       # optional, untimed gather
    slowest = m.shard.max_over_ranks(10.0 + rank)                      # timing rule: max over ranks
    samples = m.shard.sum_over_ranks(float(y.size))                    # whole-job unit count
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), full.numpy())
        np.save(os.path.join(tmp, "meta.npy"), np.array([slowest, samples]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_equivalence(msdr, orc, K, tmp_path):
    import torch.multiprocessing as mp
    from chain_helpers import tables_for
    total, nb, world = 75, 6, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, total, nb, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    slowest, samples = np.load(tmp_path / "meta.npy")
    assert slowest == 11.0 and samples == total * nb * 128
    modes = msdr.synth.mixed_modes(total)
    x = msdr.synth.batch(modes, nb * 128)
    o = orc.chain(total)
    for c, md in enumerate(modes):
        o.set_mode(c, 1, md)
        o.fir_init(c, 1, *tables_for(K, md))
    o.biquad_set_coefficients(0, 0, total, 0, K["biquad1_lowpass_coef"])
    o.biquad_set_coefficients(1, 0, total, 0, K["biquad2_notch_coef"])
    assert np.array_equal(got, o.run(x)[0])
