"""Multi-rank host logic on CPU: world_size-2 gloo run of the frame-batch sharding
(SURVEY section 8e).  The compute inside each rank is the CPU oracle through the C ABI
(test infrastructure); what is under test is the split, the per-rank handles and the
gather order -- the sharded result must equal the single-rank result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PORT_SO, PKG
from gfdm_b200 import sharding


@pytest.mark.parametrize('n,world', [(0, 1), (1, 2), (7, 2), (8, 2), (4096, 8), (5, 8), (1000003, 4)])
def test_shard_bounds_partition(n, world):
    spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(n, world, world)


def _worker(rank, world, port, so, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, PKG)
    import torch.distributed as dist
    from gfdm_b200 import capi, design, sharding as sh
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lib = capi.load(so)
        M, K, L = 5, 16, 2
        taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
        d = design.get_random_qam16(7 * M * K, np.random.default_rng(1001)).reshape(7, -1)
        mod = capi.Modulator(M, K, L, taps, lib=lib)
        dem = capi.Demodulator(M, K, L, np.conj(taps), lib=lib)
        res = sh.process_sharded(lambda x: dem.demodulate_batch(mod.modulate_batch(x)), d)
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_rank(port):
    import torch.multiprocessing as mp
    from gfdm_b200 import capi, design
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    free_port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, free_port, PORT_SO, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    M, K, L = 5, 16, 2
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
    d = design.get_random_qam16(7 * M * K, np.random.default_rng(1001)).reshape(7, -1)
    single = capi.Demodulator(M, K, L, np.conj(taps), lib=port).demodulate_batch(
        capi.Modulator(M, K, L, taps, lib=port).modulate_batch(d))
    assert sharded.shape == single.shape and np.array_equal(sharded, single)
