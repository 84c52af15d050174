"""world_size-2 gloo test (CPU) of the only multi-process logic on this path: the replica-ensemble timing
protocol (barrier, MAX over ranks, whole-job throughput).  There is no data-path collective to test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openmm_rigidbody_plugin_b200 import replicas, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, w, _ = replicas.rank_world()
        assert (r, w) == (rank, world)
        # distinct replicas: different seeds -> different systems of the same size
        sysd = synth.water_box(64, seed=replicas.replica_seed(100, rank))
        replicas.barrier(dist)
        elapsed = 0.5 + 0.25 * rank                       # rank 1 is the slow one
        slowest = replicas.max_over_ranks(elapsed, dist)
        total = replicas.ensemble_throughput(64, 10, elapsed, dist)
        checksum = float(sysd["R"].sum())
        sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(sums, torch.tensor([checksum], dtype=torch.float64))
        out.put((rank, slowest, total, [float(s.item()) for s in sums]))
    finally:
        dist.destroy_process_group()


def test_replica_protocol_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, slowest, total, sums in res:
        assert slowest == pytest.approx(0.75)                   # MAX over ranks, identical on every rank
        assert total == pytest.approx(2 * 64 * 10 / 0.75)       # all ranks' bodies / slowest rank's time
        assert sums[0] != sums[1]                               # the replicas really are different systems


def test_single_process_is_identity():
    assert replicas.max_over_ranks(1.25) == 1.25
    assert replicas.ensemble_throughput(100, 4, 2.0) == 200.0
