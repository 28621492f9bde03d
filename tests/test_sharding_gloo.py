"""CPU, world_size 2, gloo: the N>1 path of the loss - shard the sample axis, per-rank partial sums
divided by the full N, one all-reduce(sum) - gives the single-rank loss and gradient.  The per-rank
arithmetic is the device code compiled for the host (tests/hostsim)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bbm_b200 as bb
    from bbm_b200.shard import all_reduce_sum, shard_range
    from tests.hostsim.bind import HostSim, soa
    h = HostSim()
    fitted = "Aggregate(Lambertian(), CookTorrance())"
    truth = "Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))"
    hp, tp = float(np.float32(2)*np.float32(np.pi)), float(np.float32(0.5)*np.float32(np.pi))
    N = 13*8*5*6
    P = len(bb.Bsdf(fitted).parameter_values())
    first, count = shard_range(N, rank, world)
    i, o = h.spherical_dirs([13, 8, 5, 6], [0, 0, hp, tp, 0, 0, hp, tp], first, count)
    ref = h.eval(truth, i, o)
    loss, grad, _ = h.loss(fitted, 0, i, o, ref, nparams=P)
    part = torch.tensor([[loss] + list(grad)], dtype=torch.float64) * (count / N)      # partial sum / FULL N
    all_reduce_sum(part)
    if rank == 0:
        q.put(part.numpy())
    dist.destroy_process_group()


def test_two_rank_loss_matches_single_rank(golden_loss):
    _, meta = golden_loss
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rec = meta["metrics"]["nganL2"]
    assert abs(got[0, 0] - rec["double_total"]) <= 1e-4*abs(rec["double_total"])
    fd = np.array(rec["fd_gradient"])
    assert np.all(np.abs(got[0, 1:] - fd) <= 1e-4*np.abs(fd) + 1e-9)


def test_shard_ranges_cover_exactly():
    from bbm_b200.shard import partition_by_cost, shard_range
    for n in (0, 1, 7, 1458000):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in blocks) == n
            pos = 0
            for f, c in blocks:
                assert f == pos or c == 0
                pos += c
    parts = partition_by_cost([5, 1, 1, 1, 4, 4], 2)
    assert sorted(sum(parts, [])) == list(range(6))
    assert abs(sum([5, 1, 1, 1, 4, 4][j] for j in parts[0]) - 8) <= 1


def test_interleaved_shards_cover_exactly_and_are_balanced():
    """the index map of BBMCU_LOSS_SHARD_INTERLEAVED (bbm_b200.shard.interleaved_indices mirrors bbmcu_loss.cu): shards are
    disjoint, cover [0, n), keep blocks of 1024 consecutive samples, and differ in size by at most one block"""
    from bbm_b200.shard import interleaved_indices
    for n in (1, 1023, 1024, 1025, 5000, 1458000):
        for world in (1, 2, 3, 8):
            parts = [interleaved_indices(n, r, world) for r in range(world)]
            allidx = np.concatenate(parts)
            assert len(allidx) == n and np.array_equal(np.sort(allidx), np.arange(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1024
            for r, p in enumerate(parts):
                if len(p):
                    assert p[0] == r * 1024 and np.all(np.diff(p)[np.arange(len(p) - 1) % 1024 != 1023] == 1)
