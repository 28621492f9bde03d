#!/usr/bin/env python
"""Multi-GPU check + timing of the loss exchange over peer memory against the NCCL all-reduce it replaces.
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/peer_exchange_check.py [--out f.json]
Each rank owns one shard of the nganL2 loss of Aggregate(Lambertian, CookTorrance) over the MERL grid (BASELINE configs[2]).
(a) parity: per-rank totals + dist.all_reduce == fused exchange (1e-13 relative; the NCCL order of additions differs),
    the exchange result is bit-identical on every rank;  (b) time per batch of K parameter sets, both ways, device-timed,
    max over ranks."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=256)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bbm_b200 as bb
    from bbm_b200.shard import shard_range
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = bb.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    fitted = bb.Bsdf("Aggregate(Lambertian(), CookTorrance())")
    truth = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), CookTorrance([0.3,0.3,0.3], 0.2, 1.5))")
    first, count = shard_range(bb.MERL_BINS, rank, world)
    rng = np.random.default_rng(7)
    p0 = fitted.parameter_values()
    P, K = len(p0), a.K
    params = p0[None] * (1 + 0.2 * rng.random((K, P)))
    params[:, 7] = 1.2 + rng.random(K)
    L_nccl = ctx.loss("nganL2", truth, None, first=first, count=count)
    L_peer = ctx.loss("nganL2", truth, None, first=first, count=count)
    L_peer.connect_peers(K * (1 + P))
    r_nccl = torch.zeros((K, 1 + P), device=dev, dtype=torch.float64)
    r_peer = torch.zeros((K, 1 + P), device=dev, dtype=torch.float64)
    ev, ev_back = torch.cuda.Event(), torch.cuda.Event()

    def step_nccl():
        L_nccl.eval_device(fitted, params, r_nccl)
        ev.record(stream)
        torch.cuda.current_stream().wait_event(ev)
        dist.all_reduce(r_nccl)
        ev_back.record(torch.cuda.current_stream())
        stream.wait_event(ev_back)

    def step_peer():
        L_peer.eval_device(fitted, params, r_peer)

    def timed(fn):
        for _ in range(5):
            fn()
        ctx.synchronize(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            fn()
        e1.record(stream)
        ctx.synchronize(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_nccl(); step_peer()
    ctx.synchronize(); torch.cuda.synchronize()
    rel = float(((r_peer - r_nccl).abs() / (r_nccl.abs() + 1e-300)).max().item())
    gathered = [torch.zeros_like(r_peer) for _ in range(world)]
    dist.all_gather(gathered, r_peer)
    identical = all(torch.equal(g, gathered[0]) for g in gathered)
    ms_nccl = timed(step_nccl)
    ms_peer = timed(step_peer)
    ms_nccl2 = timed(step_nccl)
    ms_peer2 = timed(step_peer)
    if rank == 0:
        out = {"n_gpus": world, "K": K, "P": P, "samples": bb.MERL_BINS, "max_rel_diff_peer_vs_nccl": rel, "peer_result_bit_identical_on_all_ranks": identical,
               "ms_per_batch_nccl": min(ms_nccl, ms_nccl2), "ms_per_batch_peer": min(ms_peer, ms_peer2),
               "passes_per_s_nccl": K / (min(ms_nccl, ms_nccl2) * 1e-3), "passes_per_s_peer": K / (min(ms_peer, ms_peer2) * 1e-3), "loss0": float(r_peer[0, 0].item())}
        print(json.dumps(out))
        if a.out:
            json.dump(out, open(a.out, "w"), indent=1)
        assert rel < 1e-12 and identical, out
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
