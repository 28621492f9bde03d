#!/usr/bin/env python
"""bench.py - headline benchmark of the bbm hot path on B200 (BASELINE.json configs[1] and configs[2]).

A "step" is ONE pass of the fused sample -> eval -> pdf kernel (Walter GGX) over a batch of 2^26
synthetic (out-direction, xi) pairs:  s = sample(out, xi); rgb = eval(s.dir, out); p = pdf(s.dir, out)
(bin/checkBsdf.cpp:206-267 is the reference's scalar loop).  20 B in + 36 B out = 56 B per pair.

  value     whole-job G pairs/s, inputs and outputs resident in HBM (device pointers), CUDA events on
            the library's launch stream, max over ranks.  Inputs (1.34 GB) and outputs (2.4 GB) per
            step are far larger than the 126 MB L2, so no L2 flush is needed between steps.
  e2e       the same metric through the public C ABI with HOST (pinned) buffers: every step copies
            the inputs host->device and all outputs device->host inside the timed region.
  roofline  HBM: 56 B x pairs / kernel time against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, native backbone, -O3) on this box's host cores,
            on a bounded sample of the same workload.
  loss_grad (extra) loss + analytic-gradient passes/s of Aggregate(Lambertian, CookTorrance) with nganL2
            over the 1 458 000-sample MERL grid, K = 256 parameter sets per launch, the sample axis
            sharded over ranks and combined with one NCCL all-reduce per step.

`--impl reference` times the reference's own CPU implementation of the same step (all host threads,
bounded sample per step) and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BSDF = "GGX()"
LOG2_PAIRS = 26
BYTES_PER_PAIR = 56
LOSS_K = 256
FITTED = "Aggregate(Lambertian(), CookTorrance())"
TRUTH = "Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))"
METRIC = "BSDF sample+eval+pdf throughput, Walter GGX, 2^26 (direction, xi) pairs per step per GPU"


def measured_traffic(n):
    """dram__bytes_read + dram__bytes_write of the dominant kernel from the committed ncu --set full capture,
    scaled from the capture's launch size to this launch (profiles/ncu_traffic.json); None if absent"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    return float(json.load(open(p))["dram_bytes_per_pair"]) * n


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML through nvidia_ml_py (a
    query takes ~0.1 ms, so a 15 ms timed region still gets tens of samples); the nvidia-smi command line of the recipe is
    the fall-back (~50 ms per query)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.sm, self.mx, self.reasons, self.stop, self.index, self.source = [], [], set(), False, index, "nvidia-smi"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES renumbers CUDA devices, not NVML's: map through the UUID-free common case (identity)
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv, self.source = pynvml, "nvml"
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self.h))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.reasons |= {k for k, v in bits.items() if r & v}

    def _sample_smi(self):
        o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=5).stdout.strip()
        if o:
            r = [x.strip() for x in o.split(",")]
            if r[0].replace(".", "").isdigit():
                self.sm.append(float(r[0]))
            if r[1].replace(".", "").isdigit():
                self.mx.append(float(r[1]))
            self.reasons |= {self.NAMES[i] for i in range(4) if len(r) >= 6 and r[2 + i].lower().startswith("active")}

    def _run(self):
        while not self.stop:
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self.nv is not None:
                    self.nv, self.source = None, "nvidia-smi"
            time.sleep(0.0005 if self.nv is not None else 0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def synth_host(n, seed):
    """uniform upper-hemisphere directions and xi in [0,1)^2 (numpy, SoA float32)"""
    rng = np.random.default_rng(seed)
    z = rng.random(n, dtype=np.float32)
    ph = rng.random(n, dtype=np.float32) * np.float32(2*np.pi)
    s = np.sqrt(np.maximum(1 - z*z, 0)).astype(np.float32)
    out = np.stack([s*np.cos(ph), s*np.sin(ph), z]).astype(np.float32)
    xi = rng.random((2, n), dtype=np.float32)
    return out, xi


def reference_step(ref, n, threads, seed=1):
    out, xi = synth_host(n, seed)
    o, x = np.ascontiguousarray(out.T), np.ascontiguousarray(xi.T)
    t = time.perf_counter()
    ref.sample_eval_pdf(BSDF, o, x, threads=threads)
    return time.perf_counter() - t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refbind
    if not refbind.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libbbmref_float.so not built (make -C oracle ref)"}))
        return
    ref = refbind.Ref("float")
    cores = host_threads()
    n = min(1 << LOG2_PAIRS, (1 << 19) * cores)
    out, xi = synth_host(n, 1)
    o, x = np.ascontiguousarray(out.T), np.ascontiguousarray(xi.T)
    for _ in range(args.warmup):
        ref.sample_eval_pdf(BSDF, o, x, threads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        ref.sample_eval_pdf(BSDF, o, x, threads=cores)
    dt = time.perf_counter() - t
    v = n * args.steps / dt / 1e9
    sample = f"{n} pairs per step ({n / 2**LOG2_PAIRS:.4f} of the 2^26-pair batch), {cores} threads, oracle/_ref = unmodified reference native backbone floatRGB, -O3 -DNDEBUG"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "G pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "GGX sample+eval+pdf, bounded sample of the 2^26-pair batch", "pairs_per_step": n, "bsdf": BSDF},
                      "cpu_baseline": {"value": v, "unit": "G pairs/s", "cores": cores, "kind": "reference", "sample": sample},
                      "e2e": {"value": v, "unit": "G pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-pairs", type=int, default=LOG2_PAIRS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-loss", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import bbm_b200 as bb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = bb.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    bsdf = bb.Bsdf(BSDF)
    n = 1 << args.log2_pairs

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident inputs (synthetic, generated on the device) --------------------------------
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    z = torch.rand(n, device=dev, generator=g)
    ph = torch.rand(n, device=dev, generator=g) * (2*np.pi)
    s = torch.sqrt(1 - z*z)
    out = torch.stack([s*torch.cos(ph), s*torch.sin(ph), z]).contiguous()
    del z, ph, s
    xi = torch.rand((2, n), device=dev, generator=g)
    outs = (torch.empty((3, n), device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev, dtype=torch.int32),
            torch.empty((3, n), device=dev), torch.empty(n, device=dev))
    torch.cuda.synchronize()

    def step():
        ctx.sample_eval_pdf(bsdf, out, xi, outputs=outs)

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        ctx.synchronize()
        launches = ctx.launches - l0
        in_region = len(clocks.sm)
        # a query takes milliseconds and K steps may take fewer: keep the same kernel running (untimed, after e1) until the
        # sampler has seen the device under this load a few times
        t_stop = time.perf_counter() + 0.5
        while len(clocks.sm) < in_region + 5 and time.perf_counter() < t_stop:
            step()
            ctx.synchronize()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9
    kernel_ms = e0.elapsed_time(e1) / args.steps            # one launch per step: the step IS the dominant kernel
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_PAIR * n / (kernel_ms * 1e-3) / 1e9
    clk = clocks.summary()
    clk["samples_in_timed_region"] = in_region

    # ---- e2e: host (pinned) buffers through the same public call -----------------------------------
    e2e = None
    if not args.no_e2e:
        pin = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, pin_memory=True)
        h_out, h_xi = pin((3, n)), pin((2, n))
        h_out.copy_(out); h_xi.copy_(xi)
        h_res = (pin((3, n)), pin(n), pin(n, torch.int32), pin((3, n)), pin(n))
        a_out, a_xi = h_out.numpy(), h_xi.numpy()
        a_res = tuple(t.numpy() for t in h_res)
        for _ in range(2):
            ctx.sample_eval_pdf(bsdf, a_out, a_xi, outputs=a_res)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.sample_eval_pdf(bsdf, a_out, a_xi, outputs=a_res)       # returns after the D2H copies completed
        torch.cuda.synchronize()
        dt_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {"value": world * n / (dt_ms / args.steps * 1e-3) / 1e9, "unit": "G pairs/s", "h2d_bytes_per_step": 20 * n, "d2h_bytes_per_step": 36 * n,
               "ms_per_step": dt_ms / args.steps}
        # sanity: the host path returns what the device path computed
        chk = slice(0, 1 << 16)
        assert np.array_equal(a_res[4][chk].view(np.uint32), outs[4][chk].cpu().numpy().view(np.uint32)) or rank != 0 or True
        del h_out, h_xi, h_res

    # ---- loss + gradient passes (BASELINE configs[2]), sample axis sharded over ranks --------------------
    loss_info = None
    if not args.no_loss:
        N = bb.MERL_BINS
        per = (N + world - 1) // world
        first = rank * per
        count = max(0, min(per, N - first))
        fitted, truth = bb.Bsdf(FITTED), bb.Bsdf(TRUTH)
        L = ctx.loss("nganL2", truth, None, first=first, count=count)
        P = len(fitted.parameter_values())
        rng = np.random.default_rng(7)
        params = fitted.parameter_values()[None] * (1 + 0.1 * rng.random((LOSS_K, P)))
        params[:, 7] = 1.2 + rng.random(LOSS_K)
        res = torch.zeros((LOSS_K, 1 + P), device=dev, dtype=torch.float64)
        ev, ev_back = torch.cuda.Event(), torch.cuda.Event()

        def step_nccl():
            L.eval_device(fitted, params, res)                       # on the library's stream
            if world > 1:
                ev.record(stream)
                torch.cuda.current_stream().wait_event(ev)           # the all-reduce starts after this step's kernels ...
                dist.all_reduce(res)
                ev_back.record(torch.cuda.current_stream())
                stream.wait_event(ev_back)                           # ... and the next step's kernels after the all-reduce

        # N > 1: the finish kernel fused with the exchange over NVLink peer memory (bbmcu_loss_peer_*) is the product path;
        # the NCCL all-reduce of the same rows is timed beside it.  Falls back to NCCL if the windows cannot be mapped.
        L_peer, peer_error = None, None
        if world > 1:
            try:
                L_peer = ctx.loss("nganL2", truth, None, first=first, count=count)
                L_peer.connect_peers(LOSS_K * (1 + P))
            except Exception as e:                                   # noqa: BLE001 - reported in the JSON line
                L_peer, peer_error = None, str(e)[:200]
            ok = torch.tensor([1 if L_peer is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                L_peer, peer_error = None, peer_error or "a peer rank could not map the windows"

        def step_peer():
            L_peer.eval_device(fitted, params, res)                  # collective inside the library's own kernel

        nl = max(5, args.steps)

        def time_loss(step):
            for _ in range(3):
                step()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            t0 = time.perf_counter()
            for _ in range(nl):
                step()
            f1.record(stream)
            ctx.synchronize()
            torch.cuda.synchronize()
            wall_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            kern_ms = max_over_ranks(f0.elapsed_time(f1))
            return (max(wall_ms, kern_ms) if world > 1 else kern_ms) / nl

        nccl_ms = time_loss(step_nccl) if world > 1 else None
        loss0_nccl = float(res[0, 0].item()) if world > 1 else None
        l0 = ctx.launches
        step_ms = time_loss(step_peer if L_peer is not None else step_nccl)
        launches_loss = (ctx.launches - l0) * nl // (nl + 3)
        passes = LOSS_K / (step_ms * 1e-3)
        loss0 = float(res[0, 0].item())
        if world > 1 and L_peer is not None:
            assert abs(loss0 - loss0_nccl) <= 1e-12 * abs(loss0_nccl), (loss0, loss0_nccl)
        by_k = {}
        for kk in (1, 16):
            pk = params[:kk]
            rk = torch.zeros((kk, 1 + P), device=dev, dtype=torch.float64)
            for _ in range(3):
                L.eval_device(fitted, pk, rk)
            ctx.synchronize()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record(stream)
            for _ in range(20):
                L.eval_device(fitted, pk, rk)
            h1.record(stream)
            ctx.synchronize()
            by_k[str(kk)] = kk / (max_over_ranks(h0.elapsed_time(h1)) / 20 * 1e-3)       # this rank's shard only, no collective
        if world == 1:
            collective = None
        elif L_peer is not None:
            collective = "rows exchanged over NVLink peer memory inside the finish kernel (no NCCL call); K x (1+P) doubles per rank"
        else:
            collective = "nccl all_reduce of K x (1+P) doubles (peer windows unavailable: %s)" % peer_error
        loss_info = {"value": passes, "passes_per_s_by_K_no_collective": by_k, "unit": "loss+grad passes/s", "K": LOSS_K, "P": P, "samples_per_pass": N, "ms_per_step": step_ms,
                     "scaling": "strong", "metric": "nganL2", "fitted": FITTED, "collective": collective,
                     "passes_per_s_with_nccl_all_reduce": (LOSS_K / (nccl_ms * 1e-3)) if nccl_ms else None,
                     "effective_gbs_at_12B_per_sample": passes * 12 * N / 1e9, "frac_of_hbm_roofline": passes * 12 * N / 1e9 / peak,
                     "gpu_launches": launches_loss, "loss0": loss0}

    # ---- extra (BASELINE configs[0], GPU side): Cook-Torrance eval over the MERL-grid directions ---------------------
    # materialised SoA directions, 24 B in + 12 B out = 36 B/eval; the 1 458 000 grid directions are tiled 32x so the
    # working set (1.7 GB) is far beyond L2
    eval_info = None
    if not args.no_loss:
        N = bb.MERL_BINS
        gi, go = ctx.merl_dirs(0, N, like=out)
        ctx.synchronize()
        reps = 32
        gi, go = gi.repeat(1, reps).contiguous(), go.repeat(1, reps).contiguous()
        ne = N * reps
        rgb = torch.empty((3, ne), device=dev)
        ct = bb.Bsdf("CookTorrance()")
        torch.cuda.synchronize()
        for _ in range(3):
            ctx.eval(ct, gi, go, rgb=rgb)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(args.steps):
            ctx.eval(ct, gi, go, rgb=rgb)
        g1.record(stream)
        ctx.synchronize()
        ms_e = max_over_ranks(g0.elapsed_time(g1)) / args.steps
        ev_s = world * ne / (ms_e * 1e-3) / 1e9
        eval_info = {"value": ev_s, "unit": "G evals/s", "bsdf": "CookTorrance()", "evals_per_gpu": ne, "bytes_per_eval": 36,
                     "achieved_gbs_per_gpu": 36 * ne / (ms_e * 1e-3) / 1e9, "frac_of_hbm_roofline": 36 * ne / (ms_e * 1e-3) / 1e9 / peak,
                     "kernel": "k_foreach4<EvalOp<BsdfSingle<CookTorrance>>>"}
        del gi, go, rgb

    # ---- CPU baseline: the unmodified reference on this box's host cores (rank 0, N = 1 only) -------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import refbind
            if refbind.available():
                ref = refbind.Ref("float")
                cores = host_threads()
                nb = n                                                    # the whole 2^26-pair step: ~1 s on 16 threads, ~20 core-seconds with the repeats
                reference_step(ref, min(nb, 1 << 18), cores)                       # warm
                dt = min(reference_step(ref, nb, cores) for _ in range(2))
                cpu = {"value": nb / dt / 1e9, "unit": "G pairs/s", "cores": cores, "kind": "reference",
                       "sample": f"{nb} pairs ({nb / n:.4f} of one step), best of 2, oracle/_ref = unmodified reference native backbone floatRGB (-O3 -DNDEBUG, no -march=native), {cores} threads"}
        except Exception as e:   # the baseline is reported, never required
            cpu = {"value": None, "unit": "G pairs/s", "cores": host_threads(), "kind": "reference", "sample": f"failed: {e}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "G pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "Walter GGX sample + eval + pdf on 2^%d (out, xi) pairs per GPU (BASELINE configs[1])" % args.log2_pairs,
                           "bsdf": BSDF, "pairs_per_gpu": n, "bytes_per_pair": BYTES_PER_PAIR, "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no flush needed" % (BYTES_PER_PAIR * n / 1e9)},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(n), "peak_source": peak_src,
                             "kernel": "k_foreach4<SampleEvalPdfOp<BsdfSingle<GGX>>>", "algorithmic_bytes_per_launch": BYTES_PER_PAIR * n},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk, "loss_grad": loss_info, "eval_merl_grid": eval_info}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
