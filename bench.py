#!/usr/bin/env python
"""bench.py - benchmark of the bbm hot path on B200 (BASELINE.json configs[0..4]).

A "step" is ONE pass of the fused sample -> eval -> pdf kernel (Walter GGX) over a batch of 2^26
synthetic (out-direction, xi) pairs:  s = sample(out, xi); rgb = eval(s.dir, out); p = pdf(s.dir, out)
(bin/checkBsdf.cpp:206-267 is the reference's scalar loop).  20 B in + 36 B out = 56 B per pair.

  value     whole-job G pairs/s, inputs and outputs resident in HBM (device pointers), CUDA events on
            the library's launch stream, max over ranks.  Inputs (1.34 GB) and outputs (2.4 GB) per
            step are far larger than the 126 MB L2, so no L2 flush is needed between steps.
  e2e       the same metric through the public C ABI with HOST (pinned) buffers: every step copies
            the inputs host->device and ALL five outputs device->host inside the timed region.
  e2e_variants  the same call when the caller asks for less: (a) rgb + pdf only (the other output
            pointers NULL: not stored, not copied), (b) rgb + pdf with the inputs drawn on the device
            (bbmcu_sample_eval_pdf_generated; 0 B up), (c) pageable instead of pinned caller memory.
  roofline  HBM: 56 B x pairs / kernel time against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, native backbone, -O3) on this box's host cores.
  eval_merl_grid   configs[0]: Cook-Torrance eval over the MERL-grid directions - materialised (36 B/eval), with the
            linearizer fused into the kernel (12 B/eval), and the reference through bsdf_ptr on 1 and all host threads.
  loss_grad configs[2]: loss + analytic-gradient passes/s of Aggregate(Lambertian, CookTorrance) with nganL2
            over the 1 458 000-sample MERL grid, K = 256 parameter sets per launch, the sample axis
            sharded over ranks and combined inside the library's kernels over NVLink peer memory.
  epd       configs[3]: EPD ("Holzschuch-Pacanowski") eval over the grid and a compass fit, CPU reference beside it.
  loss_multi / sweep   configs[4]: M measured materials x K parameter sets per launch (a true DRAM workload) and a
            bounded miniature of the material x model x metric fit sweep (`--sweep` runs the full 100 x 34 x 6).

`--impl reference` times the reference's own CPU implementation of the same step (all host threads, the
full 2^26-pair batch per step) and prints the same JSON line with "impl": "reference".
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
SWEEP_MODELS_MINI = ["CookTorrance", "GGX", "NganWard", "Phong", "LowMicrofacet", "AshikhminShirleyFull"]


def workload_config(log2_pairs):
    n = 1 << log2_pairs
    return {"workload": "Walter GGX sample + eval + pdf on 2^%d (out, xi) pairs per GPU (BASELINE configs[1])" % log2_pairs,
            "bsdf": BSDF, "pairs_per_gpu": n, "bytes_per_pair": BYTES_PER_PAIR,
            "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no flush needed" % (BYTES_PER_PAIR * n / 1e9)}


def measured_traffic(n):
    """dram__bytes_read + dram__bytes_write of the dominant kernel from the committed ncu --set full capture of THIS launch
    size (profiles/ncu_traffic.json: measured at 2^26 pairs per launch); scaled only if the launch size differs"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    if "dram_bytes_per_launch" in t and int(t.get("pairs_per_launch", 0)) == n:
        return float(t["dram_bytes_per_launch"])
    return float(t["dram_bytes_per_pair"]) * n


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING a timed region (B200_PROFILING.md recipe).  NVML through nvidia_ml_py; the
    nvidia-smi command line of the recipe is the fall-back (~50 ms per query)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.sm, self.mx, self.reasons, self.stop, self.index, self.source = [], [], set(), False, index, "nvidia-smi"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
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
        t0 = time.perf_counter()                    # the first NVML query of a process can take longer than the region it is meant
        while not self.sm and time.perf_counter() - t0 < 3.0:   # to sample: let it complete before the region starts
            time.sleep(0.001)
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


def bind_to_gpu_numa_node(local):
    """per-rank CPU affinity: run (and first-touch the pinned buffers) on the NUMA node the GPU hangs off, when the box
    exposes one (/sys/bus/pci/devices/<bus id>/numa_node, local_cpulist).  Returns what was done, for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        if bus is None:
            return {"numa_node": None, "bound": False}
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, devid)
        node = int(open(path + "/numa_node").read().strip())
        cpus = open(path + "/local_cpulist").read().strip()
        if node < 0 or not cpus:
            return {"numa_node": node, "bound": False}
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids |= set(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if not ids:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, ids)
        return {"numa_node": node, "bound": True, "cpus": len(ids)}
    except Exception as e:                                  # noqa: BLE001 - affinity is an optimisation, never required
        return {"numa_node": None, "bound": False, "error": str(e)[:80]}


def synth_host(n, seed):
    """uniform upper-hemisphere directions and xi in [0,1)^2 (numpy, SoA float32)"""
    rng = np.random.default_rng(seed)
    z = rng.random(n, dtype=np.float32)
    ph = rng.random(n, dtype=np.float32) * np.float32(2*np.pi)
    s = np.sqrt(np.maximum(1 - z*z, 0)).astype(np.float32)
    out = np.stack([s*np.cos(ph), s*np.sin(ph), z]).astype(np.float32)
    xi = rng.random((2, n), dtype=np.float32)
    return out, xi


def run_reference(args):
    """the reference's own CPU implementation of the step, on the FULL batch of the workload (same config as our arm)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refbind
    if not refbind.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libbbmref_float.so not built (make -C oracle ref)"}))
        return
    ref = refbind.Ref("float")
    cores = host_threads()
    n = 1 << args.log2_pairs
    out, xi = synth_host(n, 1)
    o, x = np.ascontiguousarray(out.T), np.ascontiguousarray(xi.T)
    del out, xi
    for _ in range(min(args.warmup, 1)):                      # ~1 s per step on 16 threads: one warm-up pass pages everything in
        ref.sample_eval_pdf(BSDF, o, x, threads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        ref.sample_eval_pdf(BSDF, o, x, threads=cores)
    dt = time.perf_counter() - t
    v = n * args.steps / dt / 1e9
    sample = f"the full step: {n} pairs per step, {cores} threads, oracle/_ref = unmodified reference native backbone floatRGB, -O3 -DNDEBUG"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "G pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": workload_config(args.log2_pairs),
                      "cpu_baseline": {"value": v, "unit": "G pairs/s", "cores": cores, "kind": "reference", "sample": sample},
                      "e2e": {"value": v, "unit": "G pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def synthetic_materials(count, seed=2026):
    """BSDF strings of `count` synthetic MERL-shaped materials: Lambertian + Cook-Torrance with seeded parameters in the range
    of the reference's fits/low_cooktorrance_E1.fit entries (diffuse/specular albedo 0.005-0.5, roughness 0.01-0.5, eta 1.2-2.5)"""
    rng = np.random.default_rng(seed)
    out = {}
    for m in range(count):
        d = np.exp(rng.uniform(np.log(0.005), np.log(0.5), 3))
        s = np.exp(rng.uniform(np.log(0.005), np.log(0.5), 3))
        r = float(np.exp(rng.uniform(np.log(0.01), np.log(0.5))))
        eta = float(rng.uniform(1.2, 2.5))
        out["material-%03d" % m] = "Aggregate(Lambertian([%.5g, %.5g, %.5g]), LowCookTorrance([%.5g, %.5g, %.5g], %.5g, %.5g))" % (*d, *s, r, eta)
    return out


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
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[0], [3], [4] legs")
    ap.add_argument("--sweep", action="store_true", help="run the FULL configs[4] sweep (100 materials x 34 models x 6 metrics) instead of the miniature")
    ap.add_argument("--sweep-steps", type=int, default=30)
    ap.add_argument("--sweep-materials", type=int, default=100)
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
    # affinity only in the multi-GPU runs: at N = 1 the CPU baseline of the same process must see every host core
    numa = bind_to_gpu_numa_node(local) if world > 1 else {"numa_node": None, "bound": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = bb.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    bsdf = bb.Bsdf(BSDF)
    n = 1 << args.log2_pairs
    peak, peak_src = measured_peak()

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

    def timed(step, steps, warm=3):
        """ms per step: CUDA events on the library's stream around `steps` calls, max over ranks"""
        for _ in range(warm):
            step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            step()
        b.record(stream)
        ctx.synchronize()
        return max_over_ranks(a.elapsed_time(b)) / steps

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
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    ctx.synchronize()
    launches = ctx.launches - l0
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9
    kernel_ms = e0.elapsed_time(e1) / args.steps            # one launch per step: the step IS the dominant kernel
    achieved = BYTES_PER_PAIR * n / (kernel_ms * 1e-3) / 1e9
    # clocks: the K timed steps above last ~15 ms, fewer than an NVML query on some boxes; a SECOND timed region of the same
    # step, long enough (>= 150 ms of kernels) to be sampled from inside, gives the clocks and a sustained rate beside them
    long_steps = max(args.steps, int(np.ceil(150.0 / max(kernel_ms, 1e-3))))
    with ClockSampler(local) as clocks:
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = len(clocks.sm)
        c0.record(stream)
        for _ in range(long_steps):
            step()
        c1.record(stream)
        ctx.synchronize()
        in_region = len(clocks.sm) - before
    clk = clocks.summary()
    sustained_ms = max_over_ranks(c0.elapsed_time(c1)) / long_steps
    clk.update({"samples_in_timed_region": in_region, "region": "%d further steps of the same kernel, timed with CUDA events (%.0f ms)" % (long_steps, sustained_ms * long_steps),
                "sustained_value": world * n / (sustained_ms * 1e-3) / 1e9})

    # ---- e2e: host buffers through the same public call -----------------------------------------------------------
    e2e, e2e_variants = None, None
    if not args.no_e2e:
        pin = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, pin_memory=True)
        h_out, h_xi = pin((3, n)), pin((2, n))
        h_out.copy_(out); h_xi.copy_(xi)
        h_res = (pin((3, n)), pin(n), pin(n, torch.int32), pin((3, n)), pin(n))
        a_out, a_xi = h_out.numpy(), h_xi.numpy()
        a_res = tuple(t.numpy() for t in h_res)

        def wall(fn, steps):
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()                                                 # returns after the D2H copies completed
            torch.cuda.synchronize()
            return max_over_ranks((time.perf_counter() - t0) * 1e3) / steps

        full_ms = wall(lambda: ctx.sample_eval_pdf(bsdf, a_out, a_xi, outputs=a_res), args.steps)
        e2e = {"value": world * n / (full_ms * 1e-3) / 1e9, "unit": "G pairs/s", "h2d_bytes_per_step": 20 * n, "d2h_bytes_per_step": 36 * n,
               "ms_per_step": full_ms, "contract": "host (pinned) inputs, all five outputs delivered to the caller's host buffers", "numa": numa,
               "d2h_bytes_over_pcie_per_step": 21 * n,
               "d2h_note": "36 B per pair arrive in the caller's buffers; 21 B of them cross the link: the flag plane travels as one byte per element, sample.pdf - the "
                           "same number as pdf for this model - is written by host threads from pdf, and eval = u x the lobe's RGB scale travels as u (4 B, not 12) with "
                           "the three IEEE products formed by host threads (bbmcu_api.cu compress_fused_outputs; every byte equals the device path's, checked below; "
                           "BBMCU_HOST_TRANSFER_PLAIN=1 sends all 36)"}
        # the host path returns exactly what the device path computed
        chk = slice(0, 1 << 16)
        for host_arr, dev_t in zip(a_res, outs):
            assert np.array_equal(np.ascontiguousarray(host_arr[..., chk]).view(np.uint32), dev_t[..., chk].cpu().numpy().view(np.uint32)), "host path differs from device path"
        e2e_variants = []
        two = (None, None, None, a_res[3], a_res[4])
        red_ms = wall(lambda: ctx.sample_eval_pdf(bsdf, a_out, a_xi, outputs=two), args.steps)
        e2e_variants.append({"contract": "host (pinned) inputs, rgb + pdf copied back (other output pointers NULL)", "value": world * n / (red_ms * 1e-3) / 1e9,
                             "unit": "G pairs/s", "h2d_bytes_per_step": 20 * n, "d2h_bytes_per_step": 16 * n, "ms_per_step": red_ms})
        gen_ms = wall(lambda: ctx.sample_eval_pdf_generated(bsdf, 1 + rank, 0, n, outputs=two), args.steps)
        e2e_variants.append({"contract": "inputs drawn on the device (Philox; bbmcu_sample_eval_pdf_generated), rgb + pdf copied back", "value": world * n / (gen_ms * 1e-3) / 1e9,
                             "unit": "G pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16 * n, "ms_per_step": gen_ms})
        assert np.array_equal(a_res[4][chk].view(np.uint32), ctx.sample_eval_pdf_generated(bsdf, 1 + rank, 0, 1 << 16, like=out, want=("pdf",))[4].cpu().numpy().view(np.uint32))
        if rank == 0 and world == 1:
            np_ = 1 << 24                                            # pageable caller memory (numpy), a quarter of a... 1/4 of the batch
            p_out, p_xi = np.array(a_out[:, :np_]), np.array(a_xi[:, :np_])
            p_res = (np.empty((3, np_), np.float32), np.empty(np_, np.float32), np.empty(np_, np.int32), np.empty((3, np_), np.float32), np.empty(np_, np.float32))
            pg_ms = wall(lambda: ctx.sample_eval_pdf(bsdf, p_out, p_xi, outputs=p_res), 5)
            e2e_variants.append({"contract": "PAGEABLE host inputs and outputs (internal pinned ring), all five outputs, 2^24 pairs per call", "value": np_ / (pg_ms * 1e-3) / 1e9,
                                 "unit": "G pairs/s", "h2d_bytes_per_step": 20 * np_, "d2h_bytes_per_step": 36 * np_, "ms_per_step": pg_ms})
            del p_out, p_xi, p_res
        del h_out, h_xi, h_res

    # ---- loss + gradient passes (BASELINE configs[2]), sample axis sharded over ranks --------------------
    loss_info = None
    if not args.no_loss:
        # the sample axis is dealt to the ranks in blocks of 1024 samples (rank, rank + world, ...): contiguous eighths of
        # the MERL grid are unequal work (the first has no pair below the horizon), interleaved shards are equal
        N = bb.MERL_BINS
        shard = dict(interleaved=(rank, world))
        fitted, truth = bb.Bsdf(FITTED), bb.Bsdf(TRUTH)
        L = ctx.loss("nganL2", truth, None, **shard)
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
                L_peer = ctx.loss("nganL2", truth, None, **shard)
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
            by_k[str(kk)] = kk / (timed(lambda: L.eval_device(fitted, pk, rk), 20) * 1e-3)       # this rank's shard only, no collective
        # the same pass reading MATERIALISED direction planes (36 B per sample instead of 12): what round 1 shipped
        Lm = ctx.loss("nganL2", truth, None, materialise=True, **shard)
        mat_ms = timed(lambda: Lm.eval_device(fitted, params, res), nl)
        del Lm
        # value only (what a derivative-free search such as bbm's compass asks for): this rank's shard, no collective
        resv = torch.zeros((1, LOSS_K, 1 + P), device=dev, dtype=torch.float64)
        val_ms = timed(lambda: L.eval_multi_device(fitted, params[None], resv, grad=False), nl)
        del resv
        # the same pass through the generic dual-number tile kernel (what every model without a compact kernel runs)
        os.environ["BBMCU_LOSS_NO_COMPACT"] = "1"
        try:
            gen_ms = timed(lambda: L.eval_device(fitted, params, res), nl)
        finally:
            del os.environ["BBMCU_LOSS_NO_COMPACT"]
        if world == 1:
            collective = None
        elif L_peer is not None:
            collective = "rows exchanged over NVLink peer memory inside the finish kernel (no NCCL call); K x (1+P) doubles per rank"
        else:
            collective = "nccl all_reduce of K x (1+P) doubles (peer windows unavailable: %s)" % peer_error
        loss_info = {"value": passes, "passes_per_s_by_K_no_collective": by_k, "unit": "loss+grad passes/s", "K": LOSS_K, "P": P, "samples_per_pass": N, "ms_per_step": step_ms,
                     "scaling": "strong", "metric": "nganL2", "fitted": FITTED, "collective": collective, "linearizer": "fused into the kernel (12 B per sample)",
                     "kernel": "compact pair kernel (bbmcu_losscompact.cuh): per-sample invariants, per-set constants, closed-form jacobian",
                     "passes_per_s_with_materialised_directions_no_collective": LOSS_K / (mat_ms * 1e-3),
                     "passes_per_s_generic_tile_kernel_no_collective": LOSS_K / (gen_ms * 1e-3),
                     "value_only_passes_per_s_no_collective": LOSS_K / (val_ms * 1e-3),
                     "passes_per_s_with_nccl_all_reduce": (LOSS_K / (nccl_ms * 1e-3)) if nccl_ms else None,
                     "effective_gbs_at_12B_per_sample": passes * 12 * N / 1e9, "frac_of_hbm_roofline": passes * 12 * N / 1e9 / (world * peak),
                     "frac_note": "whole-job passes/s x 12 B x N / (n_gpus x measured HBM peak); the grid is L2-resident, the kernel is issue-bound",
                     "gpu_launches": launches_loss, "loss0": loss0}

    # ---- configs[0]: Cook-Torrance eval over the MERL-grid directions -------------------------------------------------
    eval_info = None
    if not args.no_loss:
        N = bb.MERL_BINS
        gi, go = ctx.merl_dirs(0, N, like=out)
        ctx.synchronize()
        reps = 32                                                    # the 1 458 000 grid directions tiled 32x: working set 1.7 GB, far beyond L2
        gi, go = gi.repeat(1, reps).contiguous(), go.repeat(1, reps).contiguous()
        ne = N * reps
        rgb = torch.empty((3, ne), device=dev)
        ct = bb.Bsdf("CookTorrance()")
        torch.cuda.synchronize()
        ms_e = timed(lambda: ctx.eval(ct, gi, go, rgb=rgb), args.steps)
        ev_s = world * ne / (ms_e * 1e-3) / 1e9
        del gi, go
        # linearizer fused into the kernel: 0 B in, 12 B out per eval; 32 launches over the grid write 32 different output
        # slices so the stores go to DRAM like the materialised run's
        rgb1 = [torch.empty((3, N), device=dev) for _ in range(reps)]
        ms_f = timed(lambda: [ctx.eval_merl_grid(ct, 0, N, rgb=t) for t in rgb1], max(3, args.steps // 4)) / reps
        fused_s = world * N / (ms_f * 1e-3) / 1e9
        eval_info = {"value": ev_s, "unit": "G evals/s", "bsdf": "CookTorrance()", "evals_per_gpu": ne, "bytes_per_eval": 36,
                     "achieved_gbs_per_gpu": 36 * ne / (ms_e * 1e-3) / 1e9, "frac_of_hbm_roofline": 36 * ne / (ms_e * 1e-3) / 1e9 / peak,
                     "kernel": "k_foreach4<EvalOp<BsdfSingle<CookTorrance>>>",
                     "fused_linearizer": {"value": fused_s, "unit": "G evals/s", "bytes_per_eval": 12, "evals_per_launch": N,
                                          "achieved_gbs_per_gpu": 12 * N / (ms_f * 1e-3) / 1e9, "frac_of_hbm_roofline": 12 * N / (ms_f * 1e-3) / 1e9 / peak,
                                          "kernel": "k_foreach4<EvalGridOp<BsdfSingle<CookTorrance>>>", "note": "one launch per grid (1.458 M evals, ~8 us): launch-bound; bound by instruction issue, not HBM"}}
        del rgb, rgb1

    # ---- CPU baselines: the unmodified reference on this box's host cores (rank 0, N = 1 only) -------------
    cpu = None
    ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import refbind
            if refbind.available():
                ref = refbind.Ref("float")
                cores = host_threads()
                o_h, x_h = synth_host(n, 1)
                o_a, x_a = np.ascontiguousarray(o_h.T), np.ascontiguousarray(x_h.T)
                del o_h, x_h
                ref.sample_eval_pdf(BSDF, o_a[:1 << 18], x_a[:1 << 18], threads=cores)                       # warm
                dts = []
                for _ in range(2):
                    t0 = time.perf_counter()
                    ref.sample_eval_pdf(BSDF, o_a, x_a, threads=cores)
                    dts.append(time.perf_counter() - t0)
                dt = min(dts)
                cpu = {"value": n / dt / 1e9, "unit": "G pairs/s", "cores": cores, "kind": "reference",
                       "sample": f"{n} pairs (1.0000 of one step), best of 2, oracle/_ref = unmodified reference native backbone floatRGB (-O3 -DNDEBUG, no -march=native), {cores} threads"}
                del o_a, x_a
        except Exception as e:   # the baseline is reported, never required
            cpu = {"value": None, "unit": "G pairs/s", "cores": host_threads(), "kind": "reference", "sample": f"failed: {e}"}
    if ref is not None and eval_info is not None:
        # configs[0], CPU side: Cook-Torrance eval over all 1 458 000 grid directions through bsdf_ptr (checkBsdf-style,
        # bin/checkBsdf.cpp:123-141), 1 thread exactly as shipped and all host threads
        gi_h, go_h = ctx.merl_dirs(0, bb.MERL_BINS)
        ai, ao = np.ascontiguousarray(gi_h.T), np.ascontiguousarray(go_h.T)
        cores = host_threads()
        def best(threads, reps_):
            b = 1e30
            for _ in range(reps_):
                t0 = time.perf_counter(); ref.eval("CookTorrance()", ai, ao, threads=threads); b = min(b, time.perf_counter() - t0)
            return b
        t1, tall = best(1, 3), best(cores, 5)
        eval_info["cpu_reference"] = {"unit": "G evals/s", "evals": bb.MERL_BINS, "one_thread": bb.MERL_BINS / t1 / 1e9, "all_threads": bb.MERL_BINS / tall / 1e9, "cores": cores,
                                      "how": "oracle/_ref: bsdf_import('CookTorrance()') -> bsdf_ptr::eval per direction pair, best of 3 / 5"}

    # ---- configs[3]: EPD ("Holzschuch-Pacanowski") eval over the grid and a fit, tables staged on chip ---------------------------------------
    epd_info = None
    if not args.no_extras and not args.no_loss:
        N = bb.MERL_BINS
        epd = bb.Bsdf("EPD(0.05, 0.5, [1.5, 0.5])")
        reps = 8
        gi, go = ctx.merl_dirs(0, N, like=out)
        gi, go = gi.repeat(1, reps).contiguous(), go.repeat(1, reps).contiguous()
        rgb = torch.empty((3, N * reps), device=dev)
        ms_e = timed(lambda: ctx.eval(epd, gi, go, rgb=rgb), max(3, args.steps // 2))
        del gi, go, rgb
        grid = bb.spherical_grid((31, 16), (1, 9))
        epd_truth, epd_start = "EPD(0.05, 0.8, [1.5, 0.5])", "EPD(0.1, 0.5, [1.3, 0.2])"
        from bbm_b200 import fit as bfit
        Le = ctx.loss("standardLog", bb.Bsdf(epd_truth), grid)
        fb = bb.Bsdf(epd_start)
        ctx.synchronize()
        t0 = time.perf_counter()
        opt = bfit.CompassBatched(Le, fb, fb.parameter_lower_bound(), fb.parameter_upper_bound())
        nsteps = 0
        for _ in range(60):
            if opt.is_converged():
                break
            opt.step(); nsteps += 1
        fit_s = time.perf_counter() - t0
        epd_info = {"eval": {"value": world * N * reps / (ms_e * 1e-3) / 1e9, "unit": "G evals/s", "bsdf": "EPD(0.05, 0.5, [1.5, 0.5])", "evals_per_gpu": N * reps,
                             "note": "G1 rows of the launch-uniform p staged in shared memory (2 x 1000 floats)"},
                    "fit": {"steps": nsteps, "seconds": fit_s, "steps_per_s": nsteps / fit_s, "samples": 31 * 16 * 9, "P": 4, "final_loss": float(opt.loss_value),
                            "truth": epd_truth, "start": epd_start, "metric": "standardLog", "optimizer": "compass, all 2P probes of a step in one launch"}}
        if ref is not None:
            from oracle.refbind import sph_desc
            gi_h, go_h = ctx.merl_dirs(0, N)
            sub = slice(0, 1 << 18)
            ai, ao = np.ascontiguousarray(gi_h.T[sub]), np.ascontiguousarray(go_h.T[sub])
            cores = host_threads()
            t0 = time.perf_counter(); ref.eval("EPD(0.05, 0.5, [1.5, 0.5])", ai, ao, threads=cores); te = time.perf_counter() - t0
            t0 = time.perf_counter()
            trace, _, _ = ref.compass("standardLog", sph_desc((31, 16), (1, 9)), epd_start, epd_truth, 10)
            tf = time.perf_counter() - t0
            epd_info["cpu_reference"] = {"eval_G_evals_per_s_all_threads": (1 << 18) / te / 1e9, "cores": cores, "fit_steps_per_s_one_thread": len(trace) / tf,
                                         "fit_steps_timed": int(len(trace)), "how": "oracle/_ref: bsdf_ptr eval on 2^18 grid directions; unmodified bbm::compass, 10 steps"}

    # ---- checkBsdf's sample loops and the HP renormalisation table as device kernels (rank 0's GPU) -----------------
    tools_info = None
    if not args.no_extras and rank == 0:
        from bbm_b200 import check as bcheck
        ggx = bb.Bsdf("GGX()")
        ns = 1 << 26
        bcheck.pdf_integral(ctx, ggx, 1 << 20, 1, False, "philox", 1)                      # warm-up
        t0 = time.perf_counter(); val, _ = bcheck.pdf_integral(ctx, ggx, ns, 4, False, "philox", 2); t_int = time.perf_counter() - t0
        t0 = time.perf_counter(); rp = bcheck.pdf(ctx, ggx, ns, False, True, "philox", 3); t_pdf = time.perf_counter() - t0
        t0 = time.perf_counter(); tab = ctx.hp_precompute_normalization(); t_hp = time.perf_counter() - t0
        tools_info = {"checkBsdf": {"bsdf": "GGX()", "rng": "philox (drawn in the kernel, 0 bytes in)", "samples_per_estimate": ns,
                                    "pdfInt_4_trials_s": t_int, "pdfInt_G_samples_per_s": 4 * ns / t_int / 1e9, "pdfInt_values": [float(v) for v in val],
                                    "pdf_test_s": t_pdf, "pdf_test_G_samples_per_s": ns / t_pdf / 1e9, "pdf_test_negative": list(rp["negative"]), "pdf_test_mismatch": [float(v) for v in rp["mismatch"]]},
                      "hp_normalization": {"entries": int(tab.size), "seconds": t_hp, "quadrature_terms": 5.67e9}}
        del tab

    # ---- configs[4]: many materials per launch, and the fit sweep ---------------------------------------------------------------------------
    multi_info, sweep_info = None, None
    if not args.no_extras and not args.no_loss:
        from bbm_b200 import fit as bfit
        n_mat_total = args.sweep_materials if args.sweep else 8 * world
        mats = synthetic_materials(n_mat_total)
        names = sorted(mats)
        from bbm_b200.shard import shard_range
        f0_, c0_ = shard_range(len(names), rank, world)
        mine = names[f0_:f0_ + c0_]
        tables = {m: ctx.eval_merl_grid(bb.Bsdf(mats[m]), like=out) for m in mine}          # resident on the device, 17.5 MB each
        ctx.synchronize()
        Lmulti = ctx.loss("nganL2", [tables[m] for m in mine], None)
        fitted = bb.Bsdf(FITTED)
        P = len(fitted.parameter_values())
        # (a) a true DRAM workload: MD = 64 materials (1.12 GB of tabulated data >> the 126 MB L2) x K parameter sets in one
        # launch; K = 1 value-only streams every material once per pass
        MD = 64
        big = synthetic_materials(MD, seed=77)
        tabs = [ctx.eval_merl_grid(bb.Bsdf(v), like=out) for _, v in sorted(big.items())]
        ctx.synchronize()
        Lbig = ctx.loss("nganL2", tabs, None)
        del tabs
        p1 = np.tile(fitted.parameter_values(), (MD, 1, 1))
        p16 = np.tile(p1, (1, 16, 1))
        res1 = torch.zeros((MD, 1, 1 + P), device=dev, dtype=torch.float64)
        res16 = torch.zeros((MD, 16, 1 + P), device=dev, dtype=torch.float64)
        ms_v = timed(lambda: Lbig.eval_multi_device(fitted, p1, res1, grad=False), max(5, args.steps // 2))
        ms_g1 = timed(lambda: Lbig.eval_multi_device(fitted, p1, res1), max(5, args.steps // 2))
        ms_g = timed(lambda: Lbig.eval_multi_device(fitted, p16, res16), max(3, args.steps // 4))
        bytes_v = 12 * bb.MERL_BINS * MD
        gbs = lambda ms_: bytes_v / (ms_ * 1e-3) / 1e9
        multi_info = {"materials_per_launch": MD, "bytes_of_tabulated_data_per_gpu": bytes_v,
                      "value_only_K1": {"passes_per_s": world * MD / (ms_v * 1e-3), "ms_per_launch": ms_v, "achieved_gbs_per_gpu": gbs(ms_v), "frac_of_hbm_roofline": gbs(ms_v) / peak},
                      "loss_grad_K1": {"passes_per_s": world * MD / (ms_g1 * 1e-3), "ms_per_launch": ms_g1, "achieved_gbs_per_gpu": gbs(ms_g1), "frac_of_hbm_roofline": gbs(ms_g1) / peak},
                      "loss_grad_K16": {"passes_per_s": world * MD * 16 / (ms_g * 1e-3), "ms_per_launch": ms_g},
                      "note": "one launch = 64 materials x K parameter sets (a block's direction-only work serves all the materials of its group), 12 B x 1 458 000 per material read once per launch: "
                              "1.12 GB per GPU, far beyond the 126 MB L2 - the true-DRAM case of SURVEY 8(d); each rank holds its own 64 materials (weak scaling)"}
        del Lbig, res1, res16
        # (b) the fit sweep, split by material over the ranks; no collective
        models = [m for m in bb.model_names() if m != "Merl"] if args.sweep else SWEEP_MODELS_MINI
        metrics = bb.METRICS if args.sweep else ["nganL2", "standardLog"]
        max_steps = args.sweep_steps if args.sweep else 10
        per_model = {}
        def progress(mod, met, secs, steps, nm, k):
            per_model.setdefault(mod, [0.0, 0])
            per_model[mod][0] += secs; per_model[mod][1] += steps * nm * k
        barrier()
        t0 = time.perf_counter()
        result = bfit.run_sweep_by_material(ctx, tables, models, metrics, 0, 1, max_steps, None, loss=Lmulti, progress=progress)      # this rank's materials only
        ctx.synchronize()
        mine_s = time.perf_counter() - t0
        total_s = max_over_ranks(mine_s * 1e3) / 1e3
        fits = len(models) * len(metrics) * n_mat_total
        finite = sum(1 for v in result.values() if np.isfinite(v[1]))
        if world > 1:
            times = [None] * world
            dist.all_gather_object(times, mine_s)
        else:
            times = [mine_s]
        sweep_info = {"kind": "full (BASELINE configs[4])" if args.sweep else "miniature of configs[4] (bench.py --sweep runs 100 materials x 34 models x 6 metrics)",
                      "materials": n_mat_total, "models": len(models), "metrics": len(metrics), "fits": fits, "compass_steps_per_fit": max_steps,
                      "seconds": total_s, "fits_per_s": fits / total_s, "rank_seconds": [round(float(t), 3) for t in times], "finite_fits_this_rank": finite, "fits_this_rank": len(result),
                      "partition": "by material (each rank keeps its materials' tables resident as one batched loss; one launch per compass step for all of them); no collective",
                      "seconds_by_model_this_rank": {k: round(v[0], 3) for k, v in sorted(per_model.items(), key=lambda kv: -kv[1][0])[:8]},
                      "loss_passes_this_rank": int(sum(v[1] for v in per_model.values())), "data": "synthetic Lambertian + LowCookTorrance materials (seeded), tabulated on the MERL grid"}
        if args.sweep and rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump({"sweep": sweep_info, "seconds_by_model_rank0": {k: v[0] for k, v in per_model.items()},
                       "sample_results": {"|".join(k): v for k, v in list(result.items())[:40]}}, open(os.path.join(ROOT, "gpurun_out", "sweep_full.json"), "w"), indent=1)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "G pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.log2_pairs),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(n), "peak_source": peak_src,
                             "kernel": "k_foreach4<SampleEvalPdfOpT<BsdfSingle<GGX>, false>>", "algorithmic_bytes_per_launch": BYTES_PER_PAIR * n},
                "cpu_baseline": cpu, "e2e": e2e, "e2e_variants": e2e_variants, "gpu_launches": launches, "clocks": clk, "loss_grad": loss_info, "eval_merl_grid": eval_info,
                "epd": epd_info, "loss_multi": multi_info, "sweep": sweep_info, "tools": tools_info}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
