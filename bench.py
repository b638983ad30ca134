#!/usr/bin/env python
"""Benchmark of the DFMI readout hot path (BASELINE.json: NLS fit buffers/s; demod HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE config 2, "long single channel" -- f_mod = 1 kHz, f_samp = 1 MHz,
3600 s of synthetic 'snr'-mode DFMI (m = 6, 40 dB), n = 20 periods per buffer (R = 20000 samples),
N = 10 harmonics: 180 000 buffers = 3.6e9 samples = 28.8 GB of fp64 per GPU.  A step is one whole NLS
readout of the record (demodulate + fit every buffer).  With N > 1 GPUs (torchrun, one process per
GPU) every rank holds its own 3600 s slab (weak scaling): the path has no exchange step, so ranks share
nothing but the barrier around the timed region.

value      buffers/s over all ranks, record resident in HBM (28.8 GB >> 126 MB L2: no flush needed)
e2e        the same through the reference-facing host-pointer call (dfk_nls_fit_host behind
           StandardNLSFitter.fit): pinned host record -> H2D slabs -> kernels -> D2H rows, all timed
roofline   the demodulation kernel: algorithmic bytes (8 R + 8 (2N+1) per buffer) / its CUDA-event time
cpu_baseline  the oracle port of the reference's multiprocessing schedule on the host cores, bounded sample
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F_SAMP, F_MOD, N_CYCLES, NDATA = 1e6, 1000.0, 20, 10
SECONDS = 3600.0
M_TRUE, SNR_DB = 6.0, 40.0
R = int(F_SAMP / F_MOD * N_CYCLES)
NBUF = int(SECONDS * F_SAMP) // R
INIT = [1.6, 6.0, 0.0, 0.0]
METRIC, UNIT = "nls_fit_buffers_per_sec", "buffers/s"
WORKLOAD = ("cfg2 long single channel: f_mod=1kHz f_samp=1MHz 3600s synthetic DFMI (m=6, SNR 40dB), "
            "n=20 (R=20000), ndata=10, NLS fit per buffer")


def config_dict(n_gpus):
    return {"workload": WORKLOAD, "buffers_per_gpu": NBUF, "samples_per_gpu": NBUF * R, "R": R, "ndata": NDATA,
            "record_bytes_per_gpu": NBUF * R * 8, "sharding": f"{n_gpus} contiguous time slabs, one per GPU, no collective",
            "l2": "inputs (28.8 GB per GPU) exceed the 126 MB L2; no flush between steps"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per demod launch from the committed ncu capture of this command, if there is one."""
    path = os.path.join(ROOT, "profiles", "demod_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def lm_roofline(counters, fits, nbuf, lm_ms, peak_tflops):
    """FP64 rate of the LM launches from the work the kernels counted (SURVEY 8d's per-call costs, N harmonics):
    model+Jacobian F_c = 3M + 90N + 2S, residual F_s = 3M + 12N + 2S with the Miller steps M counted on device,
    S = 60 flop per sincos, F_m = 100 per damped solve."""
    per = {k: v / fits for k, v in counters.items()}
    flop_per_fit = (per["n_state"] * (90 * NDATA + 120) + per["n_ssq"] * (12 * NDATA + 120) + per["n_solve"] * 100 +
                    3 * per["n_bessel_steps"])
    achieved = flop_per_fit * nbuf / (lm_ms * 1e-3) / 1e12
    return {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": achieved / peak_tflops if peak_tflops else None, "flop_per_fit": flop_per_fit,
            "peak_source": "dfk_probe_fp64: DFMA chains measured on this GPU in this run",
            "note": "algorithmic flops of the reference's formulation; the kernel's block-diagonal normal equations "
                    "execute about half of the model+Jacobian figure"}


# ---- clocks sampled during the timed region -------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---- CPU arm: the oracle port of the reference's multiprocessing path -----------------------------------
def cpu_sample(n_buffers, seed):
    from oracle import dfmi_oracle as orc
    return orc.snr_signal(M_TRUE, F_SAMP, F_MOD, n_buffers * R / F_SAMP, SNR_DB, seed=seed)


def time_cpu_pool(x, cores):
    from oracle import dfmi_oracle as orc
    t0 = time.perf_counter()
    rows = orc.nls_fit_pool(x, F_SAMP, F_MOD, N_CYCLES, NDATA, n_procs=cores)
    dt = time.perf_counter() - t0
    assert rows.shape[0] == len(x) // R
    return dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- CPU leg for the other BASELINE configs (bounded samples; `--impl reference --workload cfgN`) -------------------
def _cfg5_job(job):
    from oracle import dfmi_oracle as orc
    m, seed = job
    x = orc.snr_signal(float(m), 200e3, 1000.0, 1e-3, 40.0, seed=seed)
    return orc.nls_fit(x, 200e3, 1000.0, 1, 15, init_m=float(m))[0]


def run_reference_workload(args):
    """The oracle port of the reference on a bounded sample of cfg 1, 3, 4 or 5 (cfg 2 is the default arm)."""
    from multiprocessing import Pool
    from oracle import dfmi_oracle as orc
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    cores = host_cores()
    cfg = args.workload
    if cfg in ("cfg1", "cfg3"):
        secs = 10.0 if cfg == "cfg1" else 20.0
        x = orc.snr_signal(6.0, 200e3, 1000.0, secs, 40.0, seed=0)
        t0 = time.perf_counter()
        rows = orc.nls_fit_pool(x, 200e3, 1000.0, 20, 10, n_procs=cores)
        dt = time.perf_counter() - t0
        out = {"metric": METRIC, "value": len(rows) / dt, "unit": UNIT, "cores": cores,
               "sample": f"1 channel x {secs:.0f} s ({len(rows)} buffers of 4000 samples), Pool schedule, start-up included"}
    elif cfg == "cfg5":
        jobs = [(m, s) for m in range(2, 21) for s in range(100)]
        t0 = time.perf_counter()
        with Pool(cores) as pool:
            rows = pool.map(_cfg5_job, jobs, chunksize=25)
        dt = time.perf_counter() - t0
        out = {"metric": "single_buffer_fits_per_sec", "value": len(rows) / dt, "unit": "fits/s", "cores": cores,
               "sample": f"{len(jobs)} realisations (100 per m in 2..20), Pool over realisations, signal synthesis "
                         "included as in workers.py:132-189"}
    elif cfg == "cfg4":
        x = orc.snr_signal(6.0, 200e3, 1000.0, 0.1, 40.0, seed=0)
        t0 = time.perf_counter()
        orc.ekf_track(x, 200e3, 1000.0, 20)
        dt = time.perf_counter() - t0
        out = {"metric": "ekf_samples_per_sec", "value": len(x) / dt, "unit": "samples/s", "cores": 1,
               "sample": "1 channel x 0.1 s (20000 steps) on one core; channels are independent, a box scales this by "
                         f"its core count ({cores} here)"}
    else:
        raise SystemExit(f"unknown workload {cfg}")
    out.update({"impl": "reference", "kind": "port", "workload": cfg})
    emit(out)
    return 0


def run_reference(args):
    """--impl reference: the reference's CPU schedule (fitters.py:395-428: buffer 0, then a process pool over
    chunks of the rest) restated by the oracle, on all host cores, each step a bounded sample of cfg 2."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    # size the per-step sample so that the whole K + W run stays within ~2 minutes
    cal = cpu_sample(200, seed=1)
    rate = 200 / time_cpu_pool(cal, cores)
    budget_s = min(10.0, 100.0 / (args.steps + args.warmup))
    n_buffers = int(max(200, min(4000, rate * budget_s)))
    x = cpu_sample(n_buffers, seed=1)
    for _ in range(args.warmup):
        time_cpu_pool(x, cores)
    times = [time_cpu_pool(x, cores) for _ in range(args.steps)]
    dt = sum(times) / len(times)
    value = n_buffers / dt
    sample = f"{n_buffers} buffers ({n_buffers * R / F_SAMP:.0f} s of the cfg2 record) per step, pool start-up included"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "samples_per_sec": value * R, "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def bind_to_gpu_numa_node(device):
    """Pin this rank to the CPUs of its GPU's NUMA node before it allocates pinned host memory, so that the e2e leg's
    host record is first-touched on the socket the GPU hangs off (matters once several ranks stream at once)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


# ---- GPU arm --------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from deepfmkit_b200 import _lib
    from deepfmkit_b200 import fit as tunables

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = _lib.Context(local, own_stream=True)
    stream = torch.cuda.Stream(device=local)
    w0 = 2.0 * np.pi * F_MOD / F_SAMP
    opts = tunables.current_lm_opts()

    with torch.cuda.stream(stream):
        x = torch.empty(NBUF * R, dtype=torch.float64, device="cuda")
        rows = torch.empty((NBUF, _lib.ROW_STRIDE), dtype=torch.float64, device="cuda")
        ctx.use_torch_stream(stream)
        ctx.synth_snr_dev(x.data_ptr(), NBUF * R, 1, F_SAMP, F_MOD, M_TRUE, snr_db=SNR_DB, seed=1000 + rank)

        def step():
            ctx.nls_fit_dev(x.data_ptr(), NBUF, R, NDATA, w0, INIT, True, opts, rows.data_ptr())

        for _ in range(args.warmup):
            step()
        stream.synchronize()
        ctx.lm_counters(reset=True)
        ctx.profile_enable(True)
        ctx.profile_read(reset=True)
        launches0 = ctx.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with ClockSampler(local) as clocks:
            ev0.record(stream)
            for _ in range(args.steps):
                step()
            ev1.record(stream)
            stream.synchronize()
        barrier()
        elapsed_ms = ev0.elapsed_time(ev1)
        launches = ctx.launch_count() - launches0
        prof = ctx.profile_read(reset=True)
        ctx.profile_enable(False)
        counters = ctx.lm_counters(reset=True)
        head = rows[:4096].cpu().numpy()
        fp64_peak = ctx.probe_fp64_tflops() if rank == 0 else None

    # sanity: the timed work produced real fits
    assert np.all(head[:, 6] == 0) and abs(head[:, 1].mean() - M_TRUE) < 1e-3, "bench fits are wrong"

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    ms_per_step = max_ms / args.steps
    value = world * NBUF / (ms_per_step * 1e-3)

    # ---- end to end through the host-pointer entry (what StandardNLSFitter.fit calls) ----------------------
    import psutil
    # every rank pins its own host copy; keep the sum under 45% of the box's RAM (same answer on every rank)
    host_total = psutil.virtual_memory().total
    e2e_nbuf = NBUF
    while e2e_nbuf * R * 8 * world > 0.45 * host_total and e2e_nbuf > 1000:
        e2e_nbuf //= 2
    xh = torch.empty(e2e_nbuf * R, dtype=torch.float64, pin_memory=True)
    xh.copy_(x[: e2e_nbuf * R])
    torch.cuda.synchronize()
    xh_np = xh.numpy()
    rows_h = torch.empty((e2e_nbuf, _lib.ROW_STRIDE), dtype=torch.float64, pin_memory=True).numpy()
    ctx.use_own_stream()
    e2e_steps = max(1, min(args.steps, 3))
    ctx.nls_fit_host(xh_np, R, NDATA, w0, INIT, seeded=True, opts=opts, rows_out=rows_h)  # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.nls_fit_host(xh_np, R, NDATA, w0, INIT, seeded=True, opts=opts, rows_out=rows_h)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_nbuf / float(te.item())
    assert np.array_equal(rows_h[:1024, 6], head[:1024, 6])
    # the same call on ordinary (pageable) memory -- what a numpy array out of pandas is -- on a tenth of the record
    pg_nbuf = max(1000, e2e_nbuf // 10)
    x_pg = np.array(xh_np[: pg_nbuf * R])
    ctx.nls_fit_host(x_pg, R, NDATA, w0, INIT, seeded=True, opts=opts)
    t0 = time.perf_counter()
    ctx.nls_fit_host(x_pg, R, NDATA, w0, INIT, seeded=True, opts=opts)
    pg_value = pg_nbuf / (time.perf_counter() - t0)
    del xh, xh_np, x_pg

    if rank == 0:
        peak, peak_src = measured_peak()
        demod_ms = prof["demod_ms"] / max(prof["demod_regions"], 1)
        lm_ms = prof["lm_ms"] / max(prof["lm_regions"], 1)
        alg_bytes = (8 * R + 8 * (2 * NDATA + 1)) * NBUF
        achieved = alg_bytes / (demod_ms * 1e-3) / 1e9
        fits = max(args.steps * NBUF, 1)
        cores = host_cores()
        n_cpu = max(200, min(2000, 25 * cores))
        xc = cpu_sample(n_cpu, seed=1)
        cpu_dt = time_cpu_pool(xc, cores)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "samples_per_sec": value * R, "config": config_dict(world),
            "roofline": {"bound": "hbm", "kernel": "demod_fold_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_per_launch(),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": demod_ms, "share_of_step": demod_ms / ms_per_step},
            "lm": {"kernel_ms_per_step": lm_ms, "share_of_step": lm_ms / ms_per_step,
                   "fits_per_sec": NBUF / (lm_ms * 1e-3),
                   "per_fit": {k: v / fits for k, v in counters.items()},
                   "roofline": lm_roofline(counters, fits, NBUF, lm_ms, fp64_peak)},
            "cpu_baseline": {"value": n_cpu / cpu_dt, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} buffers ({n_cpu * R / F_SAMP:.0f} s of the cfg2 record), "
                                       "oracle port of the reference's Pool schedule, pool start-up included"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_nbuf * R * 8,
                    "d2h_bytes_per_step": e2e_nbuf * _lib.ROW_STRIDE * 8, "ms_per_step": float(te.item()) * 1e3,
                    "buffers_per_gpu": e2e_nbuf, "api": "dfk_nls_fit_host (StandardNLSFitter.fit), pinned host record",
                    "rank0_numa_node": numa_node,
                    "pageable_input": {"value": pg_value, "unit": UNIT, "buffers": pg_nbuf,
                                       "note": "rank 0, unpinned numpy input staged by copy threads through pinned buffers"}},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else any library prints was sent to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    # NCCL and friends print banners on fd 1; keep stdout clean for the JSON line
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="reference arm only: which BASELINE config's bounded CPU sample to time (default: the contract's cfg2)")
    args = ap.parse_args()
    if args.steps < 1 or args.warmup < 0:
        raise SystemExit("--steps must be >= 1 and --warmup >= 0")
    if args.impl == "reference":
        return run_reference(args) if args.workload == "cfg2" else run_reference_workload(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
