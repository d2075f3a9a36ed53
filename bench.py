#!/usr/bin/env python
"""bench.py -- headline benchmark of the qampy_b200 hot path (BASELINE.json metric).

Workload (``config.workload``): BASELINE config C3 per GPU -- dual-pol 64-QAM, 2 samples/symbol,
``--nsym`` (default 1e7) symbols, MCMA -> MRDE training (ntaps 45, mu 1e-3) -> apply -> BPS(64 test
angles, N = 45), processed as independent time segments of ``--seg`` output symbols (ntaps-1 overlap;
segment s == the reference's dual_mode_equalisation(segment, wxy=start taps) + bps).  Start taps
(``--start``): carried from capture to capture like a running receiver does (default; acquired once on
2^18 symbols of the first capture, outside the timed steps, reported as ``acquisition``), acquired inside
every step, or centre spike (does not converge on 8 k-symbol segments).  The line is withheld (value null)
unless the symbol error rate over ALL segments of the last timed step is below 1e-5.
One "step" = one pass of the whole chain over one capture, both training passes over every sample,
training errors returned like the reference does.  N > 1: every rank owns its own capture
(weak scaling, no collective on the data path); the reported value is the sum over ranks divided by
the max-over-ranks device time.

    python bench.py [--gpus N --steps K --warmup W]           # this project's CUDA path
    python bench.py --impl reference [...]                     # reference-equivalent CPU path (oracle port)

Prints ONE JSON line (see the task contract): value = Msamples/s with inputs resident in HBM,
e2e = same metric from pinned host buffers through H2D / D2H, roofline for the dominant kernel
(train), cpu_baseline = the oracle (reference's compile flags) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Msamples/s dual-pol 64-QAM MCMA->MRDE->BPS"
# algorithmic bytes per symbol period (dual-pol, os=2, complex64), SURVEY.md section 8d / DESIGN.md
BYTES_TRAIN, BYTES_APPLY, BYTES_BPS = 48, 48, 40
# dram__bytes_read.sum + dram__bytes_write.sum per launch at C3 from the ncu --set full captures (profiles/README.md)
TRAFFIC_TRAIN, TRAFFIC_BPS = 326.1e6, 352.3e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"],
                    help="c3 (default, the contract's N=1 workload): --nsym symbols per GPU, weak scaling; c5: BASELINE "
                         "config 5, ONE capture of 5e8 symbols per polarisation (1e9 samples) cut into contiguous "
                         "per-rank ranges of whole segments (strong scaling, no collective)")
    ap.add_argument("--nsym", type=int, default=10 ** 7, help="symbols per polarisation per GPU")
    ap.add_argument("--seg", type=int, default=-1,
                    help="output symbols per segment; -1 = smallest length >= 4096 that fills whole GPU waves of two training "
                         "warps per SM sub-partition "
                         "(pipeline.balanced_segment_symbols), 0 = one segment")
    ap.add_argument("--start", default="stream", choices=["stream", "acquire", "cold"],
                    help="what a segment's taps start from.  stream (default): the taps the receiver carries from capture "
                         "to capture -- acquired ONCE per link on the head of the first capture (2 x --acq symbols, one "
                         "serial stream per mode, before the warm-up steps; timed and reported as `acquisition`), then "
                         "every step starts its segments from the taps of the previous capture's last segment.  "
                         "acquire: every step first acquires on the head of ITS capture (inside the timed region).  "
                         "cold: centre-spike taps for every segment (round 1; does not converge: SER 3e-4)")
    ap.add_argument("--acq", type=int, default=1 << 18, help="acquisition training symbols per stage")
    ap.add_argument("--acq-layout", default="latency", choices=["latency", "throughput"])
    ap.add_argument("--no-err", action="store_true", help="do not produce / download the per-symbol training errors "
                                                          "(the reference always returns them)")
    ap.add_argument("--chunks", type=int, default=8, help="host<->device overlap chunks of the e2e path (the last one "
                                                          "is cut into 1/2 + 1/4 + 1/4 unless --no-taper)")
    ap.add_argument("--ntaps", type=int, default=45)
    ap.add_argument("--M", type=int, default=64)
    ap.add_argument("--angles", type=int, default=64)
    ap.add_argument("--bpsN", type=int, default=45)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays of the step")
    ap.add_argument("--no-dropin", action="store_true", help="skip the single-capture drop-in records (script shapes)")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 sub-record (one 1e9-sample capture over the ranks)")
    ap.add_argument("--no-pin", action="store_true", help="e2e path: leave the process on whatever CPUs it was given "
                                                          "instead of binding it to the CPUs next to its GPU")
    ap.add_argument("--no-taper", action="store_true", help="e2e path: do not cut the last chunk into 1/2 + 1/4 + 1/4")
    return ap.parse_args()


def workload_config(a, world):
    name = "C3" if a.workload == "c3" else "C5 (1e9-sample capture over %d GPU(s))" % world
    return {"workload": "%s dual-pol %d-QAM dual_mode_equalisation(mcma->mrde, ntaps=%d, mu=1e-3) + bps(%d, N=%d), "
                        "2 sps, %d symbols per GPU" % (name, a.M, a.ntaps, a.angles, a.bpsN, a.nsym),
            "symbols_per_gpu": a.nsym, "samples_per_gpu": 2 * a.nsym, "segment_symbols": a.seg or a.nsym,
            "segment_semantics": "each segment == reference call dual_mode_equalisation(segment, wxy=start taps); its taps "
                                 "filter the segment plus %d symbols on either side, bps runs on that and the segment keeps "
                                 "its own symbols (the reference's bps gives the first/last N symbols of a call no "
                                 "estimate)" % (a.bpsN if a.seg else 0),
            "start": {"stream": "taps carried from capture to capture (acquired once per link on 2 x %d symbols before "
                                "the warm-up; steps alternate between two captures of the same link)" % a.acq,
                      "acquire": "every step acquires taps on the head of its capture (2 x %d symbols, inside the timed "
                                 "region), then all segments start from them" % a.acq,
                      "cold": "centre-spike taps for every segment"}[a.start],
            "returns": "recovered symbols, phase" + ("" if a.no_err else ", err1, err2 (per-symbol training errors of both stages)"),
            "sharding": ("dp%d (independent captures per rank, no collective)" % world) if a.workload == "c3" else
                        ("dp%d (every rank synthesises and processes its own contiguous range of %d symbols of the "
                         "5e8-symbol capture; whole segments, no collective)" % (world, a.nsym)),
            "l2_policy": "inputs (%d MB/GPU) larger than L2 (126 MB); no explicit flush" % (a.nsym * 32 // 10 ** 6)}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
def pin_to_gpu_cpus(sampler):
    """Run this process on the CPUs next to its GPU (NVML's CPU affinity of the device) so that the pinned host buffers
    of the e2e path are allocated on the GPU's NUMA node: a capture that crosses the socket link moves at 77 instead of
    98 GB/s (both directions, one GPU).  Returns (previous affinity, CPUs chosen) or (None, None)."""
    try:
        nvml, handle = sampler.nvml, sampler.handle
        if nvml is None or handle is None or not hasattr(os, "sched_setaffinity"):
            return None, None
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = nvml.nvmlDeviceGetCpuAffinity(handle, words)
        near = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        before = os.sched_getaffinity(0)
        cpus = sorted(near & before)
        if not cpus or set(cpus) == set(before):
            return None, None
        os.sched_setaffinity(0, cpus)
        return before, cpus
    except Exception:
        return None, None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The region lasts tens of milliseconds, so
    the sampler polls NVML in a thread (every 2 ms) instead of `nvidia-smi -lms`, whose first line arrives
    after the region has ended; nvidia-smi is the fallback when pynvml is missing."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            # the first queries of a process take tens of milliseconds inside the driver and can hold up kernel
            # launches (scratch/step_probe.py): pay for them here, before the timed region starts
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.samples.append((float(sm), int(mask)))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.nvml is None:
            return self._smi_once()
        self.stop_flag.set()
        self.thread.join(timeout=2)
        n = self.nvml
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:
            mx = None
        sm = sorted(v for v, _ in self.samples)
        mask = 0
        for _, m in self.samples:
            mask |= m
        reasons = [name for name, bit in self.REASONS if mask & bit]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sm), "source": "nvml polled during the timed region"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                    "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi right after the timed region (pynvml unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port compiled with the reference's flags, on the host cores
# ------------------------------------------------------------------------------------------------
_CPU_STATE = {}          # taps carried from step to step ("stream"), like the GPU arm


def cpu_chain(a, nseg, seed=1234):
    """Times the reference-equivalent CPU chain on `nseg` segments of the workload, same recipe as the GPU arm
    (--start): "stream" = segments start from the taps carried over from the previous step (the first call acquires
    them on the head of its capture, untimed, like the GPU arm's warm-up), "acquire" = acquisition timed with the
    step, "cold" = centre-spike taps per segment.  Returns (seconds for the segments, seconds for the acquisition or
    0, samples processed, threads, symbol errors, symbols compared)."""
    # all host threads: torchrun exports OMP_NUM_THREADS=1 to its workers, which would cripple the CPU arm
    if "cpu_oracle" not in sys.modules:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_oracle as co
    from qampy_b200 import synth, theory
    S = a.seg or a.nsym
    H = a.bpsN if a.seg else 0          # phase-search halo (pipeline.ReceiverConfig.bps_halo), as in the GPU arm
    need_acq = a.start == "acquire" or (a.start == "stream" and "taps" not in _CPU_STATE)
    nsym = nseg * S + 2 * H + (a.ntaps - 1 + 1) // 2
    if need_acq:
        nsym = max(nsym, min(a.acq, a.nsym) + a.ntaps)
    E, syms = synth.synth_numpy(a.M, nsym, seed=seed, snr_db=28.0)
    L_seg = S * 2 + a.ntaps - 1
    L_ext = (S + 2 * H) * 2 + a.ntaps - 1
    Es = np.stack([E[:, (H + s * S) * 2: (H + s * S) * 2 + L_seg] for s in range(nseg)])   # (nseg, 2, L_seg)
    Ex = np.stack([E[:, s * S * 2: s * S * 2 + L_ext] for s in range(nseg)]) if H else Es   # + halo
    kind = "fast_native"
    lib = co.lib(kind)
    try:
        lib.qo_set_threads(int(os.cpu_count() or 1))
    except Exception:
        pass
    threads = lib.qo_max_threads()
    alphabet = theory.normalised_symbols(a.M).astype(np.complex64)
    ang = theory.bps_test_angles(a.angles, np.float32)
    s1 = theory.reshape_symbols(None, "mcma", a.M, np.complex64, 2)
    s2 = theory.reshape_symbols(None, "mrde", a.M, np.complex64, 2)
    dt_acq = 0.0
    w0 = theory.init_taps(a.ntaps, 2, np.complex64)
    if need_acq:
        # dual_mode_equalisation(E, ..., TrSyms=(A, A), apply=False) on the head of the capture: one stream per mode
        A = min(a.acq, (E.shape[1] - a.ntaps + 1) // 2)
        Ea = np.ascontiguousarray(E[None, :, :(A - 1) * 2 + a.ntaps])
        wa = w0[None].copy()
        t0 = time.perf_counter()
        co.train_segments(Ea, A, 1, 2, 1e-3, wa, [0, 1], False, s1, "mcma", mu_shared=False, kind=kind)
        co.train_segments(Ea, A, 1, 2, 1e-3, wa, [0, 1], False, s2, "mrde", mu_shared=False, kind=kind)
        dt_acq = time.perf_counter() - t0
        _CPU_STATE["taps"] = wa[0]
    if a.start != "cold":
        w0 = _CPU_STATE["taps"]
    w = np.tile(w0, (nseg, 1, 1, 1))
    tr = theory.cal_training_symbol_len(2, a.ntaps, L_seg)
    t0 = time.perf_counter()
    co.train_segments(Es, tr, 1, 2, 1e-3, w, [0, 1], False, s1, "mcma", mu_shared=False, kind=kind)
    co.train_segments(Es, tr, 1, 2, 1e-3, w, [0, 1], False, s2, "mrde", mu_shared=False, kind=kind)
    eq = co.apply_segments(Ex, 2, w, None, kind=kind)                           # (nseg, 2, S + 2 H)
    idx = co.bps_streams(eq.reshape(nseg * 2, -1), ang, alphabet, a.bpsN, kind=kind)
    ph = ang[0][idx]
    ph[:, a.bpsN:-a.bpsN] = np.unwrap(ph[:, a.bpsN:-a.bpsN] * 4) / 4
    out = eq.reshape(nseg * 2, -1) * np.exp(1j * ph)
    dt = time.perf_counter() - t0
    assert np.isfinite(out).all()
    if a.start == "stream":
        _CPU_STATE["taps"] = w[-1].copy()
    own = np.ascontiguousarray(out.reshape(nseg, 2, S + 2 * H)[:, :, H:H + S]).astype(np.complex64)
    errs, cmpd = synth.ser_segments(torch.from_numpy(own), torch.from_numpy(syms), a.M, H + np.arange(nseg) * S,
                                    seg_chunk=16)
    return dt, dt_acq, nseg * S * 2, threads, int(errs.sum()), int(cmpd.sum())


def cpu_value(a, dt, dt_acq, samples, nseg):
    """Msamples/s of the CPU arm.  "acquire": one acquisition per capture of a.nsym symbols, so the sample's time is
    scaled to the whole capture before the acquisition is added."""
    if a.start != "acquire":
        return samples / dt / 1e6
    S = a.seg or a.nsym
    nseg_total = max(1, a.nsym // S)
    return 2.0 * nseg_total * S / (dt_acq + dt * nseg_total / nseg) / 1e6


def cpu_baseline(a, target_s):
    S = a.seg or a.nsym
    max_seg = max(1, a.nsym // S)
    n0 = min(max_seg, 16)
    dt, dt_acq, samples, threads, errs, cmpd = cpu_chain(a, n0)
    nseg = int(min(max_seg, max(n0, n0 * target_s / max(dt, 1e-3))))
    if nseg > n0:
        dt, dt_acq, samples, threads, errs, cmpd = cpu_chain(a, nseg)
    else:
        nseg = n0
    return {"value": cpu_value(a, dt, dt_acq, samples, nseg), "unit": "Msamples/s", "cores": threads, "kind": "port",
            "ser": errs / max(cmpd, 1), "start": a.start,
            "sample": "%d of %d segments of %d symbols (%.1f s of CPU work%s; oracle C port built with the "
                      "reference's flags -O3 -ffast-math -march=native -fopenmp; Pythran itself is not "
                      "installable here)" % (nseg, max_seg, S, dt, ("; + %.2f s acquisition of %d symbols per capture, "
                                                                      "segment time scaled to the capture" % (dt_acq, a.acq))
                                               if a.start == "acquire" else "")}, dt


def run_reference(a, rank, world):
    if rank != 0:
        return
    S = a.seg or a.nsym
    # calibrate the per-step sample so that warmup+steps stay within a few minutes
    base, t0 = cpu_baseline(a, min(a.cpu_seconds, 6.0))
    nseg = int(base["sample"].split(" of ")[0])
    times, vals, errs, cmpd = [], [], 0, 0
    for i in range(a.warmup + a.steps):
        dt, dt_acq, samples, threads, e, c = cpu_chain(a, nseg, seed=100 + i)
        if i >= a.warmup:
            times.append(dt + dt_acq)
            vals.append(cpu_value(a, dt, dt_acq, samples, nseg))
            errs += e
            cmpd += c
    t = sum(times) / len(times)
    val = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world),
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port",
                             "sample": "each step: %d segments of %d symbols of the workload" % (nseg, S)},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "sanity": {"ser": errs / max(cmpd, 1), "symbol_errors": errs, "symbols_compared": cmpd}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measured_traffic(kernel_key, a):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full
    capture of THIS command line (profiles/r02_traffic.json, written by scratch/ncu_traffic.py from the .ncu-rep);
    None when the capture is of another workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            t = json.load(fh)
        w = t.get("workload", {})
        if (w.get("nsym"), w.get("seg"), w.get("start"), w.get("err")) != (a.nsym, a.seg, a.start, not a.no_err):
            return None
        return t["kernels"].get(kernel_key)
    except Exception:
        return None


def host_link_probe(dev, world, nbytes=256 << 20, reps=3):
    """What the box's host side can move while every rank copies in BOTH directions at once (pinned memory, one copy
    engine each way): the ceiling of any end-to-end number.  Returns aggregate GB/s over all ranks (H2D + D2H bytes
    / slowest rank's time)."""
    import torch
    import torch.distributed as dist
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for r in range(reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_in.copy_(src, non_blocking=True)
        with torch.cuda.stream(s2):
            dst.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if r > 0:
            best = dt if best is None else min(best, dt)
    return 2.0 * nbytes * world / best / 1e9


def c5_record(a, rank, world, dev):
    """BASELINE config 5 inside the default line: ONE capture of 1e9 samples (5e8 symbols per polarisation) cut into
    contiguous per-rank ranges of whole segments -- strong scaling, no collective on the data path.  Same chain,
    same recipe (--start) and same gate as the headline; timed with CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from qampy_b200 import pipeline, synth
    nsym = 5 * 10 ** 8 // world
    cfg0 = pipeline.ReceiverConfig(M=a.M, ntaps=a.ntaps, os=2)
    seg = pipeline.balanced_segment_symbols(2 * nsym, cfg0, target=8192)
    cfg = pipeline.ReceiverConfig(M=a.M, ntaps=a.ntaps, os=2, mu=(1e-3, 1e-3), methods=("mcma", "mrde"),
                                  bps_angles=a.angles, bps_N=a.bpsN, seg_symbols=seg, want_err=not a.no_err,
                                  acq_symbols=a.acq, acq_layout=a.acq_layout, bps_halo=a.bpsN)
    rx = pipeline.SegmentedReceiver(cfg, dev)
    rx.want_idx = False
    E, syms0 = synth.synth_capture(a.M, nsym, seed=7000 + 100 * rank, snr_db=28.0, device=dev)
    taps = rx.acquire(E) if a.start != "cold" else None

    def step(taps):
        if a.start == "acquire":
            taps = rx.acquire(E)
        res = rx.run(E, wxy0=taps)
        return res, (rx.carry_taps(res) if a.start == "stream" else taps)

    res, taps = step(taps)
    del res
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    nstep = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = None
    for _ in range(nstep):
        res = None               # a step's results (36 GB at N = 1) go back to the allocator before the next step asks
        res, taps = step(taps)   # for its own: no cudaMalloc inside the timed region (measured 188 -> 283 ms with it)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nstep
    g = res[0]
    firsts = g["first"] + torch.arange(g["nseg"], device=dev) * g["nsym"]
    keep = firsts + g["nsym"] <= syms0.shape[1] - 64          # block 0 of the block-wise capture
    e, c = synth.ser_segments(g["out"][keep], syms0, a.M, firsts[keep])
    stats = torch.tensor([float(e.sum()), float(c.sum())], device=dev, dtype=torch.float64)
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    ser = float(stats[0] / stats[1].clamp(min=1))
    L = E.shape[1]
    del res, E
    torch.cuda.empty_cache()
    ok = ser < 1e-5
    return {"workload": "C5: one capture of 1e9 samples (5e8 symbols per polarisation) over %d GPU(s), %d symbols per "
                        "GPU in segments of %d; same chain, recipe and gate as the headline" % (world, nsym, seg),
            "scaling": "strong", "value": (world * L / (ms * 1e-3) / 1e6) if ok else None, "unit": "Msamples/s",
            "ms_per_step": ms, "steps": nstep, "n_gpus": world, "ser": ser,
            "symbols_compared": int(stats[1].item())}


def dropin_records(dev):
    """The equaliser calls of the reference's Scripts/*_equalisation.py (their dtype, length, taps, methods, step-size
    rule) as ONE capture through the drop-in API (qampy_b200.equalisation.dual_mode_equalisation: host array in, host
    arrays out) against the oracle port with the reference's flags.  The reference's own call shape is one serial
    stream per mode: this is where the GPU path is NOT faster; reported so that the headline
    is not mistaken for it."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_oracle as co
    from qampy_b200 import equalisation as eq, synth
    shapes = (("64_qam_equalisation.py", 64, 2 ** 17, 13, (0.19e-2, 0.19e-2), ("mcma", "mddma"), (True, True)),
              ("mrde_equaliser.py", 16, 2 ** 18, 30, (1e-3, 0.5e-3), ("mcma", "mrde"), (False, False)),
              ("32_qam_equalisation.py", 32, 10 ** 6, 11, (1e-3, 1e-3), ("mcma", "sbd"), (False, False)))
    out = []
    for name, M, nsym, ntaps, mu, methods, adaptive in shapes:
        E64, _ = synth.synth_signal(M, nsym, seed=3, snr_db=25.0, beta=0.01, theta=np.pi / 3, dgd=30e-12, device=dev)
        for dt in (np.complex128, np.complex64):
            E = E64.cpu().numpy().astype(dt)
            tg = []
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rg = eq.dual_mode_equalisation(E, 2, mu, M, Ntaps=ntaps, methods=methods, adaptive_stepsize=adaptive)
                torch.cuda.synchronize()
                tg.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            rc = co.dual_mode_equalisation(E, 2, mu, M, Ntaps=ntaps, methods=methods, adaptive_stepsize=adaptive,
                                           kind="fast_native")
            tc = time.perf_counter() - t0
            out.append({"script": name, "dtype": np.dtype(dt).name, "symbols": nsym, "ntaps": ntaps,
                        "methods": list(methods), "adaptive": list(adaptive), "gpu_s": min(tg), "cpu_s": tc,
                        "speedup": tc / min(tg),
                        "rms_diff": float(np.sqrt(np.mean(np.abs(rg[0] - rc[0]) ** 2)))})
    # BASELINE configs C2 and C3 in the reference's literal call shape: ONE capture, one stream per mode, equaliser + bps
    from qampy_b200 import phaserecovery as ph, theory
    for name, M, nsym, ntaps, mu, methods, A, N in (("C2 one capture", 16, 10 ** 6, 21, (1e-3,), ("mcma",), 32, 21),
                                                    ("C3 one capture", 64, 10 ** 7, 45, (1e-3, 1e-3), ("mcma", "mrde"), 64, 45)):
        E = synth.synth_signal(M, nsym, seed=5, snr_db=28.0, device=dev)[0].cpu().numpy()
        al = theory.normalised_symbols(M).astype(np.complex64)

        def gpu():
            if len(methods) == 1:
                Eo = eq.equalise_signal(E, 2, mu[0], M, Ntaps=ntaps, method=methods[0], apply=True)[0]
            else:
                Eo = eq.dual_mode_equalisation(E, 2, mu, M, Ntaps=ntaps, methods=methods)[0]
            return ph.bps(Eo, A, al, N)[0]

        def cpu():
            if len(methods) == 1:
                Eo = co.equalise_signal(E, 2, mu[0], M, Ntaps=ntaps, method=methods[0], apply=True, kind="fast_native")[0]
            else:
                Eo = co.dual_mode_equalisation(E, 2, mu, M, Ntaps=ntaps, methods=methods, kind="fast_native")[0]
            return co.bps_driver(Eo, A, al, N, kind="fast_native")[0]
        tg = []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rg = gpu()
            torch.cuda.synchronize()
            tg.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        rc = cpu()
        tc = time.perf_counter() - t0
        out.append({"script": name + " (equaliser + bps, host arrays in and out)", "dtype": "complex64", "symbols": nsym,
                    "ntaps": ntaps, "methods": list(methods), "bps": [A, N], "gpu_s": min(tg), "cpu_s": tc,
                    "speedup": tc / min(tg), "rms_diff": float(np.sqrt(np.mean(np.abs(rg - rc) ** 2)))})
    # phaserec.bps alone on one capture (few long streams: the phase-parallel form, csrc/bps_par.cu)
    for dt, nsym in ((np.complex64, 10 ** 7), (np.complex128, 10 ** 6)):
        M, A, N = 64, 64, 45
        al = theory.normalised_symbols(M).astype(dt)
        rng = np.random.default_rng(7)
        E = (al[rng.integers(0, M, (2, nsym))] * np.exp(0.1j)
             + 0.03 * (rng.standard_normal((2, nsym)) + 1j * rng.standard_normal((2, nsym)))).astype(dt)
        tg = []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rg = ph.bps(E, A, al, N)
            torch.cuda.synchronize()
            tg.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        rc = co.bps_driver(E, A, al, N, kind="fast_native")
        tc = time.perf_counter() - t0
        out.append({"script": "phaserec.bps on one capture (host arrays in and out)", "dtype": np.dtype(dt).name,
                    "symbols": nsym, "bps": [A, N], "gpu_s": min(tg), "cpu_s": tc, "speedup": tc / min(tg),
                    "phase_max_diff": float(np.max(np.abs(rg[1] - rc[1])))})
    return out


def run_b200(a, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from qampy_b200 import _lib, pipeline, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.require_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = pipeline.ReceiverConfig(M=a.M, ntaps=a.ntaps, os=2, mu=(1e-3, 1e-3), methods=("mcma", "mrde"),
                                  bps_angles=a.angles, bps_N=a.bpsN, seg_symbols=a.seg or None,
                                  want_err=not a.no_err, acq_symbols=a.acq, acq_layout=a.acq_layout,
                                  bps_halo=a.bpsN if a.seg else 0)
    rx = pipeline.SegmentedReceiver(cfg, dev)
    rx.want_idx = False
    # captures of ONE link (same channel, different symbols and noise): the steps alternate between them
    caps = []
    if a.nsym > 2 * 10 ** 7:    # long captures are synthesised block by block (double-precision FFT temporaries)
        caps.append(synth.synth_capture(a.M, a.nsym, seed=1000 + 100 * rank, snr_db=28.0, device=dev))
    else:
        for k in range(2 if a.start == "stream" else 1):
            caps.append(synth.synth_signal(a.M, a.nsym, seed=1000 + rank + 5000 * k, snr_db=28.0, device=dev))
    L = caps[0][0].shape[1]
    groups = pipeline.plan_segments(L, cfg)
    nsym_out = sum(n * k for _, n, k, _ in groups)
    torch.cuda.synchronize()

    # acquisition (timed on its own; inside every step for --start acquire)
    taps = None
    acq_ms = None
    if a.start != "cold":
        rx.acquire(caps[0][0])                               # untimed first call (module load, attribute set-up)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        taps = rx.acquire(caps[0][0])
        e1.record()
        torch.cuda.synchronize()
        acq_ms = e0.elapsed_time(e1)

    def step(i, taps):
        E = caps[i % len(caps)][0]
        if a.start == "acquire":
            taps = rx.acquire(E)
        res = rx.run(E, wxy0=taps)
        if a.start == "stream":
            taps = rx.carry_taps(res)
        return res, taps

    for i in range(a.warmup):
        res, taps = step(i, taps)
    del res

    # The timed steps are CUDA-graph replays (one graph per capture: both trainings, FIR, BPS of the main and the
    # end-aligned group, the hand-over of the carried taps) unless --no-graph: one launch per step instead of ~40
    # host-side enqueues, so a descheduled Python thread cannot leave the GPU idle inside a 90 ms timed region.  The
    # per-stage events are external event-record nodes of the graph: they hold the stage times of the last replay.
    import gc
    graphs = None
    launches_per_step = None
    if not a.no_graph:
        try:
            taps_static = taps.clone() if taps is not None else None
            graphs = []
            torch.cuda.synchronize()
            for k in range(len(caps)):
                g = torch.cuda.CUDAGraph()
                rx.events = []
                l0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    E = caps[k][0]
                    wx = rx.acquire(E) if a.start == "acquire" else taps_static
                    res_k = rx.run(E, wxy0=wx)
                    if a.start == "stream":
                        taps_static.copy_(rx.carry_taps(res_k))
                launches_per_step = _lib.launch_count() - l0
                graphs.append((g, res_k, rx.events))
                rx.events = None
            for k in range(len(graphs)):          # untimed replays (graph upload)
                graphs[k][0].replay()
            torch.cuda.synchronize()
        except Exception as exc:
            sys.stderr.write("bench: CUDA graph capture failed (%r); timing eager launches\n" % (exc,))
            graphs = None
            rx.events = None
            torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    if graphs is None:
        rx.events = []
        rx.event_pool = [torch.cuda.Event(enable_timing=True) for _ in range(2 * 8 * (a.steps + 1))]
    sampler = ClockSampler(local_rank)
    # (eager launches are enqueued by this Python thread: a garbage-collection pause in the middle of the timed
    # region would be measured as GPU time)
    gc.collect()
    gc.disable()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graphs is None:
        for i in range(a.steps):
            res, taps = step(a.warmup + i, taps)
    else:
        for i in range(a.steps):
            graphs[(a.warmup + i) % len(graphs)][0].replay()
    e1.record()
    barrier()
    gc.enable()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if graphs is None:
        launches = _lib.launch_count() - launches0
        events, rx.events = rx.events, None
        rx.event_pool = None
        nsamples = a.steps
    else:
        last = (a.warmup + a.steps - 1) % len(graphs)
        res, events = graphs[last][1], graphs[last][2]
        launches = launches_per_step * a.steps
        nsamples = 1                               # the external events hold the last replay of that graph
        if a.start == "stream":
            taps = taps_static
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / a.steps
    value = world * L / (ms_step * 1e-3) / 1e6

    # sanity gate on the output of the LAST timed step: symbol errors of every segment after the phase search
    # (reference bar: ser < 1e-5, test/test_equalisation.py:92-98), equaliser output power
    last_syms = caps[(a.warmup + a.steps - 1) % len(caps)][1]
    errs, cmpd, worst, nbad = 0, 0, 0.0, 0
    for g in res:
        firsts = g["first"] + torch.arange(g["nseg"], device=dev) * g["nsym"]
        keep = firsts + g["nsym"] <= last_syms.shape[1] - 64      # block-wise captures keep block 0's symbols only
        if int(keep.sum()) == 0:
            continue
        e, c = synth.ser_segments(g["out"][keep], last_syms, a.M, firsts[keep])
        errs += int(e.sum())
        cmpd += int(c.sum())
        per = e.sum(1).double() / c.sum(1).clamp(min=1).double()
        worst = max(worst, float(per.max()))
        nbad += int((e.sum(1) > 0).sum())
    ser_all = errs / max(cmpd, 1)
    out_rms = float(res[0]["eq"].abs().square().mean().sqrt())
    if world > 1:
        t = torch.tensor([float(errs), float(cmpd)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tw = torch.tensor([worst], device=dev, dtype=torch.float64)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        ser_all, worst = float(t[0] / t[1].clamp(min=1)), float(tw.item())
    sanity = {"ser": ser_all, "symbol_errors": errs, "symbols_compared": cmpd, "ser_max_segment": worst,
              "segments_with_errors": nbad, "segments": sum(g["nseg"] for g in res), "eq_out_rms": out_rms,
              "gate": "ser (all segments of the last timed step, after BPS, rank 0 counts; ser over ranks) < 1e-5"}
    del res

    # per-kernel device time inside the timed region (events on the launching stream)
    per = {}
    for name, (s, e) in events:
        per.setdefault(name, []).append(s.elapsed_time(e))
    # roofline of the dominant kernel = the stage with the largest device time per step.  Algorithmic
    # bytes per symbol period: train 48 B per pass (32 B without err), apply 48 B, bps 40 B.
    stage_ms = {k: sum(v) / nsamples for k, v in per.items()}
    bytes_train = BYTES_TRAIN if cfg.want_err else BYTES_TRAIN - 16
    stage_bytes = {"train": bytes_train * nsym_out * len(cfg.methods), "apply": BYTES_APPLY * nsym_out,
                   "bps": BYTES_BPS * nsym_out,
                   "acquire": 32 * min(a.acq, nsym_out) * len(cfg.methods)}
    stage_gbs = {k: stage_bytes[k] / (stage_ms[k] * 1e-3) / 1e9 for k in stage_ms if stage_ms[k] > 0}
    kernels = {"train": "train_la_kernel<8,12,METHOD,2> (look-ahead eq_train; two launches per step: mcma, mrde)",
               "apply": "apply_2x2_os2_kernel", "bps": "bps_fast_kernel<2>",
               "acquire": "single-stream trainer (%s layout), one stream per mode; two launches: mcma, mrde" % a.acq_layout}
    limits = {"train": "instruction dispatch: two training warps per SM sub-partition (4 serial streams each) saturate it, "
                       "packed fp32x2 FMAs take two dispatch cycles (60 flop/B at ntaps 45), not HBM",
              "apply": "FP32 FMA issue (30 flop/B at ntaps 45), not HBM",
              "bps": "instruction issue of the 64-angle distance search (19 instructions per symbol and angle), not HBM",
              "acquire": "serial depth of one stream (the recurrence of ONE mode is not parallel), not HBM"}
    nlaunch = {"train": len(cfg.methods), "apply": 1, "bps": 1, "acquire": len(cfg.methods)}
    # algorithmic FP32 FMAs per symbol period: train 2 modes * 2*45 taps * (4 dot + 4 update),
    # apply 2 * 90 * 4; the BPS search is not FMA work (compares, table look-ups), so it has no entry
    stage_fma = {"train": 8.0 * 2 * a.ntaps * 2 * nsym_out * len(cfg.methods), "apply": 4.0 * 2 * a.ntaps * 2 * nsym_out}
    dom = max(stage_ms, key=stage_ms.get) if stage_ms else "train"
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = stage_gbs.get(dom, 0.0)
    roofline = {"bound": "hbm", "kernel": kernels[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(dom, a),
                "launches_per_step": nlaunch[dom],
                "algorithmic_bytes_per_launch": stage_bytes[dom] / nlaunch[dom],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "binding_limit": limits[dom], "stage_ms_per_step": stage_ms, "stage_gbs": stage_gbs,
                "stage_frac": {k: v / peak for k, v in stage_gbs.items()},
                # secondary roof, the one that actually binds these kernels: FP32 FMA issue, 148 SMs x 128 lanes
                "fp32": {"peak_tfma_per_s": 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12,
                         "stage_tfma_per_s": {k: stage_fma[k] / (stage_ms[k] * 1e-3) / 1e12
                                              for k in stage_fma if stage_ms.get(k, 0) > 0},
                         "note": "fraction of the FP32 pipe = stage_tfma_per_s / peak_tfma_per_s"}}

    # end to end: pinned host capture -> H2D -> chain -> D2H of the recovered symbols + phase + the training errors
    # of both stages (what the reference call returns), with the copies of neighbouring chunks overlapping the
    # compute (pipeline.run_host)
    e2e = None
    affinity_before, near_cpus = (None, None) if (a.no_e2e or a.no_pin) else pin_to_gpu_cpus(sampler)
    if not a.no_e2e:
        E = caps[0][0]
        Eh = torch.empty(E.shape, dtype=E.dtype, pin_memory=True)
        Eh.copy_(E)
        Ed = torch.empty_like(E)
        outs = (None, None)
        times = []
        wx = taps
        gc.collect()
        gc.disable()             # ~200 host-side enqueues per step: a collection pause in between is measured as step time
        for it in range(2 + a.steps):
            barrier()
            t0 = time.perf_counter()
            if a.start == "acquire":
                # the head of the capture goes first; the acquisition runs while the rest is still on its way
                nacq = (min(a.acq, nsym_out) - 1) * 2 + a.ntaps
                Ed[:, :nacq].copy_(Eh[:, :nacq], non_blocking=True)
                wx = rx.acquire(Ed)
            outs = pipeline.run_host(rx, Eh, outs[0], outs[1], nchunks=a.chunks, E_dev=Ed, taper=not a.no_taper,
                                     wxy0=wx, err_host=getattr(rx, "err_host", None))[:2]
            if a.start == "stream":
                wx = rx.host_carry
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= 2:
                times.append(dt)
        gc.enable()
        t = sum(times) / len(times)
        t_sorted = sorted(times)
        if world > 1:
            tt = torch.tensor([t], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        h2d = E.numel() * E.element_size()
        d2h = sum(o.numel() * o.element_size() for o in outs)
        if cfg.want_err:
            d2h += sum(o.numel() * o.element_size() for o in rx.err_host)
        # the overlapped path must give the same symbols as the device-resident one (same capture, same start taps)
        wchk = wx_prev = None
        if a.start == "stream":
            # redo the last e2e step device-resident: its start taps were the carry of the step before
            wchk = rx.host_carry
            o1 = pipeline.run_host(rx, Eh, None, None, nchunks=a.chunks, E_dev=Ed, taper=not a.no_taper, wxy0=wchk)[0]
            torch.cuda.synchronize()
            chk = rx.run(E, wxy0=wchk)
            same = bool(torch.equal(o1[:chk[0]["nseg"]].to(dev), chk[0]["ext"]["out"]))
        else:
            chk = rx.run(E, wxy0=wx)
            same = bool(torch.equal(outs[0][:chk[0]["nseg"]].to(dev), chk[0]["ext"]["out"]))
        e2e = {"value": world * L / t / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": t * 1e3,
               "ms_per_step_spread": {"min": t_sorted[0] * 1e3, "median": t_sorted[len(t_sorted) // 2] * 1e3,
                                      "max": t_sorted[-1] * 1e3,
                                      "note": "rank 0, wall clock per step; value and ms_per_step are the MEAN over the "
                                              "steps (max over ranks): the host link is shared with whatever else "
                                              "runs on the box"},
               "chunks": a.chunks,
               "matches_device_path": same, "returns": "recovered symbols + phase" + (" + err1 + err2" if cfg.want_err else ""),
               "api": "pinned host capture -> qampy_b200.pipeline.run_host (H2D / chain / D2H overlapped per "
                      "chunk of segments) -> pinned host symbols + phase + training errors"}

    link_gbs = None
    if not a.no_e2e:
        link_gbs = host_link_probe(dev, world)
        if e2e is not None:
            e2e["host_link"] = {"aggregate_gbs_both_directions": link_gbs,
                                "floor_ms_per_step": (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) * world
                                / (link_gbs * 1e9) * 1e3,
                                "note": "all ranks copying pinned memory H2D and D2H at once (256 MB each way, slowest "
                                        "rank): what the host side of this box moves, the ceiling of any e2e number"}
            if near_cpus:
                e2e["cpu_affinity"] = "process bound to the %d CPUs NVML lists next to its GPU (%d-%d) while the pinned " \
                                      "buffers are allocated and used" % (len(near_cpus), near_cpus[0], near_cpus[-1])
    if affinity_before is not None:
        os.sched_setaffinity(0, affinity_before)      # the CPU baseline below uses every core it is allowed
    c5 = None
    if a.workload == "c3" and not a.no_c5:
        caps.clear()
        torch.cuda.empty_cache()
        try:
            c5 = c5_record(a, rank, world, dev)
        except Exception as exc:
            c5 = {"error": repr(exc)}
    if rank == 0:
        ok = ser_all < 1e-5
        line = {"metric": METRIC, "value": value if ok else None, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak" if a.workload == "c3" else "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
                "clocks": clocks, "e2e": e2e if ok else None, "gpu_launches": int(launches), "roofline": roofline,
                "msymbols_per_s": value / 2, "sanity": sanity,
                "timed_steps": "CUDA-graph replays of the step (one graph per capture; stage times = last replay)"
                               if graphs is not None else "eager launches"}
        if not ok:
            line["rejected"] = "symbol error rate %.2e of the recovered symbols is not below 1e-5: the number is withheld" % ser_all
            line["withheld"] = {"value": value, "e2e": e2e}
        if acq_ms is not None:
            cap_ms = ms_step + (acq_ms if a.start == "stream" else 0.0)
            line["acquisition"] = {
                "symbols_per_stage": min(a.acq, nsym_out), "ms": acq_ms, "layout": a.acq_layout,
                "cycles_per_symbol": acq_ms * 1e-3 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / (2 * min(a.acq, nsym_out)),
                "in_timed_region": a.start == "acquire",
                "value_if_every_capture_acquires": world * L / (cap_ms * 1e-3) / 1e6,
                "note": "dual_mode_equalisation(TrSyms=(A, A), apply=False) from centre-spike taps on the head of a "
                        "capture, one serial stream per mode; --start stream pays it once per link (before the "
                        "warm-up steps), --start acquire once per capture inside every timed step"}
        if c5 is not None:
            line["c5"] = c5
        if world == 1 and not a.no_dropin:
            try:
                line["dropin"] = dropin_records(dev)
            except Exception as exc:
                line["dropin"] = {"error": repr(exc)}
        if world == 1 and not a.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_baseline(a, a.cpu_seconds)
            except Exception as exc:   # the GPU number stands on its own
                line["cpu_baseline"] = {"value": None, "error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def resolve_segments(a):
    """--seg -1: balanced segment length for this capture (same value for the GPU and the CPU arm)."""
    if a.workload == "c5":
        # one capture of 5e8 symbols per polarisation; every rank holds the samples of its own contiguous range of
        # whole segments (pipeline.rank_capture_range: ntaps-1 samples of overlap with its neighbour, no exchange)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        a.nsym = 5 * 10 ** 8 // world
        a.no_e2e = True            # a 16 GB pinned host capture is not part of this mode
    if a.seg < 0:
        from qampy_b200 import pipeline
        cfg = pipeline.ReceiverConfig(M=a.M, ntaps=a.ntaps, os=2)
        n_sm = 148
        try:
            import torch
            if torch.cuda.is_available():     # the CPU arm has no device: it uses the B200's count, like the GPU arm
                n_sm = torch.cuda.get_device_properties(int(os.environ.get("LOCAL_RANK", "0"))).multi_processor_count
        except Exception:
            pass
        # two training warps per SM sub-partition: from there on the trainer is bound by instruction dispatch, not by the
        # issue cadence of a lone warp (profiles/README.md: 1 / 2 / 3 / 4 resident warps: 1.05 / 0.92 / 0.93 / 0.90 ms per pass)
        a.seg = pipeline.balanced_segment_symbols(2 * a.nsym, cfg, target=4096, n_sm=n_sm, warps_per_sm=8)
    return a


def main():
    # The one stdout line of this script is the JSON line.  Native libraries write to file descriptor 1 directly
    # (NCCL prints its version banner there when the first communicator comes up): hand fd 1 to stderr for the
    # whole run and keep the original stdout for Python's own print().
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    a = resolve_segments(parse())
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
    else:
        run_b200(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
