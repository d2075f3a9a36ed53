#!/usr/bin/env python
"""bench.py -- headline benchmark of the qampy_b200 hot path (BASELINE.json metric).

Workload (``config.workload``): BASELINE config C3 per GPU -- dual-pol 64-QAM, 2 samples/symbol,
``--nsym`` (default 1e7) symbols, MCMA -> MRDE training (ntaps 45, mu 1e-3) -> apply -> BPS(64 test
angles, N = 45), processed as independent time segments of ``--seg`` output symbols (ntaps-1 overlap,
every segment trained from centre-spike taps; segment s == the reference called on that segment).
One "step" = one pass of the whole chain over the capture.  N > 1: every rank owns its own capture
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"],
                    help="c3 (default, the contract's N=1 workload): --nsym symbols per GPU, weak scaling; c5: BASELINE "
                         "config 5, ONE capture of 5e8 symbols per polarisation (1e9 samples) cut into contiguous "
                         "per-rank ranges of whole segments (strong scaling, no collective)")
    ap.add_argument("--nsym", type=int, default=10 ** 7, help="symbols per polarisation per GPU")
    ap.add_argument("--seg", type=int, default=-1,
                    help="output symbols per segment; -1 = smallest length >= 8192 that fills whole GPU waves "
                         "(pipeline.balanced_segment_symbols), 0 = one segment")
    ap.add_argument("--chunks", type=int, default=6, help="host<->device overlap chunks of the e2e path (the last one "
                                                          "is cut into 1/2 + 1/4 + 1/4 unless --no-taper)")
    ap.add_argument("--ntaps", type=int, default=45)
    ap.add_argument("--M", type=int, default=64)
    ap.add_argument("--angles", type=int, default=64)
    ap.add_argument("--bpsN", type=int, default=45)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-taper", action="store_true", help="e2e path: do not cut the last chunk into 1/2 + 1/4 + 1/4")
    return ap.parse_args()


def workload_config(a, world):
    name = "C3" if a.workload == "c3" else "C5 (1e9-sample capture over %d GPU(s))" % world
    return {"workload": "%s dual-pol %d-QAM dual_mode_equalisation(mcma->mrde, ntaps=%d, mu=1e-3) + bps(%d, N=%d), "
                        "2 sps, %d symbols per GPU" % (name, a.M, a.ntaps, a.angles, a.bpsN, a.nsym),
            "symbols_per_gpu": a.nsym, "samples_per_gpu": 2 * a.nsym, "segment_symbols": a.seg or a.nsym,
            "segment_semantics": "each segment == reference call on that segment (centre-spike taps)",
            "sharding": ("dp%d (independent captures per rank, no collective)" % world) if a.workload == "c3" else
                        ("dp%d (every rank synthesises and processes its own contiguous range of %d symbols of the "
                         "5e8-symbol capture; whole segments, no collective)" % (world, a.nsym)),
            "l2_policy": "inputs (%d MB/GPU) larger than L2 (126 MB); no explicit flush" % (a.nsym * 32 // 10 ** 6)}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
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
def cpu_chain(a, nseg, seed=1234):
    """Times the reference-equivalent CPU chain on `nseg` segments of the workload.  Returns
    (seconds, samples processed, threads)."""
    # all host threads: torchrun exports OMP_NUM_THREADS=1 to its workers, which would cripple the CPU arm
    if "cpu_oracle" not in sys.modules:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_oracle as co
    from qampy_b200 import synth, theory
    S = a.seg or a.nsym
    nsym = nseg * S + (a.ntaps - 1 + 1) // 2
    E, _ = synth.synth_numpy(a.M, nsym, seed=seed, snr_db=28.0)
    L_seg = S * 2 + a.ntaps - 1
    Es = np.stack([E[:, s * S * 2: s * S * 2 + L_seg] for s in range(nseg)])   # (nseg, 2, L_seg)
    kind = "fast_native"
    lib = co.lib(kind)
    try:
        lib.qo_set_threads(int(os.cpu_count() or 1))
    except Exception:
        pass
    threads = lib.qo_max_threads()
    alphabet = theory.normalised_symbols(a.M).astype(np.complex64)
    ang = theory.bps_test_angles(a.angles, np.float32)
    w = np.tile(theory.init_taps(a.ntaps, 2, np.complex64), (nseg, 1, 1, 1))
    tr = theory.cal_training_symbol_len(2, a.ntaps, L_seg)
    s1 = theory.reshape_symbols(None, "mcma", a.M, np.complex64, 2)
    s2 = theory.reshape_symbols(None, "mrde", a.M, np.complex64, 2)
    t0 = time.perf_counter()
    co.train_segments(Es, tr, 1, 2, 1e-3, w, [0, 1], False, s1, "mcma", mu_shared=False, kind=kind)
    co.train_segments(Es, tr, 1, 2, 1e-3, w, [0, 1], False, s2, "mrde", mu_shared=False, kind=kind)
    eq = co.apply_segments(Es, 2, w, None, kind=kind)                           # (nseg, 2, S)
    idx = co.bps_streams(eq.reshape(nseg * 2, -1), ang, alphabet, a.bpsN, kind=kind)
    ph = ang[0][idx]
    ph[:, a.bpsN:-a.bpsN] = np.unwrap(ph[:, a.bpsN:-a.bpsN] * 4) / 4
    out = eq.reshape(nseg * 2, -1) * np.exp(1j * ph)
    dt = time.perf_counter() - t0
    assert np.isfinite(out).all()
    return dt, nseg * S * 2, threads


def cpu_baseline(a, target_s):
    S = a.seg or a.nsym
    max_seg = max(1, a.nsym // S)
    n0 = min(max_seg, 16)
    dt, samples, threads = cpu_chain(a, n0)
    nseg = int(min(max_seg, max(n0, n0 * target_s / max(dt, 1e-3))))
    if nseg > n0:
        dt, samples, threads = cpu_chain(a, nseg)
    else:
        nseg = n0
    return {"value": samples / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": "%d of %d segments of %d symbols (%.1f s of CPU work; oracle C port built with the "
                      "reference's flags -O3 -ffast-math -march=native -fopenmp; Pythran itself is not "
                      "installable here)" % (nseg, max_seg, S, dt)}, dt


def run_reference(a, rank, world):
    if rank != 0:
        return
    S = a.seg or a.nsym
    # calibrate the per-step sample so that warmup+steps stay within a few minutes
    base, t0 = cpu_baseline(a, min(a.cpu_seconds, 6.0))
    nseg = int(base["sample"].split(" of ")[0])
    times = []
    for i in range(a.warmup + a.steps):
        dt, samples, threads = cpu_chain(a, nseg, seed=100 + i)
        if i >= a.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    val = samples / t / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world),
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port",
                             "sample": "each step: %d segments of %d symbols of the workload" % (nseg, S)},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
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
                                  bps_angles=a.angles, bps_N=a.bpsN, seg_symbols=a.seg or None)
    rx = pipeline.SegmentedReceiver(cfg, dev)
    rx.want_idx = False
    if a.nsym > 2 * 10 ** 7:    # long captures are synthesised block by block (double-precision FFT temporaries)
        E, syms = synth.synth_capture(a.M, a.nsym, seed=1000 + 100 * rank, snr_db=28.0, device=dev)
    else:
        E, syms = synth.synth_signal(a.M, a.nsym, seed=1000 + rank, snr_db=28.0, device=dev)
    L = E.shape[1]
    groups = pipeline.plan_segments(L, cfg)
    nsym_out = sum(n * k for _, n, k, _ in groups)
    torch.cuda.synchronize()

    for _ in range(a.warmup):
        res = rx.run(E)
    # sanity gate (printed with the number): equaliser output power and SER of one segment
    eq0 = res[0]["eq"][0].cpu().numpy()
    S0 = res[0]["nsym"]
    out_rms = float(np.sqrt(np.mean(np.abs(eq0) ** 2)))
    ser = synth.ser(eq0[:, :min(S0, 20000)], syms[:, :min(S0, 20000) + 200].cpu().numpy(), a.M)
    del res

    launches0 = _lib.launch_count()
    rx.events = []
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        res = rx.run(E)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    events, rx.events = rx.events, None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / a.steps
    value = world * L / (ms_step * 1e-3) / 1e6

    # per-kernel device time inside the timed region (events on the launching stream)
    per = {}
    for name, (s, e) in events:
        per.setdefault(name, []).append(s.elapsed_time(e))
    # roofline of the dominant kernel = the stage with the largest device time per step.  Algorithmic
    # bytes per symbol period: train 48 B per pass, apply 48 B, bps 40 B.
    stage_ms = {k: sum(v) / a.steps for k, v in per.items()}
    stage_bytes = {"train": BYTES_TRAIN * nsym_out * len(cfg.methods), "apply": BYTES_APPLY * nsym_out,
                   "bps": BYTES_BPS * nsym_out}
    stage_gbs = {k: stage_bytes[k] / (stage_ms[k] * 1e-3) / 1e9 for k in stage_ms if stage_ms[k] > 0}
    kernels = {"train": "train_la_kernel<8,12,METHOD,2> (look-ahead eq_train; two launches per step: mcma, mrde)",
               "apply": "apply_2x2_os2_kernel", "bps": "bps_fast_kernel<2>"}
    limits = {"train": "instruction issue of one warp per SM sub-partition (4 serial streams each; 60 flop/B at "
                       "ntaps 45), not HBM",
              "apply": "FP32 FMA issue (30 flop/B at ntaps 45), not HBM",
              "bps": "instruction issue of the 64-angle distance search (19 instructions per symbol and angle), not HBM"}
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
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this
    # workload (profiles/r01_ncu_full_summary.txt); algorithmic bytes per launch are stage_bytes / launches
    traffic = {"train": TRAFFIC_TRAIN, "apply": 450.7e6, "bps": TRAFFIC_BPS}
    roofline = {"bound": "hbm", "kernel": kernels[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic.get(dom) if a.nsym == 10 ** 7 else None,
                "launches_per_step": {"train": len(cfg.methods), "apply": 1, "bps": 1}[dom],
                "algorithmic_bytes_per_launch": stage_bytes[dom] / {"train": len(cfg.methods), "apply": 1, "bps": 1}[dom],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "binding_limit": limits[dom], "stage_ms_per_step": stage_ms, "stage_gbs": stage_gbs,
                "stage_frac": {k: v / peak for k, v in stage_gbs.items()},
                # secondary roof, the one that actually binds these kernels: FP32 FMA issue, 148 SMs x 128 lanes
                "fp32": {"peak_tfma_per_s": 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12,
                         "stage_tfma_per_s": {k: stage_fma[k] / (stage_ms[k] * 1e-3) / 1e12
                                              for k in stage_fma if stage_ms.get(k, 0) > 0},
                         "note": "fraction of the FP32 pipe = stage_tfma_per_s / peak_tfma_per_s"}}

    # end to end: pinned host capture -> H2D -> chain -> D2H of the recovered symbols + phase, with
    # the copies of neighbouring chunks overlapping the compute (pipeline.run_host)
    e2e = None
    if not a.no_e2e:
        Eh = torch.empty(E.shape, dtype=E.dtype, pin_memory=True)
        Eh.copy_(E)
        Ed = torch.empty_like(E)
        outs = (None, None)
        times = []
        for it in range(2 + a.steps):
            barrier()
            t0 = time.perf_counter()
            outs = pipeline.run_host(rx, Eh, outs[0], outs[1], nchunks=a.chunks, E_dev=Ed, taper=not a.no_taper)[:2]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= 2:
                times.append(dt)
        t = sum(times) / len(times)
        if world > 1:
            tt = torch.tensor([t], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        h2d = E.numel() * E.element_size()
        d2h = sum(o.numel() * o.element_size() for o in outs)
        # the overlapped path must give the same symbols as the device-resident one
        chk = rx.run(E)
        same = bool(torch.equal(outs[0][:chk[0]["nseg"]].to(dev), chk[0]["out"]))
        e2e = {"value": world * L / t / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": t * 1e3, "chunks": a.chunks,
               "matches_device_path": same,
               "api": "pinned host capture -> qampy_b200.pipeline.run_host (H2D / chain / D2H overlapped per "
                      "chunk of segments) -> pinned host symbols + phase"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak" if a.workload == "c3" else "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "msymbols_per_s": value / 2, "sanity": {"eq_out_rms": out_rms, "ser_segment0": ser}}
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
        a.seg = pipeline.balanced_segment_symbols(2 * a.nsym, cfg, target=8192)
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
