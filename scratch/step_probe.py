"""Does the NVML clock sampler perturb the timed region?  Step time with no sampler / 0.5 ms / 20 ms polling, and the
host time of one rx.run() call (no sync)."""
import sys, time, threading
sys.path.insert(0, ".")
import torch
import bench
from qampy_b200 import pipeline, synth

dev = torch.device("cuda", 0)
cfg = pipeline.ReceiverConfig(M=64, ntaps=45, os=2, mu=(1e-3, 1e-3), methods=("mcma", "mrde"), bps_angles=64, bps_N=45)
cfg.seg_symbols = pipeline.balanced_segment_symbols(2 * 10 ** 7, cfg, target=8192)
print("segment symbols", cfg.seg_symbols)
rx = pipeline.SegmentedReceiver(cfg, dev)
rx.want_idx = False
E, _ = synth.synth_signal(64, 10 ** 7, seed=1000, snr_db=28.0, device=dev)
for _ in range(5):
    res = rx.run(E)
torch.cuda.synchronize()


def timed(steps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        res = rx.run(E)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (t1 - t0) / steps * 1e3


for rep in range(3):
    print("no sampler      step %.3f ms  host enqueue %.3f ms" % timed())
    for period in (0.0005, 0.02):
        s = bench.ClockSampler(0)
        orig = time.sleep
        def poll(s=s, period=period):
            n = s.nvml
            while not s.stop_flag.is_set():
                t0 = time.perf_counter()
                sm = n.nvmlDeviceGetClockInfo(s.handle, n.NVML_CLOCK_SM)
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(s.handle)
                s.samples.append((float(sm), int(mask), time.perf_counter() - t0))
                time.sleep(period)
        s.thread = threading.Thread(target=poll, daemon=True)
        s.thread.start()
        r = timed()
        s.stop_flag.set(); s.thread.join()
        lat = sorted(x[2] for x in s.samples)
        print("sampler %6.1f ms step %.3f ms  host enqueue %.3f ms  samples %d  nvml call median %.2f ms max %.2f ms" %
              (period * 1e3, r[0], r[1], len(lat), lat[len(lat) // 2] * 1e3, lat[-1] * 1e3))
