"""BASELINE config C4 end to end (array level): dual-pol 256-QAM, 61 frames of 2**16 symbols, pilot sequence 2**10,
phase pilot every 32, ntaps 45: frame sync -> frequency offset -> pilot equaliser over 59 frames -> pilot CPE.
Times the CUDA path stage by stage; with `--cpu N` also the CPU oracle on N frames for comparison."""
import sys, time, types, warnings
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); warnings.filterwarnings("ignore")
import numpy as np, torch
from qampy_b200 import pilots, synth, theory
import qampy_b200.equalisation as be
dev = torch.device('cuda', 0)
M, fl, sl, rat, nfr = 256, 2 ** 16, 2 ** 10, 32, 61
d = synth.synth_pilot_signal(M, fl, sl, rat, nfr, snr_db=35, freq_off=100e6, linewidth=100e3, delay=4000, seed=11, device=dev)
rx, seq, php, idx_pil = d["E"].cpu().numpy(), d["pilot_seq"].cpu().numpy(), d["ph_pilots"].cpu().numpy(), d["idx_pil"]
idx = np.nonzero(idx_pil)[0][sl:]
frames = list(range(59))
def chain(backend, frames, batched=True):
    t = [time.perf_counter()]
    al, shiftf, foe, _, ok = pilots.sync2frame(rx, seq, 2, fl, backend=backend); t.append(time.perf_counter())
    rx3 = pilots.corr_foe(al, foe, 2); t.append(time.perf_counter())
    taps, eq, _ = pilots.pilot_equaliser_nframes(rx3, seq, shiftf, 2, fl, (1e-3, 1e-3), 45, synctaps=17, foe_comp=False,
                                                 frames=frames, methods=("cma", "sbd"), backend=backend, batched=batched)
    t.append(time.perf_counter())
    outs = [pilots.pilot_based_cpe_new(eq[:, k * fl:(k + 1) * fl], php, idx, fl, num_average=5, nframes=1)[0] for k in range(len(frames))]
    t.append(time.perf_counter())
    return np.diff(t), eq, outs
for rep in range(2):
    dt, eq, outs = chain(be, frames)
print('CUDA  frame_sync %.3f s  corr_foe %.3f s  pilot_eq(59 frames) %.3f s  cpe %.3f s  total %.3f s  -> %.1f Msymbols/s' %
      (*dt, dt.sum(), 2 * 59 * fl / dt.sum() / 1e6))
alphabet = theory.normalised_symbols(M).astype(np.complex64)
def ser_of(out, nfr_check=(0, 1, 57, 58)):
    sy = d["symbols"].cpu().numpy(); e = []
    for f in nfr_check:
        data = out[:, f * fl:(f + 1) * fl][:, ~idx_pil]; ref = sy[:, f * fl:(f + 1) * fl][:, ~idx_pil]
        dec = np.concatenate([alphabet[np.argmin(np.abs(data[:, a:a + 8192, None] - alphabet[None, None, :]), axis=-1)] for a in range(0, data.shape[1], 8192)], axis=1)
        e.append(float(np.mean(np.abs(dec - ref) > 1e-3)))
    return e
print('SER (array chain) frames 0,1,57,58:', ser_of(np.hstack(outs)))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = pilots.pilot_receiver(rx, seq, php[:, :idx.size], idx_pil, fl, 2, frames, to_host=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print('CUDA device-resident chain (host capture in, host result out): %.3f s -> %.1f Msymbols/s' % (t1 - t0, 2 * 59 * fl / (t1 - t0) / 1e6))
Ed = d["E"]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res2 = pilots.pilot_receiver(Ed, seq, php[:, :idx.size], idx_pil, fl, 2, frames, to_host=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print('CUDA device-resident chain (capture and result stay in HBM): %.3f s -> %.1f Msymbols/s' % (t1 - t0, 2 * 59 * fl / (t1 - t0) / 1e6))
print('SER (device chain) frames 0,1,57,58:', ser_of(res["out"]))
dt2, _, _ = chain(be, frames, batched=False)
print('CUDA frame-by-frame pilot_eq %.3f s' % dt2[2])
if '--cpu' in sys.argv:
    import cpu_oracle as co
    n = int(sys.argv[sys.argv.index('--cpu') + 1])
    ob = types.SimpleNamespace(equalise_signal=lambda *a, **k: co.equalise_signal(*a, kind="fast_native", **k),
                               apply_filter=lambda *a, **k: co.apply_filter(*a, kind="fast_native", **k))
    dtc, _, _ = chain(ob, list(range(n)), batched=False)
    print('CPU oracle (reference flags) frame_sync %.3f s corr_foe %.3f s pilot_eq(%d frames) %.3f s cpe %.3f s' % (dtc[0], dtc[1], n, dtc[2], dtc[3]))
