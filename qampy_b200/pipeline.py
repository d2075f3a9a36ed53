"""Device-resident receiver chain over independent time segments (BASELINE configs C3/C5).

    capture (nmodes, L)  --segment_view-->  (nseg, nmodes, L_seg)      (no copy, ntaps-1 overlap)
        stage 1 training (e.g. MCMA)        one launch, nseg*nmodes serial streams in parallel
        stage 2 training (e.g. MRDE)        restarts at sample 0 of the segment with the stage-1 taps
        apply final taps + decimate         (nseg, nmodes, S) equalised symbols, stays in HBM
        blind phase search                  (nseg*nmodes) streams of S symbols, fused tail

Semantics (what the parity tests check): segment ``s`` gives exactly what the reference gives when
``dual_mode_equalisation`` (``qampy/core/equalisation/equalisation.py:400-466``) followed by ``bps``
(``qampy/core/phaserecovery.py:93-159``) is called on that segment's samples alone, with the taps
initialised as the caller says (centre spike, or a warm start such as the taps of a previous
capture -- the reference's own ``wxy=`` mechanism).  Segment ``s`` owns output symbols
``[s*S, (s+1)*S)`` and reads input samples ``[s*S*os, s*S*os + S*os + ntaps - 1)``; if S does not
divide the capture, one extra segment of the same length is aligned to the end of the capture.  A single segment (``seg_symbols=None``) is the reference call on the whole
capture.  Nothing here touches the host between stages and no collective is involved: ranks of a
multi-GPU job own disjoint ranges of segments (``shard_segments``).
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import device, theory


@dataclass
class ReceiverConfig:
    M: int = 64
    ntaps: int = 45
    os: int = 2
    mu: tuple = (1e-3, 1e-3)
    methods: tuple = ("mcma", "mrde")
    niter: tuple = (1, 1)
    bps_angles: int = 64
    bps_N: int = 45
    seg_symbols: int = None     # output symbols per segment; None = one segment (reference semantics)
    want_err: bool = False      # keep the per-symbol training error (reference returns it; 16 B/symbol/pass)
    bps_halo: int = 0           # symbols of equalised signal on either side of a segment that its phase search also
    #                             sees (0 = none).  The reference's bps leaves the first and last N symbols of a call
    #                             without an estimate (phaserecovery.py:150-155: idx 0 -> rotated by angles[0]), so
    #                             back-to-back segments need bps_halo >= bps_N for a seamless capture.
    acq_symbols: int = 1 << 18  # training symbols per stage of SegmentedReceiver.acquire (tap acquisition)
    acq_layout: str = "latency"  # training-kernel layout of the (single-stream) acquisition


def plan_segments(L, cfg):
    """Split the N = (L - ntaps + 1)//os output symbols of a capture into equal segments of
    cfg.seg_symbols.  Returns a list of groups (first_output_symbol, n_symbols, n_segments, drop):
    the main group holds the N//S back-to-back segments; if S does not divide N, a second group holds
    ONE more segment of the same length aligned to the END of the capture (it overlaps its
    predecessor; only its last N % S symbols -- everything after ``drop`` -- are kept when stitching).
    All segments therefore have the same length and cost, and the extra one runs concurrently.
    A single segment (seg_symbols None or not shorter than the capture) is the reference call on the whole capture
    and has no halo."""
    N = (L - cfg.ntaps + 1) // cfg.os
    if N <= 0:
        raise ValueError("capture shorter than the filter")
    S = cfg.seg_symbols
    H = int(cfg.bps_halo)
    if S is None or S >= N - 2 * H:
        return [(0, N, 1, 0)]
    # with a halo the segments tile [H, N - H): the H symbols at either end of the capture have no neighbour to
    # borrow from (the reference gives them no phase estimate either)
    nfull, rem = divmod(N - 2 * H, S)
    groups = [(H, S, nfull, 0)]
    if rem:
        groups.append((N - H - S, S, 1, S - rem))
    return groups


def balanced_segment_symbols(L, cfg, target=8192, nmodes=2, n_sm=148, streams_per_warp=4, warps_per_sm=4):
    """Segment length >= ``target`` for which the trained streams fill whole waves of the GPU.

    The training kernel runs ``streams_per_warp`` (segment, mode) streams per warp and is latency bound:
    a launch takes as long as its busiest SM sub-partition, so 4.1 warps per SM cost ~45 % more than 4.0
    (profiles/README.md).  This picks the smallest S such that all segments -- including the extra
    end-aligned one -- need at most k * n_sm * warps_per_sm warps for the smallest possible k."""
    N = (L - cfg.ntaps + 1) // cfg.os
    wave = n_sm * warps_per_sm                                        # warps per full wave

    def warps(S):
        nfull, rem = divmod(N, S)
        w = -(-nfull * nmodes // streams_per_warp)
        return w + (-(-nmodes // streams_per_warp) if rem and nfull else 0)   # + the end-aligned extra launch

    S = min(target, N)
    k0 = warps(S) // wave                   # whole waves at the target length
    if k0 == 0:
        return S                            # less than one wave: nothing to balance
    while S < N and warps(S) > k0 * wave:   # grow S (by at most a factor 1 + 1/k0) to drop the partial wave
        S += 1
    return S


def shard_segments(nseg, rank, world):
    """Contiguous block of segment indices owned by ``rank`` (no exchange between ranks)."""
    base, extra = divmod(nseg, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks of the default process group (identity without one).
    Used for the multi-GPU timing rule: a step takes as long as its slowest rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rank_capture_range(nsym_total, ntaps, os, seg_symbols, rank, world, halo=0):
    """Sample range [a, b) of a long capture that ``rank`` must hold to produce ITS contiguous block of
    segments (whole segments, ntaps-1 samples -- plus ``halo`` symbols on either side when the phase search borrows
    from the neighbours, ReceiverConfig.bps_halo -- of overlap with the next rank, no exchange), plus the
    range of output symbols it owns.  A rank that runs plan_segments on its range with the same halo produces exactly
    its own symbols.  SURVEY.md section 8e."""
    L = nsym_total * os
    N = (L - ntaps + 1) // os
    nseg = (N - 2 * halo) // seg_symbols
    lo, hi = shard_segments(nseg, rank, world)
    first_sym, last_sym = halo + lo * seg_symbols, halo + hi * seg_symbols
    if rank == world - 1:
        last_sym = N - halo               # the last rank also takes the remainder (end-aligned extra segment)
    a = (first_sym - halo) * os
    b = min(L, a + (last_sym - first_sym + 2 * halo) * os + ntaps - 1) if last_sym > first_sym else a
    return a, b, first_sym, last_sym


class SegmentedReceiver:
    def __init__(self, cfg, dev=None, cdtype=np.complex64, nmodes=2):
        self.cfg = cfg
        self.dev = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())
        self.cdtype = np.dtype(cdtype)
        self.tdtype = torch.complex64 if self.cdtype == np.complex64 else torch.complex128
        self.rdtype = torch.float32 if self.cdtype == np.complex64 else torch.float64
        self.nmodes = nmodes
        self.syms = []
        for m in cfg.methods:
            if m in theory.DATA_AIDED:
                raise NotImplementedError("data-aided methods need per-segment training sequences")
            t = theory.reshape_symbols(None, m, cfg.M, self.cdtype.type, nmodes)
            self.syms.append(torch.from_numpy(np.ascontiguousarray(t)).to(self.dev))
        alphabet = theory.normalised_symbols(cfg.M).astype(self.cdtype)
        self.bps_tables = device.BpsTables(cfg.bps_angles, alphabet, self.cdtype.type, self.dev)
        self.w0 = torch.from_numpy(theory.init_taps(cfg.ntaps, nmodes, self.cdtype.type)).to(self.dev)
        self.side = None
        self._streams = None
        self.events = None      # set to a list to collect (name, (start, end)) CUDA events per launch
        self.want_idx = True
        self.event_pool = None  # optional list of pre-created timing events for _tic
        self.trace = None       # set to a list: run_host appends (label, timing event) per chunk and copy (scratch/e2e_trace.py)
        self._nvtx_open = False

    def _tic(self, name):
        # NVTX range per stage (tracing): costs nothing without a profiler attached; the range
        # is closed by _toc / the next _tic
        if self._nvtx_open:
            torch.cuda.nvtx.range_pop()
        torch.cuda.nvtx.range_push("qampy_b200." + name)
        self._nvtx_open = True
        if self.events is not None:
            # events from a pool made before the timed region where there is one (bench.py): creating them costs a
            # driver call each
            pool = self.event_pool
            if torch.cuda.is_current_stream_capturing():
                # inside a CUDA graph: external event-record nodes, re-recorded by every replay
                ev = (torch.cuda.Event(enable_timing=True, external=True),
                      torch.cuda.Event(enable_timing=True, external=True))
            else:
                ev = (pool.pop(), pool.pop()) if pool is not None and len(pool) >= 2 else \
                    (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            self.events.append((name, ev))
            return ev[1]
        return None

    def _toc(self):
        if self._nvtx_open:
            torch.cuda.nvtx.range_pop()
            self._nvtx_open = False

    def _run_group(self, E, first, nsym, nseg, drop, wxy0, between, stage_done=None):
        cfg = self.cfg
        Ev = device.segment_view(E[:, first * cfg.os:], nseg, nsym, cfg.os, cfg.ntaps)
        L_seg = Ev.shape[2]
        w0 = self.w0 if wxy0 is None else wxy0          # (nmodes, nmodes, ntaps): same start for all segments
        assert w0.dim() == 3, "wxy0 must be (nmodes, nmodes, ntaps)"
        w = w0.unsqueeze(0).repeat(nseg, 1, 1, 1)        # fresh copy per segment (trained in place)
        trsyms = theory.cal_training_symbol_len(cfg.os, cfg.ntaps, L_seg)
        errs = []
        for stage in range(len(cfg.methods)):
            mu = torch.full((nseg, self.nmodes), float(cfg.mu[stage]), dtype=self.rdtype, device=self.dev)
            err = None
            if cfg.want_err:
                err = torch.empty((nseg, self.nmodes, trsyms * cfg.niter[stage]), dtype=self.tdtype,
                                  device=self.dev)
            t = self._tic("train")
            device.train_equaliser(Ev, trsyms, cfg.niter[stage], cfg.os, mu, w, None, False,
                                   self.syms[stage], cfg.methods[stage], err)
            t and t.record()
            errs.append(err)
            if stage_done is not None:
                stage_done(stage, err)      # e.g. run_host: the stage's error array can leave while the chain goes on
        # the final taps of a segment filter its own samples plus the halo on either side (H = 0: the same view)
        H = cfg.bps_halo if first >= cfg.bps_halo else 0
        next_ = nsym + 2 * H
        Ea = Ev if H == 0 else device.segment_view(E[:, (first - H) * cfg.os:], nseg, next_, cfg.os, cfg.ntaps,
                                                    step_symbols=nsym)
        t = self._tic("apply")
        eq = device.apply_filter_to_signal(Ea, cfg.os, w)              # (nseg, nmodes, nsym + 2 H)
        t and t.record()
        bin_ = eq if between is None else between(eq)
        t = self._tic("bps")
        out, ph, idx = device.bps(bin_.reshape(nseg * self.nmodes, next_), self.bps_tables, cfg.bps_N,
                                  want_idx=self.want_idx)
        t and t.record()
        if idx is None:
            idx = ph
        shp = (nseg, self.nmodes, next_)
        self._toc()
        ext = dict(eq=eq, out=out.reshape(shp), ph=ph.reshape(shp), idx=idx.reshape(shp))
        own = {k: v[:, :, H:H + nsym] for k, v in ext.items()}         # the segment's own symbols (views)
        return dict(own, ext=ext, halo=H, taps=w, err=errs, first=first, nsym=nsym, nseg=nseg, drop=drop)

    def acquire(self, E, nsym=None, wxy0=None):
        """Tap acquisition on the head of a capture: the reference call ``dual_mode_equalisation(E, os, mu, M,
        wxy=wxy0 or Ntaps=ntaps, TrSyms=(A, A), methods=cfg.methods, apply=False)``
        (``core/equalisation/equalisation.py:400-466``; ``_lms_init`` :391-397 cuts the field to the first
        ``(A-1)*os + ntaps`` samples) -- ONE stream per mode, as deep as A symbols per stage, so it runs in the
        single-stream layout of the trainer.  Short cold-started segments do not converge (64-QAM, mu 1e-3: SER
        3e-4 after 8454 symbols, 0 after 2^18; scratch/conv_study.py); the taps acquired here are what every
        segment of :meth:`run` starts from (``wxy0``), and what a streaming receiver carries from capture to
        capture.  Returns taps (nmodes, nmodes, ntaps)."""
        cfg = self.cfg
        assert E.is_cuda and E.dim() == 2 and E.shape[0] == self.nmodes and E.stride(1) == 1
        N = (E.shape[1] - cfg.ntaps + 1) // cfg.os
        A = min(int(nsym or cfg.acq_symbols), N)
        Ev = E[:, :(A - 1) * cfg.os + cfg.ntaps].unsqueeze(0)
        w = (self.w0 if wxy0 is None else wxy0).clone().unsqueeze(0).contiguous()
        for stage in range(len(cfg.methods)):
            mu = torch.full((1, self.nmodes), float(cfg.mu[stage]), dtype=self.rdtype, device=self.dev)
            t = self._tic("acquire")
            device.train_equaliser(Ev, A, cfg.niter[stage], cfg.os, mu, w, None, False, self.syms[stage],
                                   cfg.methods[stage], None, layout=cfg.acq_layout)
            t and t.record()
        self._toc()
        return w[0]

    @staticmethod
    def carry_taps(res):
        """Taps a streaming receiver hands to the next capture: those of the chronologically last full segment."""
        return res[0]["taps"][-1]

    def run(self, E, wxy0=None, between=None):
        """E: (nmodes, L) complex CUDA tensor.  Returns one result dict per segment group (see
        plan_segments); use :func:`stitch` for (nmodes, N) arrays.  The end-aligned extra segment (if
        any) is enqueued on a side stream so that it overlaps the main group."""
        assert E.is_cuda and E.dim() == 2 and E.shape[0] == self.nmodes and E.stride(1) == 1
        groups = plan_segments(E.shape[1], self.cfg)
        main = torch.cuda.current_stream()
        res = [None] * len(groups)
        if len(groups) > 1:
            if self.side is None:
                self.side = torch.cuda.Stream()
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                events, self.events = self.events, None      # per-launch events only on the main stream
                res[1] = self._run_group(E, *groups[1], wxy0, between)
                self.events = events
                if not torch.cuda.is_current_stream_capturing():     # (a captured graph owns its memory pool)
                    for v in list(res[1].values()) + list(res[1]["ext"].values()):
                        if torch.is_tensor(v):
                            v.record_stream(main)
        res[0] = self._run_group(E, *groups[0], wxy0, between)
        if len(groups) > 1:
            main.wait_stream(self.side)
        return res


def _host_chunks(groups, nchunks, taper=True):
    """Split the main group into ``nchunks`` runs of whole segments; the end-aligned extra segment (if
    any) becomes the last chunk.  Yields (first_symbol, nsym, nseg, drop, seg_index0).

    ``taper``: the last run is cut once more into 1/2 + 1/4 + 1/4 of its segments.  The wall time of the overlapped
    path ends with (last H2D) -> chain of the last run (as deep as any: a segment is a serial recurrence) -> its
    D2H, so the shorter the last run, the shorter the copy that nothing can hide."""
    first, nsym, nseg, drop = groups[0]
    nchunks = max(1, min(nchunks, nseg))
    bounds = [shard_segments(nseg, c, nchunks) for c in range(nchunks)]
    if taper and nchunks > 1 and bounds[0][1] - bounds[0][0] >= 8:
        # and the first run into 1/4 + 1/4 + 1/2: nothing can be downloaded before the first run's first training
        # stage is done, and the download is the longer direction of the link -- the sooner it starts the better
        lo, hi = bounds.pop(0)
        n = hi - lo
        cuts = [lo, lo + n // 4, lo + n // 2, hi]
        bounds = [(cuts[i], cuts[i + 1]) for i in range(3)] + bounds
    if taper and nchunks > 1 and bounds[-1][1] - bounds[-1][0] >= 8:
        lo, hi = bounds.pop()
        n = hi - lo
        cuts = [lo, lo + n // 2, lo + n // 2 + n // 4, hi]
        bounds += [(cuts[i], cuts[i + 1]) for i in range(3)]
    for lo, hi in bounds:
        if hi > lo:
            yield first + lo * nsym, nsym, hi - lo, 0, lo
    if len(groups) > 1:
        f2, n2, k2, d2 = groups[1]
        yield f2, n2, k2, d2, nseg


def run_host(rx, E_host, out_host=None, ph_host=None, nchunks=6, E_dev=None, taper=True, wxy0=None, err_host=None):
    """End-to-end form of :meth:`SegmentedReceiver.run` for a capture in (pinned) HOST memory.

    The capture is cut into ``nchunks`` runs of whole segments; the H2D copy of chunk c+1, the chain of
    chunk c (on its own stream) and the D2H copy of the recovered symbols + phases of chunk c-1 overlap,
    so the wall time approaches max(copy in, compute, copy out) instead of their sum.  Results land in
    ``out_host`` / ``ph_host`` with shape (nseg_total, nmodes, S + 2*bps_halo) (segment major, like the device
    layout; with a halo a segment's own symbols are columns [halo, halo + S); the last row is the end-aligned extra segment when S does not divide the capture).  With
    ``cfg.want_err`` the per-symbol training errors of every stage (what the reference returns as ``(err1, err2)``,
    ``equalisation.py:462-464``) are downloaded as well, into ``err_host`` = one pinned
    (nseg_total, nmodes, TrSyms*Niter) array per stage.  ``wxy0``: the taps every segment starts from.
    Returns (out_host, ph_host, groups) and, with ``cfg.want_err``, leaves the arrays in ``rx.err_host``."""
    cfg = rx.cfg
    nmodes, L = E_host.shape
    groups = plan_segments(L, cfg)
    nseg_total = sum(g[2] for g in groups)
    S = groups[0][1]
    H = cfg.bps_halo if groups[0][0] >= cfg.bps_halo else 0     # rows hold the halo too: own symbols are [H, H + S)
    if out_host is None:
        out_host = torch.empty((nseg_total, nmodes, S + 2 * H), dtype=rx.tdtype, pin_memory=True)
    if ph_host is None:
        ph_host = torch.empty((nseg_total, nmodes, S + 2 * H), dtype=rx.rdtype, pin_memory=True)
    if E_dev is None:
        E_dev = torch.empty((nmodes, L), dtype=rx.tdtype, device=rx.dev)
    if cfg.want_err and err_host is None:
        tr = theory.cal_training_symbol_len(cfg.os, cfg.ntaps, S * cfg.os + cfg.ntaps - 1)
        err_host = [torch.empty((nseg_total, nmodes, tr * cfg.niter[k]), dtype=rx.tdtype, pin_memory=True)
                    for k in range(len(cfg.methods))]
    rx.err_host = err_host if cfg.want_err else None
    if rx._streams is None or len(rx._streams["comp"]) < min(nchunks + 5, 21):
        # one compute stream per chunk: the training kernel is latency bound, so the chains of different
        # chunks must run side by side rather than queue behind each other
        # one download stream per kind of result (errors of stage k, recovered symbols): a stream is a queue in host
        # order (chunk by chunk), and on ONE queue the first-stage errors of chunk c+1, ready long before the
        # second-stage errors of chunk c, would wait behind them with the link idle
        rx._streams = dict(h2d=torch.cuda.Stream(), d2h=torch.cuda.Stream(),
                           d2h_err=[torch.cuda.Stream() for _ in range(len(cfg.methods))],
                           comp=[torch.cuda.Stream() for _ in range(min(nchunks + 5, 21))])
    st = rx._streams
    main = torch.cuda.current_stream()
    all_streams = [st["h2d"], st["d2h"]] + st["d2h_err"] + st["comp"]
    for s_ in all_streams:
        s_.wait_stream(main)
    events, rx.events = rx.events, None
    keep = []

    def mark(label):
        # optional timeline of the overlapped path (rx.trace): one timing event on the current stream
        if rx.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            rx.trace.append((label, ev))

    mark("start")
    copied_to = 0                                                   # samples [0, copied_to) are on the device
    for ci, (first, nsym, nseg, drop, seg0) in enumerate(_host_chunks(groups, nchunks, taper)):
        need = min(L, (first + nsym * nseg + H) * cfg.os + cfg.ntaps - 1)
        ev_in = torch.cuda.Event()
        with torch.cuda.stream(st["h2d"]):
            if need > copied_to:
                for k in range(nmodes):          # row by row: contiguous slices -> plain async DMA
                    E_dev[k, copied_to:need].copy_(E_host[k, copied_to:need], non_blocking=True)
                copied_to = need
            ev_in.record()
            mark("h2d %d" % ci)
        comp = st["comp"][ci % len(st["comp"])]
        ev_out = torch.cuda.Event()
        with torch.cuda.stream(comp):
            comp.wait_event(ev_in)
            def stage_done(stage, err, seg0=seg0, nseg=nseg):
                # a stage's training errors leave as soon as the stage is done, on their own copy stream: the download
                # (the larger direction of the link) starts a training pass after the first upload instead of a
                # whole chain after it
                if err is None:
                    return
                ev = torch.cuda.Event()
                ev.record()
                mark("train%d %d" % (stage, ci))
                with torch.cuda.stream(st["d2h_err"][stage]):
                    st["d2h_err"][stage].wait_event(ev)
                    err_host[stage][seg0:seg0 + nseg].copy_(err, non_blocking=True)
                    mark("d2h err%d %d" % (stage, ci))

            res = rx._run_group(E_dev, first, nsym, nseg, drop, wxy0, None, stage_done if cfg.want_err else None)
            ev_out.record()
            mark("chain %d" % ci)
        with torch.cuda.stream(st["d2h"]):
            st["d2h"].wait_event(ev_out)
            out_host[seg0:seg0 + nseg].copy_(res["ext"]["out"], non_blocking=True)
            ph_host[seg0:seg0 + nseg].copy_(res["ext"]["ph"], non_blocking=True)
            mark("d2h out %d" % ci)
        keep.append(res)
        if drop == 0:
            rx.host_carry = res["taps"][-1]                         # taps of the last full segment (carry_taps)
    for s_ in all_streams:
        main.wait_stream(s_)
    rx.events = events
    rx._keep = keep                                                 # alive until the caller synchronises
    return out_host, ph_host, groups


def stitch(groups, key):
    """Concatenate a per-symbol result (``eq``, ``out``, ``ph``, ``idx``) of all segments to (nmodes, N)."""
    parts = [g[key].permute(1, 0, 2).reshape(g[key].shape[1], -1)[:, g["drop"]:] for g in groups]
    return torch.cat(parts, dim=1)
