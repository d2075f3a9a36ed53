"""Multi-rank path on CPU: world_size 2 over gloo.  Each rank owns a contiguous block of whole
segments of one long capture (ntaps-1 samples of overlap, NO exchange on the data path); the CPU
oracle stands in for the CUDA kernels.  Checks that the per-rank ranges tile the capture, that the
union of the per-rank results is identical to the single-process result, and the max-over-ranks
timing reduction bench.py uses."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_oracle as co
    from qampy_b200 import pipeline, synth, theory
    M, ntaps, S, nsym = 16, 21, 1024, 9000
    E, _ = synth.synth_numpy(M, nsym, seed=5, snr_db=25.0)          # every rank can regenerate the capture
    a, b, s0, s1 = pipeline.rank_capture_range(nsym, ntaps, 2, S, rank, world)
    mine = E[:, a:b]                                                  # what this rank would hold in HBM
    cfg = pipeline.ReceiverConfig(M=M, ntaps=ntaps, seg_symbols=S, bps_angles=32, bps_N=21)
    outs = []
    for first, n, nseg, drop in pipeline.plan_segments(mine.shape[1], cfg):
        for s in range(nseg):
            seg = mine[:, (first + s * n) * 2: (first + s * n) * 2 + n * 2 + ntaps - 1]
            eq, w, _ = co.dual_mode_equalisation(seg, 2, (1e-3, 1e-3), M, Ntaps=ntaps, methods=("mcma", "mrde"))
            outs.append(eq[:, drop:] if nseg == 1 and drop else eq)
    out = np.concatenate(outs, axis=1)
    assert out.shape[1] == s1 - s0
    # data-path results are never exchanged; only for the test do we collect them on rank 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (s0, s1, out))
    t = pipeline.max_over_ranks(10.0 + 5.0 * rank)                    # slowest rank defines the step time
    if rank == 0:
        np.savez(os.path.join(outdir, "res.npz"), t=t, **{"out%d" % r: g[2] for r, g in enumerate(gathered)},
                 spans=np.array([[g[0], g[1]] for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_tile_the_capture_without_exchange():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_oracle as co
    from qampy_b200 import pipeline, synth
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, d), nprocs=world, join=True)
        res = np.load(os.path.join(d, "res.npz"))
        assert float(res["t"]) == 15.0
        spans = res["spans"]
        M, ntaps, S, nsym = 16, 21, 1024, 9000
        N = (nsym * 2 - ntaps + 1) // 2
        assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == N
        # single-process result over the same segmentation
        E, _ = synth.synth_numpy(M, nsym, seed=5, snr_db=25.0)
        cfg = pipeline.ReceiverConfig(M=M, ntaps=ntaps, seg_symbols=S, bps_angles=32, bps_N=21)
        ref = np.zeros((2, N), np.complex64)
        nfull = N // S
        for s in range(nfull):
            seg = E[:, s * S * 2: s * S * 2 + S * 2 + ntaps - 1]
            ref[:, s * S:(s + 1) * S] = co.dual_mode_equalisation(seg, 2, (1e-3, 1e-3), M, Ntaps=ntaps,
                                                                  methods=("mcma", "mrde"))[0]
        both = np.concatenate([res["out0"], res["out1"]], axis=1)
        assert both.shape == (2, N)
        # all whole segments are bit-identical to the single-process run (same code, same inputs)
        assert np.array_equal(both[:, :nfull * S], ref[:, :nfull * S])
        assert np.isfinite(both).all()
