"""Multi-GPU partitioning of the hot path: one process per GPU over torch.distributed.

* Streams and pairs are independent units: `round_robin` assigns them to ranks, no data-path
  collective (the reference's only parallelism is an in-process goroutine pool,
  fingerprint/analyzers/spectral.go:234-285,443-517).
* ONE long cross-correlation shards by lag range: every rank holds both (small) sequences, evaluates
  its lags, and the ranks exchange 16 B each (|peak|, global lag index) with an all-gather; the global
  arg-max follows findPeak's rule (larger |c|, ties -> smaller index, algorithms/stats/correlation.go:
  535-541).  A second all-gather of 9 doubles per rank carries the peak-relative partial sums
  (noise power, side lobe, second peak, the peak's neighbours) for SNR / sharpness / peak-to-sidelobe.
The collective backend is whatever the process group uses: NCCL over NVLink on the GPU box, gloo in
the CPU tests.
"""
from __future__ import annotations

import ctypes as C

from . import capi


def round_robin(n_items: int, world: int, rank: int) -> list[int]:
    """Indices of the streams / pairs rank `rank` owns (same rule as the C library uses across devices)."""
    return list(range(rank, n_items, world))


def lag_range(n_lags: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous lag-index range [lo, hi) of rank `rank`: ceil(n_lags / world) lags each."""
    step = -(-n_lags // world)
    lo = min(n_lags, rank * step)
    return lo, min(n_lags, lo + step)


def actual_max_lag(max_lag: int, na: int, nb: int) -> int:
    return max(0, min(max_lag, na - 1, nb - 1))  # correlation.go:452-461


def nccl_setup(lib: "capi.SonarLib", device=None, group=None):
    """Gives the C library its own NCCL communicator over the ranks of `group`: rank 0 draws the unique id
    (sonar_nccl_unique_id), torch.distributed only carries those 128 bytes, every rank calls sonar_nccl_init.  After this
    `lib.xcorr_lag_sharded` runs the whole sharded correlation inside the library (one ncclAllGather, no host hops)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid = torch.frombuffer(bytearray(lib.nccl_unique_id()), dtype=torch.uint8).to(uid.device)
    dist.broadcast(uid, src=0, group=group)
    lib.nccl_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    return world, rank


def xcorr_lag_sharded(lib: "capi.SonarLib", a, b, max_lag: int, device=None, group=None):
    """CrossCorrelation.Compute of ONE pair with the lags split over the ranks of `group`.

    Returns (summary, (lo, hi), local_corr): every rank gets the identical global summary.
    """
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nl = 2 * actual_max_lag(max_lag, len(a), len(b)) + 1
    lo, hi = lag_range(nl, world, rank)
    sh, pk = lib.xcorr_shard(a, b, max_lag, lo, hi)
    try:
        mine = torch.tensor([pk.abs_peak, float(pk.index)], dtype=torch.float64, device=device)  # 16 B / rank
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine, group=group)
        peaks = [capi.XcorrShardPeak(float(g[0]), int(g[1])) for g in (t.cpu() for t in got)]
        gidx = lib.xcorr_merge_peaks(peaks)
        m = lib.xcorr_shard_metrics(sh, gidx)
        names = [n for n, _ in capi.XcorrShardMetrics._fields_]
        mine = torch.tensor([getattr(m, n) for n in names], dtype=torch.float64, device=device)
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine, group=group)
        parts = []
        for t in got:
            pm = capi.XcorrShardMetrics()
            for n, v in zip(names, t.cpu().tolist()):
                setattr(pm, n, v)
            parts.append(pm)
        summary = lib.xcorr_merge_metrics(parts, len(a), len(b), max_lag, gidx)
        local = lib.xcorr_shard_corr(sh, hi - lo)
    finally:
        lib.xcorr_shard_close(sh)
    return summary, (lo, hi), local
