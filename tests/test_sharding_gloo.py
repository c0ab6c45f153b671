"""N > 1 host logic on CPU: world_size-2 gloo process groups driving the lag-shard protocol
 with the oracle as the per-rank compute, plus the stream/pair partitioning rules."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle", "libsonar_oracle.so")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("sonido-sonar_b200")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = pkg.capi.SonarLib(ORACLE)
    rng = np.random.default_rng(7)  # same inputs on every rank (both sequences are replicated)
    base = np.convolve(rng.standard_normal(5000), np.ones(16) / 16, mode="same")
    a, b = base[300:3300], base[300 - 141:3300 - 141] + 0.05 * rng.standard_normal(3000)
    summ, (lo, hi), local = pkg.sharding.xcorr_lag_sharded(lib, a, b, 700)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), lo=lo, hi=hi, local=local,
             **{k: v for k, v in summ.as_dict().items()})
    # stream partition: every index owned exactly once
    mine = pkg.sharding.round_robin(11, world, rank)
    import torch
    t = torch.zeros(11, dtype=torch.int64)
    t[mine] = 1
    dist.all_reduce(t)
    assert t.tolist() == [1] * 11
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_lag_sharded_xcorr_matches_unsharded(tmp_path, oracle, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    base = np.convolve(rng.standard_normal(5000), np.ones(16) / 16, mode="same")
    a, b = base[300:3300], base[300 - 141:3300 - 141] + 0.05 * rng.standard_normal(3000)
    corr, whole = oracle.xcorr(a, b, 700)
    assert whole.peak_lag == 141
    pieces = []
    for r in range(world):
        g = np.load(os.path.join(tmp_path, f"r{r}.npz"))
        for k, v in whole.as_dict().items():
            if k == "n_candidates":
                continue
            got = g[k].item()
            assert got == pytest.approx(v, rel=1e-12, abs=0), (k, got, v)
        pieces.append((int(g["lo"]), g["local"]))
    assert np.array_equal(np.concatenate([p[1] for p in sorted(pieces, key=lambda x: x[0])]), corr)


def test_partition_rules(pkg):
    sh = pkg.sharding
    assert [sh.lag_range(20671, 8, r) for r in (0, 7)] == [(0, 2584), (18088, 20671)]
    assert sh.lag_range(5, 8, 7) == (5, 5)  # empty shard is legal
    assert sorted(sum((sh.round_robin(4096, 8, r) for r in range(8)), [])) == list(range(4096))
    assert sh.actual_max_lag(10335, 51676, 51676) == 10335 and sh.actual_max_lag(10335, 100, 5000) == 99
