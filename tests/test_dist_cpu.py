"""CPU, world_size 2, gloo: the host-side logic of the sharded path -- contiguous particle sharding, the lock-step
exact-median protocol (histogram all-reduce between radix passes) and the max-over-ranks reduction bench.py uses.
The CUDA kernels are replaced by NumPy stand-ins that implement the same per-pass contract as bode_svgd_hist_pass /
bode_svgd_select_digit (include/bode_b200.h), so what is exercised is the protocol and the collectives."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bayesian_ode_b200  # noqa: F401
    from bayesian_ode_b200.samplers.stein import radix_select_protocol, shard_range
    from oracle import samplers as osamp

    n, d = 96, 7
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d)).astype(np.float32)
    lo, hi = shard_range(n, rank, world)
    # "all-gather" of the positions: every rank contributes its block
    parts = [torch.zeros(shard_range(n, r, world)[1] - shard_range(n, r, world)[0], d) for r in range(world)]
    dist.all_gather(parts, torch.from_numpy(X[lo:hi]))
    Xall = torch.cat(parts).numpy()
    assert np.array_equal(Xall, X)
    d2 = osamp.sq_dists(X[lo:hi].astype(np.float64), Xall.astype(np.float64)).astype(np.float32)      # local rows x all columns
    bits = d2.view(np.uint32).ravel()

    windows = [(20, 11), (9, 11), (0, 9)]
    total = n * n
    st = dict(prefix=[0, 0], rank=[(total - 1) // 2, total // 2])
    hist = torch.zeros(2, 2048, dtype=torch.int64)

    def hist_pass(ps):
        shift, nb = windows[ps]
        himask = 0 if ps == 0 else (0xFFFFFFFF << (shift + nb)) & 0xFFFFFFFF
        hist.zero_()
        for w in range(2):
            if w == 1 and st["prefix"][0] == st["prefix"][1]:
                continue
            sel = (bits & himask) == st["prefix"][w]
            dig = (bits[sel] >> shift) & ((1 << nb) - 1)
            hist[w] += torch.from_numpy(np.bincount(dig, minlength=2048).astype(np.int64))

    def select_digit(ps):
        shift, nb = windows[ps]
        two = st["prefix"][0] != st["prefix"][1]
        newp, newr = [], []
        for w in range(2):
            h = hist[1 if (w == 1 and two) else 0].numpy()
            cum = np.cumsum(h)
            dsel = int(np.searchsorted(cum, st["rank"][w], side="right"))
            newp.append(st["prefix"][w] | (dsel << shift))
            newr.append(st["rank"][w] - (int(cum[dsel - 1]) if dsel > 0 else 0))
        st["prefix"], st["rank"] = newp, newr

    radix_select_protocol(hist_pass, lambda: dist.all_reduce(hist), select_digit)
    a, b = (np.array([p], dtype=np.uint32).view(np.float32)[0] for p in st["prefix"])
    med = np.float32(0.5) * (a + b)
    full = osamp.sq_dists(X.astype(np.float64), X.astype(np.float64)).astype(np.float32)
    assert med == np.median(full), (med, np.median(full))
    # max-over-ranks timing reduction (bench.py) and a final gather of per-rank results
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(world)
    ret[rank] = float(med)
    dist.barrier()
    dist.destroy_process_group()


def _main():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29641, ret), nprocs=world, join=True)
    assert len(ret) == world and ret[0] == ret[1]
    print("GLOO_OK", ret[0])


def test_sharded_median_protocol_world2_gloo():
    """Runs in its own interpreter: forking / process-group state must not leak into the pytest process (a BLAS call
    after an in-process mp.Manager() fork was seen to deadlock)."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.abspath(__file__)], capture_output=True, text=True, timeout=240)
    assert "GLOO_OK" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


def test_shard_range_partitions_exactly():
    import bayesian_ode_b200  # noqa: F401
    from bayesian_ode_b200.samplers.stein import shard_range
    for n in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


if __name__ == "__main__":
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    _main()
