"""GPU (>= 2 devices): the particle-sharded SVGD step equals the single-GPU step on the concatenated particles
(all-gather of [theta | grad], distributed exact median via histogram all-reduce).  Run by hand / by the driver with
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/test_multigpu.py
and through pytest (which spawns that command) when two GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    from bayesian_ode_b200.samplers import SVGD
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    data = problems.make_dataset(seed=0)
    Z = problems.inducing_grid(data["Y"], 5)
    U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
    P = 256
    U = U0[None] + 0.1 * torch.randn(P, 25, 2, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    lo, hi = rank * P // world, (rank + 1) * P // world

    def run(Uslice, distributed, median_comm="p2p", gather_comm="p2p"):
        f = bode.NPDEField(Uslice, Z, 1.0, 0.75, 0.1)
        post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
        f.bind_flat_grads()
        smp = SVGD([f.U, f.logsn], lr=1e-4, median_comm=median_comm, gather_comm=gather_comm)
        if distributed:
            assert smp.median_comm == median_comm, smp.median_comm
            assert smp.gather_comm == (gather_comm if median_comm == "p2p" else "nccl"), smp.gather_comm
        if not distributed:
            smp.world, smp.rank, smp.n_total = 1, 0, smp.P_local
        for it in range(3):
            if it != 1:
                smp.prefetch()            # positions gathered + Gram operands built on the side stream, beside the solve
            post.loss_and_grad_()
            smp.phi(update_lr=1e-4)
        return f.theta.clone(), smp._ws.med_gamma.clone()

    th_d, mg_d = run(U[lo:hi], True, "p2p")          # peer-mapped workspaces: push-kernel gathers, NVLink reads + flag barriers in the median kernels
    th_g, mg_g = run(U[lo:hi], True, "p2p", "nccl")  # peer-memory median, NCCL all-gathers
    th_n, mg_n = run(U[lo:hi], True, "nccl")         # collective protocol throughout
    assert torch.equal(mg_d, mg_n) and torch.equal(th_d, th_n)
    assert torch.equal(mg_d, mg_g) and torch.equal(th_d, th_g)
    # single-GPU reference on every rank (world forced to 1 before the workspace is built)
    f = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
    f.bind_flat_grads()
    w = dist.get_world_size
    try:
        torch.distributed.is_initialized_backup = torch.distributed.is_initialized
        torch.distributed.is_initialized = lambda: False
        smp = SVGD([f.U, f.logsn], lr=1e-4)
    finally:
        torch.distributed.is_initialized = torch.distributed.is_initialized_backup
    for _ in range(3):
        post.loss_and_grad_()
        smp.phi(update_lr=1e-4)
    th_s, mg_s = f.theta[lo:hi], smp._ws.med_gamma
    assert torch.equal(mg_d, mg_s), (mg_d, mg_s)                      # bit-exact distributed median
    err = float((th_d - th_s).abs().max() / th_s.abs().max())
    assert err < 1e-6, err
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_OK world=%d err=%.2e" % (world, err))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_svgd_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29731", os.path.abspath(__file__)], capture_output=True, text=True, timeout=600)
    assert "MULTIGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


if __name__ == "__main__":
    _worker()
