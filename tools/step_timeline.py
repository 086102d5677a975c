#!/usr/bin/env python
"""Developer driver: where the kernels of one SVGD sampler step (c3) sit on the two streams.  Every C-ABI call of the step is
followed by a one-thread kernel that writes %globaltimer (bode_stamp) on the stream it was issued on; the step is captured in ONE
CUDA graph, replayed, and the stamps are read back relative to the start of the step (each stamp costs its stream ~2 us).  One process per GPU:
    python tools/step_timeline.py                                   (1 GPU)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/step_timeline.py"""
import os, sys, statistics
import ctypes as C
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200 import problems, _lib
from bayesian_ode_b200.samplers import SVGD
from bayesian_ode_b200.samplers.stein import _Workspace

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = int(os.environ.get("P", "4096"))
data = problems.make_dataset("VDP", seed=0, N=5, R=3.0, T=40, t_end=7.0, noise=0.1)
Z = problems.inducing_grid(data["Y"], 5)
U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
U = U0[None] + 0.1 * torch.randn(P, 25, 2, generator=torch.Generator().manual_seed(1234 + rank), dtype=torch.float64)
field = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
post = bode.NPDEPosterior(field, data["x0"], data["t"], torch.from_numpy(data["Y"]), method="rk4", grad_mode="discrete")
field.bind_flat_grads()
smp = SVGD([field.U, field.logsn], lr=1e-4, side_sms=int(os.environ.get("SIDE_SMS", "40")), gather_comm=os.environ.get("GATHER", "p2p"),
           median_comm=os.environ.get("MEDIAN", "p2p"))
smp.check_finite = "deferred"

marks = []          # (tag, stream name) of the step being captured; mark i writes slot i of `stamps`
stamps = torch.zeros(64, dtype=torch.int64, device="cuda")
recording = [False]


def mark(tag):
    if recording[0]:
        side = smp._side is not None and torch.cuda.current_stream() == smp._side
        _lib.check(_lib.load().bode_stamp(C.c_void_p(stamps.data_ptr() + 8 * len(marks)), _lib.stream_ptr()))
        marks.append((tag, "side" if side else "main"))


def wrap(obj, name, tag):
    f = getattr(obj, name)
    def g(*a, **k):
        r = f(*a, **k)
        mark(tag(*a, **k) if callable(tag) else tag)
        return r
    setattr(obj, name, g)


ws = smp._ws
wrap(ws, "sqdist", lambda *a, **k: "prep_x" if k.get("stages") == _lib.SVGD_PREPARE else "gram2")
wrap(ws, "median", "median")
if world > 1:
    wrap(ws, "peer_gather", lambda which, X: "gather_%s" % ("X" if which == 0 else "G"))
wrap(post, "loss_and_grad_", "solve")
lib = _lib.load()
_phi = lib.bode_svgd_phi_staged
def phi_wrapped(stages, *a):
    r = _phi(stages, *a)
    mark("prep_v" if int(stages) == _lib.SVGD_PREPARE else "phi2")
    return r
lib.bode_svgd_phi_staged = phi_wrapped
for _name in ("bode_svgd_window_select", "bode_svgd_radix_fallback"):       # the two launches of the median chain, separately
    def _mk(fn, tag):
        def w(*a):
            r = fn(*a)
            mark(tag)
            return r
        return w
    setattr(lib, _name, _mk(getattr(lib, _name), _name.replace("bode_svgd_", "")))


def step():
    mark("start")
    smp.prefetch()
    mark("solve>")
    post.loss_and_grad_()
    smp.phi(update_lr=1e-4)


for _ in range(3):
    step()
torch.cuda.synchronize()
recording[0] = True
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
recording[0] = False
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
acc = {i: [] for i in range(len(marks))}
for it in range(30):
    flush.zero_()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    st = stamps.cpu().tolist()
    for i in range(len(marks)):
        acc[i].append((st[i] - st[0]) * 1e-3)
if rank == 0:
    print("gather=%s median=%s" % (smp.gather_comm, smp.median_comm))
    print("world %d, P=%d per GPU, side_sms=%d: end of each call, us after the start of the step (median of 30 replays)" % (world, P, smp.side_sms))
    for i, (tag, s) in enumerate(marks):
        print("  %-5s %-10s %7.1f" % (s, tag, statistics.median(acc[i])))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
