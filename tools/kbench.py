#!/usr/bin/env python
"""Developer micro-benchmark of the fused npde kernel (not the contract bench): CUDA-event timing over P."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode  # noqa: E402
from oracle import npde  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, nargs="+", default=[512, 4096, 32768, 262144])
    ap.add_argument("--M", type=int, default=5)
    ap.add_argument("--T", type=int, default=40)
    ap.add_argument("--method", default="rk4")
    ap.add_argument("--mode", default="discrete")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--lanes", type=int, default=0)
    args = ap.parse_args()
    bode._lib.load().bode_npde_set_lanes_per_pair(args.lanes)
    data = npde.make_vdp_data(seed=0, T=args.T)
    Z = npde.inducing_grid(data["Y"], args.M)
    U0 = npde.gradient_matching_init(data["Y"], data["t"].astype(np.float64), Z, 1.0, 0.75)
    rng = np.random.default_rng(1)
    stages = {"euler": 1, "midpoint": 2, "rk4": 4}[args.method]
    for P in args.P:
        U = U0[None] + 0.1 * rng.standard_normal((P, args.M ** 2, 2))
        f = bode.NPDEField(torch.from_numpy(U), torch.from_numpy(Z), 1.0, 0.75, 0.1)
        post = bode.NPDEPosterior(f, torch.from_numpy(data["x0"]), torch.from_numpy(data["t"]), torch.from_numpy(data["Y"]),
                                  method=args.method, grad_mode=args.mode)
        for _ in range(3):
            post.loss_and_grad_()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters + 1)]
        ev[0].record()
        for i in range(args.iters):
            post.loss_and_grad_()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.iters))
        med = ts[len(ts) // 2]
        steps = args.T - 1
        flop = 23.0 * stages * 5 * args.M ** 2 * steps * P
        print(json.dumps({"P": P, "M": args.M, "method": args.method, "mode": args.mode, "steps": steps,
                          "ms_med": round(med, 4), "ms_min": round(ts[0], 4),
                          "prk_steps_per_s": P * steps / (med * 1e-3), "alg_tflops": flop / (med * 1e-3) / 1e12}))
        del f, post


if __name__ == "__main__":
    main()
