#!/usr/bin/env python
"""End-to-end throughput of ONE detexDecompressTextureLinear call spread over several GPUs from one process
(detexb200SetDevices, SURVEY.md 8e behind the C ABI): blocks/s for device sets of 1, 2, 4, ... GPUs with pinned and
with pageable (malloc'd) caller buffers, next to the plain cudaMemcpy ceiling of the same bytes over the same
devices.  Run under `gpurun --gpus N`:

  python tools/e2e_devices.py [--workloads bc1,bc7] [--steps 5]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import detex_b200 as D  # noqa: E402
from oracle import harness as H  # noqa: E402

WORK = {"bc1": ("BC1", 8192, 8192, 0), "bc7": ("BPTC", 16384, 16384, 8), "bc6h": ("BPTC_FLOAT", 8192, 8192, 14)}


def memcpy_ceiling(devs, in_bytes, out_bytes, steps):
    """H2D of in_bytes and D2H of out_bytes split evenly over the devices, all at once, pinned"""
    parts = []
    for d in devs:
        with torch.cuda.device(d):
            hi = torch.empty(in_bytes // len(devs), dtype=torch.uint8).pin_memory()
            ho = torch.empty(out_bytes // len(devs), dtype=torch.uint8).pin_memory()
            di = torch.empty_like(hi, device="cuda:%d" % d)
            do = torch.empty_like(ho, device="cuda:%d" % d)
            parts.append((d, hi, ho, di, do, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))

    def step():
        for d, hi, ho, di, do, s1, s2 in parts:
            with torch.cuda.stream(s1):
                di.copy_(hi, non_blocking=True)
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
        for p in parts:
            p[5].synchronize()
            p[6].synchronize()
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="bc1,bc7")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    lib = D.load()
    n_dev = lib.detexb200DeviceCount()
    if args.threads:
        D.set_host_threads(args.threads)
    cpu = H.best_cpu_decoder()[0]
    sets = [[0]] + [list(range(k)) for k in (2, 4, 8) if k <= n_dev]
    rows = []
    for key in args.workloads.split(","):
        fmt, w, h, force = WORK[key]
        tf, bs, ps = H.FORMATS[fmt]
        n = (w // 4) * (h // 4)
        blocks = H.synth_blocks(fmt, n, seed=1)
        if force:
            blocks = H.force_modes(fmt, blocks)
        want, ok_w, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=True, threads=os.cpu_count() or 1)
        n_out = w * h * ps
        pin_in = torch.from_numpy(blocks.copy()).pin_memory()
        pin_out = torch.zeros(n_out, dtype=torch.uint8).pin_memory()
        pag_in = blocks.copy()
        pag_out = np.zeros(n_out, dtype=np.uint8)
        for devs in sets:
            D.set_devices(devs if len(devs) > 1 else [])       # one device: the plain per-thread path
            res = {"workload": key, "devices": len(devs)}
            for kind, ip, op, view in (("pinned", pin_in.data_ptr(), pin_out.data_ptr(), pin_out.numpy()),
                                       ("pageable", pag_in.ctypes.data, pag_out.ctypes.data, pag_out)):
                tex = D.DetexTexture(tf, ip, w, h, w // 4, h // 4)
                view[:] = 0
                for _ in range(2):
                    ok = lib.detexDecompressTextureLinear(C.byref(tex), op, tf & 0xFFFF)
                assert bool(ok) == ok_w and np.array_equal(view, want), (key, devs, kind)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    lib.detexDecompressTextureLinear(C.byref(tex), op, tf & 0xFFFF)
                dt = (time.perf_counter() - t0) / args.steps
                res[kind + "_mblocks_s"] = n / dt / 1e6
                res[kind + "_gbs_d2h"] = n_out / dt / 1e9
            dt = memcpy_ceiling(devs, n * bs, n_out, args.steps)
            res["memcpy_ceiling_mblocks_s"] = n / dt / 1e6
            res["memcpy_ceiling_gbs_d2h"] = n_out / dt / 1e9
            rows.append(res)
            print("%-5s %d GPU: pinned %8.1f M blocks/s (%5.1f GB/s D2H)  pageable %8.1f  memcpy ceiling %8.1f (%5.1f GB/s)  bit-exact" %
                  (key, len(devs), res["pinned_mblocks_s"], res["pinned_gbs_d2h"], res["pageable_mblocks_s"],
                   res["memcpy_ceiling_mblocks_s"], res["memcpy_ceiling_gbs_d2h"]), flush=True)
    D.set_devices([])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"host_cores": os.cpu_count(), "gpus": n_dev, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "e2e_devices.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
