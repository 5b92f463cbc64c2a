#!/usr/bin/env python
"""Device-resident throughput of every format in both layouts (development aid, not the bench
of record): blocks/s and fraction of the measured HBM peak, CUDA-event timed, rotating buffers."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import detex_b200 as D  # noqa: E402
from bench import ALGO_BYTES, measured_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--formats", default=",".join(D.FORMAT_NAMES))
    ap.add_argument("--blocks", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--modes", default="U", help="U, M or both (M only affects BC7/BC6H)")
    ap.add_argument("--bc7-single-modes", default="", help="comma list of BC7 modes: one run per mode with EVERY block in that mode")
    args = ap.parse_args()
    peak, _ = measured_peak()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    rows = []
    for fmt in args.formats.split(","):
        tf = D.FORMATS[fmt]
        bs, ps = D.block_size(tf), D.pixel_size(tf & 0xFFFF)
        n = args.blocks
        wib = 2048
        hib = n // wib
        dists = ["U"]
        if fmt in ("BPTC", "BPTC_FLOAT", "BPTC_SIGNED_FLOAT") and args.modes in ("M", "both"):
            dists = ["M"] if args.modes == "M" else ["U", "M"]
        if fmt == "BPTC" and args.bc7_single_modes:
            dists += ["m%s" % m for m in args.bc7_single_modes.split(",")]
        for dist in dists:
            force = 0 if dist == "U" else (8 if fmt == "BPTC" else 14)
            if dist.startswith("m"):
                force = 100 + int(dist[1:])
            n_sets = max(2, int(600e6 // (n * (bs + 16 * ps))) + 1)
            sets = []
            for k in range(n_sets):
                b = torch.empty(n * bs, dtype=torch.uint8, device=dev)
                D.generate_blocks_device(b.data_ptr(), n, bs, seed=1 + k, force_modes=force, stream=stream.cuda_stream)
                sets.append((b, torch.empty(n * 16 * ps, dtype=torch.uint8, device=dev)))
            for layout, lname in ((D.LAYOUT_LINEAR, "linear"), (D.LAYOUT_TILED, "tiled")):
                def step(k):
                    b, o = sets[k % n_sets]
                    D.decode_device(tf, b.data_ptr(), wib * 4, hib * 4, o.data_ptr(), layout=layout, stream=stream.cuda_stream)
                for k in range(3):
                    step(k)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for k in range(args.steps):
                    step(k)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                gbs = n * ALGO_BYTES[fmt] / ms / 1e6
                rows.append(dict(format=fmt, dist=dist, layout=lname, us=ms * 1e3, gblocks_s=n / ms / 1e6, gbs=gbs, frac=gbs / peak))
                print("%-20s %-2s %-6s %8.1f us  %7.2f Gblk/s  %7.0f GB/s  %5.1f%% of measured peak" %
                      (fmt, dist, lname, ms * 1e3, n / ms / 1e6, gbs, 100 * gbs / peak), flush=True)
            del sets
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
