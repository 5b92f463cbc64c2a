#!/usr/bin/env python
"""Device throughput of the pixel-format conversion kernel and of decode + conversion through the
L2-resident slab loop (development aid): GB/s of algorithmic traffic (source bytes + target bytes;
for decode + conversion: compressed bytes + converted bytes), CUDA-event timed."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import detex_b200 as D  # noqa: E402
from bench import measured_peak  # noqa: E402
from oracle.harness import PIXEL_FORMATS as PF, pixel_size  # noqa: E402


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pixels", type=int, default=1 << 25)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="", help="comma-separated compressed formats: decode cases only, of these formats")
    ap.add_argument("--target", default="", help="with --only: run only the case with this target (e.g. RGBX8, hdr:RGBX8, native)")
    ap.add_argument("--layout", default="", help="linear or tiled: run only this layout")
    args = ap.parse_args()
    only = [f for f in args.only.split(",") if f]
    peak, _ = measured_peak()
    n = args.pixels
    for a, b in () if only else (("RGBA8", "BGRA8"), ("RGBA8", "RGB8"), ("FLOAT_RGBX16", "FLOAT_RGB16"), ("R16", "R8"),
                 ("RGBX8", "RGBX16"), ("FLOAT_RGBX32", "FLOAT_RGBX16"), ("FLOAT_RGBX16", "RGBX8"), ("SIGNED_RG16", "RG8")):
        sf, tf = PF[a], PF[b]
        src = torch.randint(0, 255, (n * pixel_size(sf),), dtype=torch.uint8, device="cuda")
        dst = torch.empty(n * pixel_size(tf), dtype=torch.uint8, device="cuda")
        ms = timed(lambda: D.convert_pixels_device(src.data_ptr(), n, sf, dst.data_ptr(), tf), args.steps)
        gbs = n * (pixel_size(sf) + pixel_size(tf)) / ms / 1e6
        print("convert %-14s -> %-14s %8.1f us  %6.2f Gpx/s  %6.0f GB/s  %5.1f%% of measured peak" %
              (a, b, ms * 1e3, n / ms / 1e6, gbs, 100 * gbs / peak))
    wib = hib = 2048
    nb = wib * hib
    for fmt, target in (("BPTC_FLOAT", "hdr:RGBX8"), ("BC1", None), ("BC1", "BGRA8"), ("BC1", "RGB8"), ("BPTC", "RGB8"), ("BPTC_FLOAT", None),
                        ("BPTC_FLOAT", "FLOAT_RGB16"), ("BPTC_FLOAT", "RGBX8"), ("BPTC_FLOAT", "BGRA8"), ("BPTC_FLOAT", "RGB8"),
                        ("BPTC_FLOAT", "RGBX16"), ("EAC_RG11", "RG8"), ("BC3", "RGBA16")):
        if only and fmt not in only:
            continue
        if args.target and (target or "native") != args.target:
            continue
        tf = D.FORMATS[fmt]
        bs = D.block_size(tf)
        hdr = target is not None and target.startswith("hdr:")       # detexb200DecodeHDRDevice, range [0, 2]
        pf = D.native_pixel_format(tf) if target is None else PF[target.split(":")[-1]]
        blocks = torch.empty(nb * bs, dtype=torch.uint8, device="cuda")
        D.generate_blocks_device(blocks.data_ptr(), nb, bs, seed=1, force_modes=(8 if fmt == "BPTC" else 14 if fmt == "BPTC_FLOAT" else 0))
        for layout, lname in ((D.LAYOUT_LINEAR, "linear"), (D.LAYOUT_TILED, "tiled")):
            if args.layout and lname != args.layout:
                continue
            out = [torch.empty(nb * 16 * pixel_size(pf), dtype=torch.uint8, device="cuda") for _ in range(3)]
            k = [0]

            def step():
                k[0] += 1
                if hdr:
                    D.decode_hdr_device(fmt, blocks.data_ptr(), wib * 4, hib * 4, out[k[0] % 3].data_ptr(), pf,
                                        gamma=1.0, range_min=0.0, range_max=2.0, layout=layout)
                else:
                    D.decode_device(fmt, blocks.data_ptr(), wib * 4, hib * 4, out[k[0] % 3].data_ptr(), pixel_format=pf, layout=layout)
            ms = timed(step, args.steps)
            gbs = nb * (bs + 16 * pixel_size(pf)) / ms / 1e6
            print("decode  %-10s -> %-12s %-6s %8.1f us  %6.2f Gblk/s  %6.0f GB/s  %5.1f%% of measured peak" %
                  (fmt, target or "native", lname, ms * 1e3, nb / ms / 1e6, gbs, 100 * gbs / peak))


if __name__ == "__main__":
    main()
