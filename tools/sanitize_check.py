#!/usr/bin/env python
"""Small decode workload for compute-sanitizer (memcheck / racecheck / synccheck): the mode-sorted
kernels (BC7, BC6H), the BC7 dataflow kernel when DETEXB200_BC7_KERNEL=flow, and one plain kernel per
layout, each on a few tiles with a partial last tile, compared with the CPU checker.

  compute-sanitizer --tool racecheck python tools/sanitize_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import detex_b200 as D  # noqa: E402
from oracle import harness as H  # noqa: E402


def converted_targets(cpu):
    """the staged / fused variants of the sorted kernels: BC6H -> 8-bit (tone classes, staged LINEAR), -> RGBX16,
    -> FLOAT_RGB16 (packed half tiles, 512-block tiles, unpadded bins), BC7 -> RGB8 (staged in both layouts)"""
    PF = H.PIXEL_FORMATS
    n = 0
    # the restatement is the checker here: the compiled reference's FLOAT_RGBX16 -> FLOAT_RGB16 step writes
    # through an uninitialised pointer (convert.c:704, harness.REFERENCE_UB_STEPS)
    cpu = H.load_oracle()
    for fmt, w, h, targets in (("BPTC_FLOAT", 256, 264, ("RGBX8", "RGB8", "BGRA8", "RGBX16", "FLOAT_RGB16")),
                               ("BPTC", 256, 200, ("RGB8", "BGRA8"))):
        nb = (w // 4) * (h // 4)
        blocks = H.synth_blocks(fmt, nb, seed=5)
        blocks = np.concatenate([H.force_modes(fmt, blocks[:16 * (nb // 2)]), blocks[16 * (nb // 2):]])
        for t in targets:
            for linear in (True, False):
                want, ok_w, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=linear, pixel_format=PF[t])
                fn = D.detexDecompressTextureLinear if linear else D.detexDecompressTextureTiled
                got, ok_g = fn(fmt, blocks, w, h, pixel_format=PF[t])
                assert ok_g == ok_w and np.array_equal(got, want), (fmt, t, linear)
                n += 1
    return n


def main():
    cpu = H.best_cpu_decoder()[0]
    n = 0
    if "--converted-only" in sys.argv:
        print("sanitize_check OK: %d converted decodes" % converted_targets(cpu))
        return
    for fmt, w, h in (("BPTC", 256, 200), ("BPTC_FLOAT", 256, 264), ("BPTC_SIGNED_FLOAT", 512, 132),
                      ("ETC2", 128, 64), ("BC1", 128, 64)):
        nb = (w // 4) * (h // 4)
        blocks = H.synth_blocks(fmt, nb, seed=3)
        if fmt.startswith("BPTC"):
            blocks = np.concatenate([H.force_modes(fmt, blocks[:16 * (nb // 2)]), blocks[16 * (nb // 2):]])
        for linear in (True, False):
            want, ok_w, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=linear)
            fn = D.detexDecompressTextureLinear if linear else D.detexDecompressTextureTiled
            got, ok_g = fn(fmt, blocks, w, h)
            assert ok_g == ok_w and np.array_equal(got, want), (fmt, linear)
            n += 1
    n += converted_targets(cpu)
    print("sanitize_check OK: %d decodes, BC7 kernel = %s" % (n, os.environ.get("DETEXB200_BC7_KERNEL", "sorted")))


if __name__ == "__main__":
    main()
