"""Helper of tests/test_gpu_parity.py::test_bc6h_kernel_structures: run in a child process with
DETEXB200_BC6H_SPLIT=0 or 1 (the library reads the variable once per process).  BC6H has two structures of the
mode-sorted kernel -- the one-piece decoder (direct, staged or TMA-stored output) and the SPLIT one (mode-dependent
half on sorted groups, mode-independent half in image order) -- and the default picks per target format; the switch
forces one of them for every BC6H launch.  Compared with the CPU checker: both formats, distributions U (invalid
blocks) and M, both layouts, the fused targets, sizes with partial tiles / tiles that wrap over block rows / rows
of whole 64-block boxes (TMA store) / a clipped last block row.  Prints OK <n> on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import detex_b200 as D  # noqa: E402
from oracle import harness as H  # noqa: E402


def main():
    assert os.environ.get("DETEXB200_BC6H_SPLIT") in ("0", "1")
    # the restatement is the checker: the compiled reference's FLOAT_RGBX16 -> FLOAT_RGB16 step writes through an
    # uninitialised pointer (convert.c:704, harness.REFERENCE_UB_STEPS)
    cpu = H.load_oracle()
    n_checked = 0
    sizes = ((1024, 512), (256, 1024), (120, 200), (400, 200), (2048, 30))
    for fmt, targets in (("BPTC_FLOAT", (None, "FLOAT_RGB16", "RGBX8", "BGRA8", "RGB8", "RGBX16")), ("BPTC_SIGNED_FLOAT", (None,))):
        for w, h in sizes:
            n = ((w + 3) // 4) * ((h + 3) // 4)
            base = H.synth_blocks(fmt, n, seed=w + h)
            for blocks in (base, H.force_modes(fmt, base)):
                for target in targets:
                    pf = None if target is None else D.PIXEL_FORMATS[target]
                    for linear in (True, False):
                        want, ok_w, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=linear, pixel_format=pf, threads=4)
                        fn = D.detexDecompressTextureLinear if linear else D.detexDecompressTextureTiled
                        got, ok_g = fn(fmt, blocks, w, h, pixel_format=pf)
                        assert ok_g == ok_w and np.array_equal(got, want), (fmt, target, w, h, linear)
                        n_checked += 1
    print("OK %d" % n_checked)


if __name__ == "__main__":
    main()
