"""Helper of tests/test_gpu_parity.py::test_bc7_flow_kernel: run in a child process with
DETEXB200_BC7_KERNEL=flow (the library reads the variable once per process) and compare the BC7
dataflow kernel with the CPU checker: distributions U and M, every single mode, both layouts,
sizes that leave a partial last tile and tiles that wrap over block rows, invalid blocks, the
fused BGRA8 / RGB8 targets.  Prints OK <n> on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import detex_b200 as D  # noqa: E402
from oracle import harness as H  # noqa: E402


def main():
    assert os.environ.get("DETEXB200_BC7_KERNEL") == "flow"
    cpu = H.best_cpu_decoder()[0]
    n_checked = 0
    cases = []
    base = H.synth_blocks("BPTC", 131072, seed=9)
    cases.append((base, 2048, 1024))                                     # U: invalid blocks, mode 0 heavy
    cases.append((H.force_modes("BPTC", base), 1024, 2048))            # M
    for m in range(8):
        cases.append((H.force_modes("BPTC", base[:16 * 8192], [m]), 512, 256))
    cases.append((H.force_modes("BPTC", base[:16 * 1500]), 120, 200))   # 30 x 50 blocks: partial tile, wib < tile
    cases.append((base[:16 * 5000], 400, 200))                           # 100 x 50 blocks
    for blocks, w, h in cases:
        for linear in (True, False):
            want, ok_w, _ = H.cpu_decode(cpu, "BPTC", blocks, w, h, linear=linear, threads=4)
            fn = D.detexDecompressTextureLinear if linear else D.detexDecompressTextureTiled
            got, ok_g = fn("BPTC", blocks, w, h)
            assert ok_g == ok_w and np.array_equal(got, want), ("BPTC", w, h, linear)
            n_checked += 1
    m_blocks = H.force_modes("BPTC", base[:16 * 16384])
    for target in ("BGRA8", "RGB8", "RGBX8"):
        pf = D.PIXEL_FORMATS[target]
        for linear in (True, False):
            want, ok_w, _ = H.cpu_decode(cpu, "BPTC", m_blocks, 512, 512, linear=linear, pixel_format=pf)
            fn = D.detexDecompressTextureLinear if linear else D.detexDecompressTextureTiled
            got, ok_g = fn("BPTC", m_blocks, 512, 512, pixel_format=pf)
            assert ok_g == ok_w and np.array_equal(got, want), (target, linear)
            n_checked += 1
    print("OK %d" % n_checked)


if __name__ == "__main__":
    main()
