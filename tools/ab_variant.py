#!/usr/bin/env python
"""A/B helper for kernel experiments that are compile-time switches (-DDETEXB200_...).

  python tools/ab_variant.py build table2 -DDETEXB200_BC7_TABLE2        # -> detex_b200/libexp_table2.so (travels with gpurun)
  python tools/ab_variant.py run table2 tools/sweep.py --formats BPTC --modes both
  python tools/ab_variant.py run table2 -m pytest tests/test_gpu_parity.py -m gpu -q

`run` executes the given script (or `-m module`) in this process with detex_b200.lib.LIB_PATH pointing at the
variant, so tools, the benchmark and the GPU tests exercise it unchanged.  Variant libraries are git-ignored
(*.so); delete them when the experiment is decided."""
import os
import runpy
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def variant_path(name):
    return os.path.join(ROOT, "detex_b200", "libexp_%s.so" % name)


def main():
    if len(sys.argv) < 3 or sys.argv[1] not in ("build", "run"):
        sys.exit(__doc__)
    cmd, name, rest = sys.argv[1], sys.argv[2], sys.argv[3:]
    if cmd == "build":
        from detex_b200 import build as B
        nvcc = B.find_nvcc()
        if nvcc is None:
            sys.exit("nvcc not found")
        out = variant_path(name)
        subprocess.run([nvcc] + B.NVCC_FLAGS + rest + ["-o", out] + [os.path.join(B.CSRC, s) for s in B.SOURCES],
                       check=True, cwd=B.CSRC)
        print(out)
        return
    lib = variant_path(name)
    if not os.path.exists(lib):
        sys.exit("no such variant library: " + lib)
    import detex_b200.lib as L
    L.LIB_PATH = lib
    if rest and rest[0] == "-m":
        sys.argv = rest[1:]
        if rest[1] == "pytest":
            import pytest
            sys.exit(pytest.main(rest[2:]))
        runpy.run_module(rest[1], run_name="__main__")
    else:
        sys.argv = rest
        runpy.run_path(rest[0], run_name="__main__")


if __name__ == "__main__":
    main()
