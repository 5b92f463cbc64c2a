#!/usr/bin/env python
"""Generate tests/golden/convert_golden.npz from the UNMODIFIED reference
(oracle/_ref/libdetex_ref.so, built by oracle/Makefile with the reference's own flags).

For every ordered pair of pixel formats that the reference's detexConvertPixels can connect
(convert.c:887-1045) -- except chains through the two table entries whose reference function
dereferences an uninitialised pointer (convert.c:704) -- 96 seeded source pixels and the
reference's output.  Also the reference's conversion table itself (source, target pairs in
table order, read from the exported detex_conversion_table symbol) and three HDR parameter sets.

Runs only in the build container; the .npz travels with the repository.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N = 96
HDR_SETS = [(1.0, -0.25, 2.5), (1.0, 0.1, 0.7), (2.2, 0.0, 1.0)]
HDR_PAIRS = [("FLOAT_R16_HDR", "R16"), ("FLOAT_RGBX16_HDR", "RGBX16"), ("FLOAT_RG32_HDR", "FLOAT_RG32"),
             ("FLOAT_RGBX32_HDR", "RGBX16"), ("FLOAT_RG16_HDR", "RG8")]


def main():
    ref, orc = H.load_reference(), H.load_oracle()
    assert ref is not None, "compiled reference missing; run make -C oracle"
    arrays = {}
    # the table: struct { uint32 source; uint32 target; void *fn } (convert.c:757-761)
    class Entry(C.Structure):
        _fields_ = [("source", C.c_uint32), ("target", C.c_uint32), ("fn", C.c_void_p)]
    n_entries = 73
    tab = (Entry * n_entries).in_dll(ref, "detex_conversion_table")
    arrays["table"] = np.array([[e.source, e.target] for e in tab], dtype=np.uint32)
    assert all(e.fn for e in tab)
    names = list(H.PIXEL_FORMATS)
    pairs = []
    for a in names:
        for b in names:
            sf, tf = H.PIXEL_FORMATS[a], H.PIXEL_FORMATS[b]
            chain = H.conversion_chain(orc, sf, tf)
            if chain is not None and any((chain[i], chain[i + 1]) in H.REFERENCE_UB_STEPS for i in range(len(chain) - 1)):
                continue
            src = H.synth_pixels(sf, N, seed=len(pairs) + 11)
            ok, out = H.convert_pixels(ref, src, N, sf, tf)
            pairs.append((sf, tf, int(ok)))
            if ok:
                arrays["%04x_%04x/src" % (sf, tf)] = src
                arrays["%04x_%04x/out" % (sf, tf)] = out
    arrays["pairs"] = np.array(pairs, dtype=np.uint32)
    for k, (g, lo, hi) in enumerate(HDR_SETS):
        ref.detexSetHDRParameters(g, lo, hi)
        for a, b in HDR_PAIRS:
            sf, tf = H.PIXEL_FORMATS[a], H.PIXEL_FORMATS[b]
            src = H.synth_pixels(sf, N, seed=1000 + k)
            # gamma != 1: the reference inherits the thread's rounding mode (see oracle/convert_oracle.c)
            ok, out = H.convert_pixels(ref, src, N, sf, tf, primed=(g != 1.0))
            assert ok
            arrays["hdr%d/%04x_%04x/src" % (k, sf, tf)] = src
            arrays["hdr%d/%04x_%04x/out" % (k, sf, tf)] = out
    ref.detexSetHDRParameters(1.0, 0.0, 1.0)
    arrays["hdr_sets"] = np.array(HDR_SETS, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "convert_golden.npz"), **arrays)
    print("pairs:", len(pairs), "connected:", sum(p[2] for p in pairs), "table entries:", n_entries)


if __name__ == "__main__":
    main()
