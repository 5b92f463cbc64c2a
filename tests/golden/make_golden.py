#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference for the bundled KTX fixtures
and oracle/_ref/libdetex_ref.so, the reference compiled with its own flags by
oracle/Makefile).  Outputs, all under tests/golden/:

  fixtures.npz            for each bundled test-texture-<FMT>.ktx (reference
                          validate.c:31-57 lists them): the compressed payload as loaded
                          by the reference's detexLoadKTXFile and the pixels produced by
                          the reference's detexDecompressTextureLinear in the native format
  synthetic_golden.json   FNV-1a-64 of reference outputs on seeded synthetic blocks
                          (splitmix64 generator of SURVEY.md 8c): per format, LINEAR and
                          TILED layouts, invalid-block counts, per-mode forced sets for
                          BC7 / BC6H, clipped (non multiple of 4) sizes
  reference_enum_values.txt  numeric value of every DETEX_* constant (from detex.h)

The SURVEY.md 8c hash tables could not be reproduced with a standard FNV-1a-64 (not
even the input hash, which only depends on the generator); the invalid-block counts
there DO reproduce exactly, so the generator is the same and these files are the pins.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REFDIR = "/root/reference"
KTX = ["BC1", "BC1A", "BC2", "BC3", "BPTC", "BPTC_FLOAT", "RGTC1", "RGTC2", "SIGNED_RGTC1",
       "SIGNED_RGTC2", "ETC1", "ETC2", "ETC2_PUNCHTHROUGH", "ETC2_EAC", "EAC_R11", "EAC_RG11",
       "EAC_SIGNED_R11"]
CLIPPED_SIZES = [(1021, 37), (5, 3), (1, 1), (130, 67), (64, 61)]


def main():
    ref = H.load_reference()
    assert ref is not None, "compiled reference missing; run make -C oracle"
    ref.detexLoadKTXFile.argtypes = [C.c_char_p, C.POINTER(C.POINTER(H.DetexTexture))]
    ref.detexLoadKTXFile.restype = C.c_bool
    arrays = {}
    for name in KTX:
        t = C.POINTER(H.DetexTexture)()
        assert ref.detexLoadKTXFile(("%s/test-texture-%s.ktx" % (REFDIR, name)).encode(), C.byref(t))
        tx = t.contents
        tf, bs, ps = H.FORMATS[name]
        assert tx.format == tf
        n = tx.width_in_blocks * tx.height_in_blocks * bs
        data = np.ctypeslib.as_array(C.cast(tx.data, C.POINTER(C.c_uint8)), shape=(n,)).copy()
        out = np.zeros(tx.width * tx.height * ps, dtype=np.uint8)
        ok = ref.detexDecompressTextureLinear(t, out.ctypes.data, tf & 0xFFFF)
        arrays[name + "/blocks"] = data
        arrays[name + "/pixels"] = out
        arrays[name + "/meta"] = np.array([tx.width, tx.height, int(ok)], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "fixtures.npz"), **arrays)

    gold = {"generator": "splitmix64 counter, SURVEY.md 8c", "hash": "FNV-1a-64 of the output buffer",
            "uniform": {}, "forced_modes": {}, "clipped": {}}
    n = 4096  # 64 x 64 blocks = 256 x 256 px
    for f in H.FORMAT_NAMES:
        b = H.synth_blocks(f, n, seed=1)
        lin, ok, _ = H.cpu_decode(ref, f, b, 256, 256, linear=True)
        til, ok2, _ = H.cpu_decode(ref, f, b, 256, 256, linear=False)
        tile_bytes = 16 * H.FORMATS[f][2]
        # invalid blocks decode to all-zero tiles; count via the per-block API
        fn = getattr(ref, "detexDecompressBlock" + f)
        tmp = np.zeros(256, dtype=np.uint8)
        bad = sum(0 if fn(b.ctypes.data + i * H.FORMATS[f][1], 0xFFFFFFFF, 0, tmp.ctypes.data) else 1
                  for i in range(n))
        gold["uniform"][f] = {"seed": 1, "blocks": n, "size": [256, 256], "ok": bool(ok), "invalid": bad,
                              "linear": "%016x" % H.fnv1a64(lin), "tiled": "%016x" % H.fnv1a64(til),
                              "tile_bytes": tile_bytes}
        gold["clipped"][f] = {}
        for (w, h) in CLIPPED_SIZES:
            wib, hib = (w + 3) // 4, (h + 3) // 4
            bb = H.synth_blocks(f, wib * hib, seed=5)
            o, okc, _ = H.cpu_decode(ref, f, bb, w, h, linear=True)
            gold["clipped"][f]["%dx%d" % (w, h)] = {"seed": 5, "ok": bool(okc), "linear": "%016x" % H.fnv1a64(o)}
    for f, nm in (("BPTC", 8), ("BPTC_FLOAT", 14), ("BPTC_SIGNED_FLOAT", 14)):
        gold["forced_modes"][f] = {}
        for m in range(nm):
            b = H.force_modes(f, H.synth_blocks(f, n, seed=3), [m])
            lin, ok, _ = H.cpu_decode(ref, f, b, 256, 256, linear=True)
            gold["forced_modes"][f][str(m)] = {"seed": 3, "ok": bool(ok), "linear": "%016x" % H.fnv1a64(lin)}
        b = H.force_modes(f, H.synth_blocks(f, n, seed=3))
        lin, ok, _ = H.cpu_decode(ref, f, b, 256, 256, linear=True)
        gold["forced_modes"][f]["cycle"] = {"seed": 3, "ok": bool(ok), "linear": "%016x" % H.fnv1a64(lin)}
    with open(os.path.join(OUT, "synthetic_golden.json"), "w") as fp:
        json.dump(gold, fp, indent=1, sort_keys=True)
    print("wrote fixtures.npz and synthetic_golden.json")


if __name__ == "__main__":
    main()
