#!/usr/bin/env python
"""Pageable-buffer e2e of detexDecompressTextureLinear under the library's tuning knobs (one process per setting: the
library reads DETEXB200_HOST_THREADS / DETEXB200_POOL_SLAB_KB / DETEXB200_POOL_COPY once).  Prints one line."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import detex_b200 as D  # noqa: E402
from oracle import harness as H  # noqa: E402


def host_copy_bandwidth(threads, mb=64):
    """GB/s of numpy copies of `mb` MiB per thread, `threads` at once (what the host memory system gives memcpy)"""
    src = [np.ones(mb << 20, dtype=np.uint8) for _ in range(threads)]
    dst = [np.empty(mb << 20, dtype=np.uint8) for _ in range(threads)]

    def work(i):
        for _ in range(4):
            np.copyto(dst[i], src[i])
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return threads * 4 * (mb << 20) / (time.perf_counter() - t0) / 1e9


def main():
    fmt, w, h = "BC1", 8192, 8192
    tf, bs, ps = H.FORMATS[fmt]
    n = (w // 4) * (h // 4)
    blocks = H.synth_blocks(fmt, n, seed=1)
    out = np.zeros(w * h * ps, dtype=np.uint8)
    lib = D.load()
    tex = D.DetexTexture(tf, blocks.ctypes.data, w, h, w // 4, h // 4)
    for _ in range(3):
        lib.detexDecompressTextureLinear(C.byref(tex), out.ctypes.data, tf & 0xFFFF)
    steps = 8
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.detexDecompressTextureLinear(C.byref(tex), out.ctypes.data, tf & 0xFFFF)
    dt = (time.perf_counter() - t0) / steps
    pin_in = torch.from_numpy(blocks.copy()).pin_memory()
    pin_out = torch.zeros(out.size, dtype=torch.uint8).pin_memory()
    tex2 = D.DetexTexture(tf, pin_in.data_ptr(), w, h, w // 4, h // 4)
    for _ in range(2):
        lib.detexDecompressTextureLinear(C.byref(tex2), pin_out.data_ptr(), tf & 0xFFFF)
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.detexDecompressTextureLinear(C.byref(tex2), pin_out.data_ptr(), tf & 0xFFFF)
    dp = (time.perf_counter() - t0) / steps
    extra = ""
    if "--register" in sys.argv:        # what would pinning the caller's buffer in place cost?
        rt = torch.cuda.cudart()
        for mb in (4, 64, 256):
            a = np.zeros(mb << 20, dtype=np.uint8)
            t0 = time.perf_counter()
            rc = rt.cudaHostRegister(a.ctypes.data, a.size, 0)
            t1 = time.perf_counter()
            rt.cudaHostUnregister(a.ctypes.data)
            t2 = time.perf_counter()
            extra += "  cudaHostRegister %d MiB: %.2f ms (%.1f GB/s), unregister %.2f ms (rc %s);" % (mb, (t1 - t0) * 1e3, a.size / (t1 - t0) / 1e9, (t2 - t1) * 1e3, rc)
    if "--host-bw" in sys.argv:
        extra = "  host copy GB/s by threads: " + ", ".join("%d: %.1f" % (t, host_copy_bandwidth(t)) for t in (1, 2, 4, 8, 16))
    print("threads=%s ring=%s slab_kb=%s copy=%s: pageable %.0f M blocks/s (%.2f ms, %.1f GB/s out)  pinned %.0f M  ratio %.2f%s" %
          (os.environ.get("DETEXB200_HOST_THREADS", "def"), os.environ.get("DETEXB200_POOL_RING", "def"), os.environ.get("DETEXB200_POOL_SLAB_KB", "def"),
           os.environ.get("DETEXB200_POOL_COPY", "nt"), n / dt / 1e6, dt * 1e3, out.size / dt / 1e9, n / dp / 1e6, dp / dt, extra), flush=True)


if __name__ == "__main__":
    main()
