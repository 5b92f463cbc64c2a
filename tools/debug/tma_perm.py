"""debug aid: where do the 16-byte row segments written by the TMA-store kernel come from?"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import detex_b200 as D
from oracle import harness as H

cpu = H.best_cpu_decoder()[0]
for (w, h) in ((1024, 1024), (8192, 512), (4096, 64)):
    fmt = "BPTC"
    n = (w // 4) * (h // 4)
    blocks = H.force_modes(fmt, H.synth_blocks(fmt, n, seed=1))
    want, _, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=True, threads=8)
    d_blocks = torch.from_numpy(blocks.copy()).cuda()
    d_out = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda")
    D.decode_device(fmt, d_blocks.data_ptr(), w, h, d_out.data_ptr(), layout=D.LAYOUT_LINEAR)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    g16 = got.reshape(-1, 16)
    w16 = want.reshape(-1, 16)
    bad = np.nonzero((g16 != w16).any(axis=1))[0]
    print("%dx%d: %d of %d row segments differ" % (w, h, bad.size, g16.shape[0]))
    if bad.size == 0:
        continue
    # index: segment bytes -> position in want
    index = {}
    for i in range(w16.shape[0]):
        index.setdefault(w16[i].tobytes(), i)
    wib = w // 4
    def coords(s):       # segment index -> (pixel row y, block column bx)
        return divmod(s, wib)
    for s in list(bad[:24]) + list(bad[bad.size // 2: bad.size // 2 + 8]):
        src = index.get(g16[s].tobytes())
        y, bx = coords(s)
        if src is None:
            print("  seg (y=%d, bx=%d): content not found in the expected image (zero=%s)" % (y, bx, not g16[s].any()))
        else:
            sy, sbx = coords(src)
            print("  seg (y=%d, bx=%d) holds expected (y=%d, bx=%d)   [block id %d row %d <- block id %d row %d]" %
                  (y, bx, sy, sbx, (y // 4) * wib + bx, y % 4, (sy // 4) * wib + sbx, sy % 4))
