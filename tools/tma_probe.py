"""Diagnostic: what a TMA box of the box mode leaves in shared memory (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xuance_b200 import _lib
dev = "cuda:0"


def probe(C, W, rows, box_c, box_px, box_h, step, coords, label, rows_shown=range(10)):
    # element value encodes (row, pixel, channel): row*4096 + px*64 + c  (exact in bf16? no -> use int16 view of bf16 bits)
    idx = (torch.arange(rows)[:, None, None] * 4096 + torch.arange(W)[None, :, None] * 64 + torch.arange(C)[None, None, :]).to(torch.int16)
    t = idx.view(torch.bfloat16).to(dev).contiguous().unsqueeze(0)          # [1, rows, W, C], bits = index
    nbytes = box_h * box_px * box_c * 2
    out = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)
    _lib.call("xb_debug_tma_box", _lib.ptr(t), t[0].numel(), 1, C, W, rows, box_c, box_px, box_h, step, *coords, nbytes,
              _lib.ptr(out), nbytes + 1024)
    torch.cuda.synchronize()
    raw = out.cpu().numpy()
    vals = raw[:nbytes].view(np.int16)
    print("==", label, "box", (box_c, box_px, box_h, step), "coords", coords, "bytes", nbytes)
    untouched = int((raw[:nbytes] == 0xEE).sum())
    print("   untouched bytes in box region:", untouched, " bytes touched after region:", int((raw[nbytes:] != 0xEE).sum()))
    # print the first 4 smem "rows" of 128 B as 16-byte chunks decoded (row, px, c of first element of the chunk)
    for r in rows_shown:
        chunks = []
        for ch in range(8):
            v = int(vals[(r * 128 + ch * 16) // 2]) & 0xffff
            chunks.append("%d/%d/%d" % (v // 4096, (v // 64) % 64, v % 64))
        print("   smem row %2d:" % r, " ".join(chunks))


probe(64, 10, 48, 64, 10, 12, 1, (0, 0, 2, 0), "conv3-type, in-bounds")
probe(64, 10, 48, 64, 10, 12, 1, (0, -1, 2, 0), "conv3-type, w0 = -1")
# the pixel-pair view of a 32-channel activation (22 pixels = 11 pairs of 64 channels), row step 2 from row 5: smem row
# 10*k + x must hold tensor row 5 + 2k, pair 1 + x
probe(64, 11, 96, 64, 10, 12, 2, (0, 1, 5, 0), "conv2-type (pixel pairs), step 2 from row 5", rows_shown=(0, 1, 9, 10, 11, 20, 30, 110, 119))
probe(64, 11, 96, 64, 10, 12, 2, (0, 0, -2, 0), "conv2-type (pixel pairs), step 2 from row -2", rows_shown=(0, 9, 10, 20, 30))
