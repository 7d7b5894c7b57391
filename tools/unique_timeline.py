#!/usr/bin/env python
"""Per-block phase time stamps of the persistent unique kernel (b200ps_debug_buffer): where the 38 x 32768-id
dedup of one DeepFM batch spends its time.  Prints, per stamp, min / median / max over blocks in us from the
first block's start."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elasticdl_b200 import _lib  # noqa: E402
from elasticdl_b200.workloads.deepfm import GROUP_ROWS, synthetic_batch  # noqa: E402


def main():
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    G, B = len(GROUP_ROWS), int(os.environ.get("B", 32768))
    bounds = (ctypes.c_int64 * G)(*GROUP_ROWS)
    ws = torch.zeros(lib.b200ps_unique_bounded_workspace(G, B, bounds), dtype=torch.uint8, device=dev)
    uniq = torch.empty(G * B, dtype=torch.int64, device=dev)
    inv = torch.empty(G * B, dtype=torch.int32, device=dev)
    n = torch.empty(G, dtype=torch.int32, device=dev)
    dbg = torch.zeros(1024 * 8, dtype=torch.int64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for dist in ("zipf", "uniform"):
        batches = [synthetic_batch(B, 10 + i, dev, dist)[0].to(torch.int32).contiguous() for i in range(4)]
        for bps in (0, 1):
            for it in range(6):
                ids = batches[it % 4]
                if it == 5:
                    lib.b200ps_debug_buffer(dbg.data_ptr(), dbg.numel() * 8)
                    dbg.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.b200ps_unique_bounded_ex(None, ids.data_ptr(), 1, G, B, bounds, uniq.data_ptr(), inv.data_ptr(),
                                                  n.data_ptr(), ws.data_ptr(), ws.numel(), bps, st)
                e1.record()
                assert rc == 0
                torch.cuda.synchronize()
            lib.b200ps_debug_buffer(None, 0)
            t = dbg.cpu().numpy().reshape(-1, 8)
            t = t[t[:, 0] > 0]
            t0 = t[:, 0].min()
            names = ["start", "A done", "barrier1 passed", "B done", "barrier2 passed", "C done"]
            out = {"dist": dist, "blocks_per_sm": bps, "blocks": int(len(t)), "event_us": e0.elapsed_time(e1) * 1e3,
                   "unique_total": int(n.sum().item())}
            for j, nm in enumerate(names):
                col = (t[:, j] - t0) / 1e3
                out[nm] = [round(float(col.min()), 1), round(float(np.median(col)), 1), round(float(col.max()), 1)]
            # which blocks are the slow ones: blocks are laid out segment by segment
            dur = (t[:, 1] - t[:, 0]) / 1e3
            order = np.argsort(-dur)[:5]
            out["slowest_A_blocks"] = [[int(i), round(float(dur[i]), 1)] for i in order]
            durb = (t[:, 3] - t[:, 2]) / 1e3
            out["slowest_B_blocks"] = [[int(i), round(float(durb[i]), 1)] for i in np.argsort(-durb)[:5]]
            print(json.dumps(out))


if __name__ == "__main__":
    main()
