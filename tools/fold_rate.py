#!/usr/bin/env python
"""How many grad_value reductions could be folded before they leave the SM?  (CPU, numpy; no GPU needed.)

The backward kernel issues one 128-byte vector red per (point, bilinear corner) whose weight is non-zero
(ir_ads_b200/csrc/msda_fast.cuh).  Two contributions can be pre-added on the SM only if they go to the SAME
(image, pixel, head) row and are held by the same warp / CTA at the same time.  This script counts, for the
bench's synthetic distributions, the fraction of reds that survive folding at three scopes:

  warp-step   the 4 rows (x-adjacent queries of one head, STRIP order) x 4 corners a warp holds for ONE point
              index -- what a __match_any_sync fold can see
  warp-row    the same 4 rows over ALL L*P points of the rows (needs the warp to keep 64 x 4 partial rows)
  strip       one 128-thread STRIP CTA: 16 consecutive queries of one head, all points
  tile WxH    an (image, head, W x H query tile of one level) CTA, all points, per SAMPLED level -- what an
              on-SM accumulation window (shared memory or a sorted segment sum) could fold

Output: surviving reds / issued reds (1.0 = nothing folds).  Usage: python tools/fold_rate.py [cfg2|cfg5|...]
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ir_ads_b200.workloads import WORKLOADS, make_workload_inputs  # noqa: E402


def corner_ids(loc, levels):
    """loc [Q,H,L,P,2] -> dest [Q,H,L,P,4] int64 pixel ids inside the image (-1 = padded / gated corner)."""
    Q, H, L, P, _ = loc.shape
    out = np.full((Q, H, L, P, 4), -1, dtype=np.int64)
    start = 0
    for l, (hl, wl) in enumerate(levels):
        x = loc[:, :, l, :, 0].astype(np.float64) * wl - 0.5
        y = loc[:, :, l, :, 1].astype(np.float64) * hl - 0.5
        ok = (x > -1) & (y > -1) & (x < wl) & (y < hl)
        x0 = np.floor(x).astype(np.int64)
        y0 = np.floor(y).astype(np.int64)
        for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            xx, yy = x0 + dx, y0 + dy
            v = ok & (xx >= 0) & (xx < wl) & (yy >= 0) & (yy < hl)
            out[:, :, l, :, k] = np.where(v, start + yy * wl + xx, -1)
        start += hl * wl
    return out


def surviving(ids_2d):
    """ids_2d [groups, n]: per group, number of distinct non-negative ids; returns (distinct, issued)."""
    s = np.sort(ids_2d, axis=1)
    valid = s >= 0
    new = np.ones_like(valid)
    new[:, 1:] = s[:, 1:] != s[:, :-1]
    return int((valid & new).sum()), int(valid.sum())


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    wl = WORKLOADS[name]
    res = {"workload": name, "levels": [list(x) for x in wl.levels], "points": wl.num_points}
    for dist in ("model", "test"):
        _, _, _, loc, _ = make_workload_inputs(wl, dist=dist, seed=0, batch=1)
        loc = loc[0].numpy()                                  # [Q,H,L,P,2]
        Q, H, L, P, _ = loc.shape
        ids = corner_ids(loc, wl.levels)                      # [Q,H,L,P,4]
        r = {}
        # warp-step: 4 consecutive queries, one head, one (l,p)
        q4 = (Q // 4) * 4
        g = ids[:q4].reshape(Q // 4, 4, H, L, P, 4).transpose(0, 2, 3, 4, 1, 5).reshape(-1, 16)
        d, n = surviving(g)
        r["warp_step_4rows_x_4corners"] = round(d / n, 4)
        # warp-row: 4 consecutive queries, one head, per level all P points (levels never share pixels)
        g = ids[:q4].reshape(Q // 4, 4, H, L, P, 4).transpose(0, 2, 3, 1, 4, 5).reshape(-1, 4 * P * 4)
        d, n = surviving(g)
        r["warp_row_4rows_all_points"] = round(d / n, 4)
        for n_rows in (16, 32, 64):
            qn = (Q // n_rows) * n_rows
            g = ids[:qn].reshape(Q // n_rows, n_rows, H, L, P, 4).transpose(0, 2, 3, 1, 4, 5).reshape(-1, n_rows * P * 4)
            d, n = surviving(g)
            r[f"strip_{n_rows}rows_all_points"] = round(d / n, 4)
        q16 = (Q // 16) * 16
        per_level = []
        for l in range(L):
            gl = ids[:q16].reshape(Q // 16, 16, H, L, P, 4)[:, :, :, l].transpose(0, 2, 1, 3, 4).reshape(-1, 16 * P * 4)
            d, n = surviving(gl)
            per_level.append(round(d / n, 4))
        r["strip_16rows_per_sampled_level"] = per_level
        # 2-D query tiles (encoder form only: query i is pixel i)
        if wl.num_query == 0:
            for tw, th in ((8, 4), (8, 8), (16, 8), (16, 16), (32, 16)):
                tot_d = np.zeros(L, dtype=np.int64)
                tot_n = np.zeros(L, dtype=np.int64)
                start = 0
                for (hl, wl_) in wl.levels:
                    blk = ids[start:start + hl * wl_].reshape(hl, wl_, H, L, P, 4)
                    for ty in range(0, hl, th):
                        for tx in range(0, wl_, tw):
                            t = blk[ty:ty + th, tx:tx + tw]                  # [th', tw', H, L, P, 4]
                            t = t.reshape(-1, H, L, P * 4).transpose(1, 2, 0, 3).reshape(H * L, -1)
                            s = np.sort(t, axis=1)
                            valid = s >= 0
                            new = np.ones_like(valid)
                            new[:, 1:] = s[:, 1:] != s[:, :-1]
                            dd = (valid & new).sum(1).reshape(H, L).sum(0)
                            nn = valid.sum(1).reshape(H, L).sum(0)
                            tot_d += dd
                            tot_n += nn
                    start += hl * wl_
                r[f"tile_{tw}x{th}_per_sampled_level"] = [round(float(a) / float(b), 4) for a, b in zip(tot_d, tot_n)]
                r[f"tile_{tw}x{th}_all_levels"] = round(float(tot_d.sum()) / float(tot_n.sum()), 4)
        res[dist] = r
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
