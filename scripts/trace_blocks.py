#!/usr/bin/env python
"""Reads a C5_TRACE_FILE (ctx launch block sm t_start_ns t_end_ns per line) and prints, per launch:
duration, blocks that did real work, busy-slot occupancy over time, per-SM spread."""
import sys
import numpy as np

def main(path):
    rows = []
    for line in open(path):
        p = line.split()
        if len(p) == 6:
            rows.append((int(p[0], 16), int(p[1]), int(p[2]), int(p[3]), int(p[4]), int(p[5])))
    a = np.array(rows, dtype=np.int64)
    t_origin = a[:, 4][a[:, 4] > 0].min()
    for ctx in np.unique(a[:, 0]):
        for launch in np.unique(a[a[:, 0] == ctx][:, 1]):
            r = a[(a[:, 0] == ctx) & (a[:, 1] == launch)]
            r = r[r[:, 4] > 0]
            dur = (r[:, 5] - r[:, 4]) / 1e3
            work = dur > 20.0                       # blocks that walked rays (us)
            t0, t1 = r[:, 4].min(), r[:, 5].max()
            w = r[work]
            line = f"ctx {ctx & 0xffff:04x} launch {launch:2d}: start {1e-3 * (t0 - t_origin):9.1f} us, length {1e-3 * (t1 - t0):7.1f} us, blocks {len(r)}, working {int(work.sum())}"
            if len(w):
                per_sm = np.bincount(w[:, 3], minlength=148)
                busy = np.zeros(148)
                for sm in range(148):
                    q = w[w[:, 3] == sm]
                    if len(q):
                        busy[sm] = (q[:, 5].max() - q[:, 4].min()) / 1e3
                line += (f", block us p50 {np.median(dur[work]):6.1f} p90 {np.percentile(dur[work], 90):6.1f} max {dur[work].max():6.1f}"
                         f", working blocks per SM min {per_sm.min()} max {per_sm.max()}, SM span us min {busy[busy > 0].min():6.1f} max {busy.max():6.1f}"
                         f", slot-time {dur[work].sum() / (1e-3 * (t1 - t0)) / 148:5.2f} blocks/SM avg")
            print(line)

if __name__ == "__main__":
    main(sys.argv[1])
