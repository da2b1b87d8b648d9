#!/usr/bin/env python
"""Reads a WT_TRACE dump (tools/explore_sched.py) and reports how busy the SMs were over the launch:
per-SM item counts, the distribution of item durations and the number of resident warps over time."""
import sys
import numpy as np


def report(path):
    t = np.load(path).astype(np.int64)
    smid, cta, t0, t1 = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    ok = t1 > 0
    smid, cta, t0, t1 = smid[ok], cta[ok], t0[ok], t1[ok]
    base = t0.min()
    t0, t1 = t0 - base, t1 - base
    total = t1.max()
    dur = t1 - t0
    print(f"{path}: {len(dur)} items on {len(np.unique(smid))} SMs, launch span {total/1e3:.1f} us")
    print(f"  item duration us: min {dur.min()/1e3:.1f} p10 {np.percentile(dur,10)/1e3:.1f} median {np.median(dur)/1e3:.1f} p90 {np.percentile(dur,90)/1e3:.1f} max {dur.max()/1e3:.1f}")
    per_sm = np.bincount(smid)
    per_sm = per_sm[per_sm > 0]
    print(f"  items per SM: min {per_sm.min()} max {per_sm.max()} mean {per_sm.mean():.2f}")
    print(f"  start us: p50 {np.median(t0)/1e3:.1f} p99 {np.percentile(t0,99)/1e3:.1f} max {t0.max()/1e3:.1f};  end us: p1 {np.percentile(t1,1)/1e3:.1f} p10 {np.percentile(t1,10)/1e3:.1f} p50 {np.median(t1)/1e3:.1f} p90 {np.percentile(t1,90)/1e3:.1f}")
    # resident work items over time (all SMs), 20 bins
    edges = np.linspace(0, total, 21)
    ev = np.concatenate([np.stack([t0, np.ones_like(t0)], 1), np.stack([t1, -np.ones_like(t1)], 1)])
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    lvl = np.cumsum(ev[:, 1])
    avg = []
    for a, b in zip(edges[:-1], edges[1:]):
        m = (ev[:, 0] >= a) & (ev[:, 0] < b)
        avg.append(lvl[m].mean() if m.any() else (avg[-1] if avg else 0))
    print("  resident items over 20 equal time bins:", " ".join(f"{x:.0f}" for x in avg))
    busy = dur.sum() / (total * max(1, lvl.max()))
    print(f"  sum(item time) / (span x peak residency) = {busy:.3f}")
    # last finisher per SM
    last = np.zeros(smid.max() + 1)
    np.maximum.at(last, smid, t1)
    last = last[last > 0]
    print(f"  per-SM last finish us: min {last.min()/1e3:.1f} median {np.median(last)/1e3:.1f} max {last.max()/1e3:.1f}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        report(p)
