"""Reads the per-CTA timeline of the packed-half scan (SS_DTW_H2_TIMELINE=<file> python tools/dtw_sweep.py ...) and prints
where a step's SM time goes: set-up (barrier init, TMEM allocation), the DP loop, the CTA's tail (list merge, store, TMEM
release), the gaps between consecutive CTAs of an SM, and the idle time at the end of the step.
  python tools/h2_timeline.py gpurun_out/tl_100000.txt"""
import sys
from collections import defaultdict

import numpy as np

rows = np.loadtxt(sys.argv[1], dtype=np.int64, comments="#", ndmin=2)
rows = rows[rows[:, 1] > 0]
cta, t0, t1, t2, t3, smid, kind, group, L, ntiles = rows[:, :10].T
T0, T1 = t0.min(), t3.max()
span = (T1 - T0) / 1e6
nsm = len(np.unique(smid))
print("%d CTAs on %d SMs, step span %.3f ms" % (len(rows), nsm, span))
busy = (t3 - t0).sum() / 1e6
print("SM-time: span x SMs = %.1f ms; inside CTAs %.1f ms (%.1f %%)" % (span * nsm, busy, 100 * busy / (span * nsm)))
for name, a, b in (("set-up", t0, t1), ("DP loop", t1, t2), ("tail", t2, t3)):
    d = (b - a) / 1e3
    print("  %-8s total %.2f ms of SM-time (%.2f %%), per CTA median %.1f us, p90 %.1f us, max %.1f us"
          % (name, d.sum() / 1e3, 100 * d.sum() / 1e3 / (span * nsm), np.median(d), np.percentile(d, 90), d.max()))
gaps, idle_end, idle_start = [], [], []
per_sm = defaultdict(list)
for i in range(len(rows)):
    per_sm[smid[i]].append((t0[i], t3[i]))
for s, lst in per_sm.items():
    lst.sort()
    idle_start.append(lst[0][0] - T0)
    idle_end.append(T1 - lst[-1][1])
    gaps += [lst[j + 1][0] - lst[j][1] for j in range(len(lst) - 1)]
gaps = np.array(gaps) / 1e3
print("  gaps between consecutive CTAs of an SM: total %.2f ms of SM-time (%.2f %%), median %.1f us, p90 %.1f us, max %.1f us"
      % (gaps.sum() / 1e3, 100 * gaps.sum() / 1e3 / (span * nsm), np.median(gaps), np.percentile(gaps, 90), gaps.max()))
ie, is_ = np.array(idle_end) / 1e3, np.array(idle_start) / 1e3
print("  idle before an SM's first CTA: total %.2f ms (%.2f %%); after its last: total %.2f ms (%.2f %%), mean %.1f us, max %.1f us"
      % (is_.sum() / 1e3, 100 * is_.sum() / 1e3 / (span * nsm), ie.sum() / 1e3, 100 * ie.sum() / 1e3 / (span * nsm), ie.mean(), ie.max()))
for k in sorted(np.unique(kind)):
    m = kind == k
    d = (t3 - t0)[m] / 1e3
    steps = ((L[m] + 1) // 2) * ntiles[m]
    print("kind NB=%d: %d CTAs, duration median %.1f us (min %.1f, max %.1f); first start %.3f ms, last end %.3f ms; %.3f us per (tile, 2-row step)"
          % (k, m.sum(), np.median(d), d.min(), d.max(), (t0[m].min() - T0) / 1e6, (t3[m].max() - T0) / 1e6, ((t2 - t1)[m].sum() / 1e3) / steps.sum()))

# the SM clock the DP phase ran at (clock64 cycles / globaltimer ns), and the two halves of every slice
if rows.shape[1] >= 12:
    cyc, th = rows[:, 10], rows[:, 11]
    f = cyc / np.maximum(t2 - t1, 1)
    print("SM clock over the DP phases (clock64 / globaltimer): median %.1f MHz, p10 %.1f, p90 %.1f" % (1e3 * np.median(f), 1e3 * np.percentile(f, 10), 1e3 * np.percentile(f, 90)))
    for k in sorted(np.unique(kind)):
        m = (kind == k) & (ntiles >= 4) & (th > 0)
        if not m.any():
            continue
        h1 = (ntiles[m] + 1) // 2
        a = (th - t1)[m] / 1e3 / h1
        b = (t2 - th)[m] / 1e3 / np.maximum(ntiles[m] - h1, 1)
        print("kind NB=%d: us per tile, first half of the slice %.3f, second half %.3f (ratio %.3f); SM cycles per tile %.0f"
              % (k, a.mean(), b.mean(), a.mean() / b.mean(), (cyc[m] / ntiles[m]).mean()))
