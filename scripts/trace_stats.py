#!/usr/bin/env python
"""Per-CTA timeline of a -DSPGG_TRACE build (debug): python scripts/trace_stats.py <trace file>..."""
import sys
import numpy as np
for f in sys.argv[1:]:
    t = np.fromfile(f, dtype=np.uint64).reshape(-1, 4)
    s = t[:, 0].astype(np.int64); e = t[:, 1].astype(np.int64); sm = t[:, 2]; tiles = t[:, 3]
    t0 = s.min(); dur = (e - s) / 1e3
    print(f"{f}: ctas {len(t)} first end {(e.min()-t0)/1e3:.1f} last end {(e.max()-t0)/1e3:.1f} us; dur min {dur.min():.1f} med {np.median(dur):.1f} max {dur.max():.1f}; tiles min {tiles.min()} max {tiles.max()}")
    half = len(t) // 2
    print(f"   first-wave CTAs dur med {np.median(dur[:half]):.1f}, second-wave {np.median(dur[half:]):.1f}")
    ends = {}
    for i in range(len(t)):
        ends[int(sm[i])] = max(ends.get(int(sm[i]), 0), (e[i] - t0) / 1e3)
    ev = np.array(list(ends.values()))
    print("   per-SM finish: min %.1f med %.1f max %.1f; #SM %d" % (ev.min(), np.median(ev), ev.max(), len(ev)))
