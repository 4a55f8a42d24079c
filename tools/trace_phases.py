#!/usr/bin/env python
"""Phase timeline of the fused cycle kernel from an OSC_TRACE build (see osc_kindyn.cuh: trace_point).

  make -C sai_primitives_b200/csrc BUILD=../_build_trace OUT=../libsai_b200_osc_trace.so EXTRA=-DOSC_TRACE
  SAI_B200_OSC_LIB=.../libsai_b200_osc_trace.so OSC_TRACE_FILE=gpurun_out/trace.bin python bench.py --steps 4 --warmup 3 --no-cpu
  python tools/trace_phases.py gpurun_out/trace.bin
"""
import sys
import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.uint64)
MAGIC = 0x4F53435452414345
launches = []
pos = 0
while pos < len(raw):
    assert raw[pos] == MAGIC, "bad trace file"
    grid = int(raw[pos + 1])
    n = grid * 64
    launches.append(raw[pos + 2: pos + 2 + n].reshape(grid, 32, 2))
    pos += 2 + n
print("launches in file:", len(launches))
which = int(sys.argv[2]) if len(sys.argv) > 2 else len(launches) // 2
t = launches[which]
grid = t.shape[0]
npts = min(31, int(t[0, 31, 1]))
smid = t[:, 31, 0].astype(int)
clk = t[:, :npts, 0].astype(np.int64)
gt = t[:, :npts, 1].astype(np.int64)
g0 = gt[:, 0].min()
start = gt[:, 0] - g0
end = gt[:, npts - 1] - g0
print("launch %d: grid %d, %d trace points/block, kernel span %.1f us" % (which, grid, npts, end.max() / 1e3))
order = np.argsort(start)
first_wave = start < np.percentile(start, 55)
print("block start times (us): min %.2f p25 %.2f p50 %.2f p75 %.2f max %.2f" % tuple(np.percentile(start, [0, 25, 50, 75, 100]) / 1e3))
print("block end   times (us): min %.2f p25 %.2f p50 %.2f p75 %.2f max %.2f" % tuple(np.percentile(end, [0, 25, 50, 75, 100]) / 1e3))
dur = clk[:, 1:] - clk[:, :-1]
print("\nper-phase SM cycles (median over blocks | first-wave median | later blocks median) and skew of phase-entry time")
for k in range(npts - 1):
    a = dur[:, k]
    sk = gt[first_wave, k] - g0
    print("  phase %2d: %7.0f | %7.0f | %7.0f   entry-time spread of first-wave blocks p5..p95: %.2f..%.2f us"
          % (k, np.median(a), np.median(a[first_wave]), np.median(a[~first_wave]) if (~first_wave).any() else 0,
             np.percentile(sk, 5) / 1e3, np.percentile(sk, 95) / 1e3))
tot = clk[:, npts - 1] - clk[:, 0]
print("block total cycles: median %.0f  first wave %.0f  later %.0f" % (np.median(tot), np.median(tot[first_wave]), np.median(tot[~first_wave]) if (~first_wave).any() else 0))
print("blocks per SM: min %d max %d; SMs used %d" % (np.bincount(smid).min(), np.bincount(smid).max(), len(np.unique(smid))))
