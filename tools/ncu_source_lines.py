#!/usr/bin/env python
"""Per-source-line share of executed instructions and stall samples of one kernel of an .ncu-rep (needs -lineinfo and
--import-source on).  usage: ncu_source_lines.py rep kernel-regex [top]"""
import csv, subprocess, sys, collections

def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    cur_file, hdr = None, None
    lines = collections.OrderedDict()
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r; i_inst = hdr.index("Instructions Executed"); i_smp = hdr.index("# Samples"); continue
        if hdr and r[0].isdigit() and len(r) > i_inst:
            try:
                key = (cur_file, int(r[0]))
                a = lines.setdefault(key, [0.0, 0.0, r[1].strip()[:90]])
                a[0] += float(r[i_inst] or 0); a[1] += float(r[i_smp] or 0)
            except ValueError:
                pass
    ti = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
    print("kernel %s: %.0f warp instructions, %.0f samples" % (kern, ti, ts))
    byfile = collections.defaultdict(lambda: [0.0, 0.0])
    for (f, l), v in lines.items():
        byfile[f][0] += v[0]; byfile[f][1] += v[1]
    for f, v in sorted(byfile.items(), key=lambda x: -x[1][1]):
        print("  %-26s instructions %5.1f %%  samples %5.1f %%" % (f, 100 * v[0] / max(ti, 1), 100 * v[1] / max(ts, 1)))
    print("  top lines by samples:")
    for (f, l), v in sorted(lines.items(), key=lambda x: -x[1][1])[:top]:
        print("  %5.1f %% smp %5.1f %% inst  %s:%d  %s" % (100 * v[1] / max(ts, 1), 100 * v[0] / max(ti, 1), f, l, v[2]))

if __name__ == "__main__":
    main()
