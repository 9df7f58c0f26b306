#!/usr/bin/env python
"""Per-source-line roll-up straight from an ncu report captured with --import-source on (development tool):
`ncu -i REP --page source --csv --print-source cuda,sass` already correlates SASS with the CUDA lines embedded in the report,
so this works for old reports whose sources have changed since. usage: ncu_src.py REP [kernel_substring] [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kname = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    agg = {}
    cur_file, cur_fn, hdr, cur_line = None, None, None, None
    tot_i = tot_s = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            cur_fn = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            ci = {n: [i for i, h in enumerate(hdr) if h == n] for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed")}
            continue
        if hdr is None or (kname and kname not in (cur_fn or "")):
            continue
        if r[0]:  # a source line row
            cur_line = (cur_file, int(r[0]), r[1].strip()[:100])
            continue
        if len(r) < len(hdr) or r[2] in ("...", "-"):
            continue
        try:
            sm = int(r[ci["# Samples"][0]]); ie = int(r[ci["Instructions Executed"][0]]); te = int(r[ci["Thread Instructions Executed"][0]])
        except ValueError:
            continue
        a = agg.setdefault(cur_line, [0, 0, 0, 0])
        a[0] += sm; a[1] += ie; a[2] += te; a[3] += 1
        tot_i += ie; tot_s += sm
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for key, (sm, ie, te, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%5.1f%% inst %5.1f%% samp  thr/inst %4.1f  sass %3d  %s:%d  %s" % (100.0 * ie / max(1, tot_i), 100.0 * sm / max(1, tot_s), te / max(1, ie), n, key[0], key[1], key[2]))


if __name__ == "__main__":
    main()
