#!/usr/bin/env python
"""Per-source-line roll-up of an ncu report (development tool).

ncu's CSV source page is per SASS instruction; this joins it with `nvdisasm -g` line info of the matching cubin so the
hot source lines (instructions executed, stall samples) can be read without the GUI.
usage: ncu_lines.py report.ncu-rep lib.so kernel_substring [top_n] [mangled_substring]
(kernel_substring selects the kernel section of the report; mangled_substring, default = kernel_substring, selects the
function in the cubin, e.g. inflate_tok_kernelILi14E for one template instantiation)
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def r_join(r):
    return ",".join(r)


def main():
    rep, so, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    mangled = sys.argv[5] if len(sys.argv) > 5 else kname
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    sass_lines = None
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if mangled in out:
            sass_lines = out.splitlines()
            break
    assert sass_lines, "kernel not found in any cubin"
    # walk the function: remember the current "//## File ..., line N" and assign to each instruction
    infn = False
    cur = None
    inst_lines = []
    for ln in sass_lines:
        if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
            infn = mangled in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            inst_lines.append(cur)
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csvtxt.splitlines()))
    # the report may hold several kernels: take the section whose "Kernel Name" row matches
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = next((i for i in starts if kname in r_join(rows[i])), starts[0] if starts else 0)
    nxt = next((i for i in starts if i > sec), len(rows))
    rows = rows[sec:nxt]
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {n: hdr.index(n) for n in ("Instructions Executed", "# Samples", "Source", "Thread Instructions Executed")}
    data = [r for r in rows[hdr_i + 1:] if len(r) > max(ci.values())]
    if len(data) != len(inst_lines):
        print("warning: %d SASS rows in report vs %d in cubin" % (len(data), len(inst_lines)))
    agg = {}
    tot_i = tot_s = 0
    for k, r in enumerate(data):
        key = inst_lines[k] if k < len(inst_lines) else None
        ie = int(r[ci["Instructions Executed"]] or 0)
        sm = int(r[ci["# Samples"]] or 0)
        te = int(r[ci["Thread Instructions Executed"]] or 0)
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += ie
        a[1] += sm
        a[2] += te
        tot_i += ie
        tot_s += sm
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    srcs = {}
    for key, (ie, sm, te) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = ""
        if key:
            fn, ln = key
            if fn not in srcs:
                p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "compu_b200", "csrc", fn)
                srcs[fn] = open(p).read().splitlines() if os.path.exists(p) else []
            if 0 < ln <= len(srcs[fn]):
                text = srcs[fn][ln - 1].strip()[:90]
        print("%5.1f%% samp %5.1f%% inst  thr/inst %4.1f  %s:%s  %s" % (100.0 * sm / max(1, tot_s), 100.0 * ie / max(1, tot_i),
                                                                    te / max(1, ie), key[0] if key else "?", key[1] if key else "?", text))


if __name__ == "__main__":
    main()
