#!/usr/bin/env python
"""Copies what tools/capture_profiles.sh brought back (gpurun_out/<prefix>_*) into profiles/ under the round-2 names and rewrites
profiles/traffic.json (DRAM bytes per launch from the `--set full` summaries, keyed by the hash of the kernel sources).
usage: install_profiles.py <prefix>   (development tool; run in the repo root after the capture)"""
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = {
    "inflate_ncu.md": "r2_inflate_cfg2_ncu.md", "inflate_tok_lines.txt": "r2_inflate_tok_lines.txt",
    "inflate_lz_lines.txt": "r2_inflate_lz_lines.txt", "deflate_ncu.md": "r2_deflate_chain_1gib_ncu.md",
    "deflate_match_lines.txt": "r2_deflate_match_lines.txt", "runs_ncu.md": "r2_runs_gzip_cfg4_ncu.md",
    "runs_candidates_lines.txt": "r2_runs_candidates_lines.txt", "launches.csv": "r2_bench_all_launches.csv",
    "bench.json": "r2_bench_line_n1.json",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def dram_bytes(md, launch_ids):
    total = 0.0
    for sec in open(md).read().split("\n## ")[1:]:
        lid = int(re.search(r"launch id (\d+)", sec).group(1))
        if lid not in launch_ids:
            continue
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(re.escape(key) + r" \| ([0-9.]+) \| (\w+)", sec)
            total += float(m.group(1)) * SCALE[m.group(2)]
    return total


def main():
    prefix = sys.argv[1]
    for src, dst in NAMES.items():
        shutil.copyfile(os.path.join(ROOT, "gpurun_out", "%s_%s" % (prefix, src)), os.path.join(ROOT, "profiles", dst))
    line = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_line_n1.json")))
    h = bench.kernel_source_hash()
    assert line["config"]["kernel_src_hash"] == h, "the capture was taken on other kernel sources (%s, here %s)" % (line["config"]["kernel_src_hash"], h)
    inf = dram_bytes(os.path.join(ROOT, "profiles", "r2_inflate_cfg2_ncu.md"), {0, 1})
    dfl = dram_bytes(os.path.join(ROOT, "profiles", "r2_deflate_chain_1gib_ncu.md"), set(range(11)))
    t = {
        "src_hash": h,
        "inflate_cfg2_dram_bytes_per_launch": int(inf),
        "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum at 65 536 x 64 KiB streams on a B200 (profiles/"
                  "r2_inflate_cfg2_ncu.md, captured by tools/capture_profiles.sh on the kernel sources with this src_hash): "
                  "inflate_tok_kernel + inflate_lz_kernel = %.2f GB; algorithmic bytes 6.32 GB" % (inf / 1e9),
        "deflate_cfg3_dram_bytes_per_launch": int(dfl * 4),
        "deflate_cfg3_source": "sum over the 11 kernels of one chain (checksum, chains, match, parse, histogram, plan, layouts, emit) of "
                               "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full on 1 GiB (profiles/r2_deflate_chain_1gib_ncu.md: "
                               "%.1f GB per GiB), times 4 for the 4 GiB of cfg3; algorithmic bytes 6.26 GB" % (dfl / 1e9),
    }
    json.dump(t, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=2)
    print("installed %s: src_hash %s, inflate %.2f GB per launch, deflate chain %.1f GB per GiB" % (prefix, h, inf / 1e9, dfl / 1e9))


if __name__ == "__main__":
    main()
