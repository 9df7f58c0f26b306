"""Regenerates tests/golden/known_answers.json from the fixtures with madler zlib (Python's zlib module).

The numbers pin the CPU oracle (oracle/compu_oracle.c) to SURVEY.md §8(c) "Known answers to carry forward".
"""
import json, os, zlib

HERE = os.path.dirname(os.path.abspath(__file__))


def sizes(data, level):
    out = {}
    for name, wbits in (("raw", -15), ("zlib", 15), ("gzip", 31)):
        c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, zlib.Z_DEFAULT_STRATEGY)
        out[name] = len(c.compress(data) + c.flush())
    return out


def main():
    res = {"zlib_version": zlib.ZLIB_RUNTIME_VERSION, "files": {}}
    for fn in ("10x10y", "alice29.txt"):
        data = open(os.path.join(HERE, fn), "rb").read()
        res["files"][fn] = {
            "len": len(data),
            "crc32": "%08x" % zlib.crc32(data),
            "adler32": "%08x" % zlib.adler32(data),
            "levels": {str(l): sizes(data, l) for l in (1, 6, 9)},
        }
    c = zlib.compressobj(9, zlib.DEFLATED, 31, 8, 0)
    res["gzip_l9_10x10y_hex"] = (c.compress(open(os.path.join(HERE, "10x10y"), "rb").read()) + c.flush()).hex()
    json.dump(res, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
