#!/usr/bin/env python
"""Per-CUDA-source-line warp-stall samples of one kernel from an .ncu-rep (needs -lineinfo and
--import-source on):  python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main(rep, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    cur_file, data = "", []
    for r in csv.reader(io.StringIO(raw)):
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) >= 5 and r[0].isdigit():
            try:
                data.append((int(r[4]), cur_file, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot = sum(d[0] for d in data) or 1
    print("total warp-stall samples: %d" % tot)
    for v, f, ln, src in sorted(data, reverse=True)[:top]:
        print("%7d %5.1f%%  %s:%d  %s" % (v, 100.0 * v / tot, f, ln, src[:100]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
