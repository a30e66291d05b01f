"""Aggregate an `ncu -i <rep> --page source --csv --print-source cuda,sass` dump by CUDA source line:
share of warp instructions, of FP64 instructions (DFMA/DMUL/DADD) and of stall samples per line.

    python tools/ncu_by_line.py gpurun_out/<name>_srcsass.csv [top]
"""
import collections
import csv
import sys


def main(path, top=45):
    cur_file, hdr, ix = None, None, None
    per_line = collections.defaultdict(lambda: [0, 0, 0])
    launch = 0
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            launch += 1
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            continue
        if hdr is None or launch > 1:   # first captured launch only
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        num = lambda x: int(x) if x not in ("", "-") else 0
        inst, samp = num(r[ix["Instructions Executed"]]), num(r[ix["# Samples"]])
        key = (cur_file, line, r[1].strip()[:90])
        per_line[key][0] += inst
        per_line[key][1] += samp
        toks = r[3].split()
        if toks:
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            if op.startswith(("DFMA", "DMUL", "DADD")):
                per_line[key][2] += inst
    tot = sum(v[0] for v in per_line.values())
    tots = sum(v[1] for v in per_line.values())
    totf = sum(v[2] for v in per_line.values())
    print(f"{path}: warp instructions {tot}, FP64 {totf} ({100 * totf / max(tot, 1):.1f} %), samples {tots}")
    for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / tot:5.1f}% inst {100 * v[2] / max(totf, 1):5.1f}% fp64 {100 * v[1] / max(tots, 1):5.1f}% samp  "
              f"{key[0]}:{key[1]}  {key[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
