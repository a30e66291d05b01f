"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line."""
import collections
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    cur_file, hdr = None, None
    per_line = collections.defaultdict(lambda: [0, 0])
    stall_tot = collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        if r[0] not in ("", "-"):
            d = dict(zip(hdr, r))
            try:
                line = int(r[0])
            except ValueError:
                continue
            num = lambda x: int(x) if x not in ("", "-", None) else 0
            inst = num(d["Instructions Executed"])
            samp = num(d["# Samples"])
            key = (cur_file, line, r[1].strip()[:100])
            per_line[key][0] += inst
            per_line[key][1] += samp
            for k, v in d.items():
                if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0", "-"):
                    stall_tot[k] += int(v)
    tot_inst = sum(v[0] for v in per_line.values())
    tot_s = sum(v[1] for v in per_line.values())
    print("total warp instructions", tot_inst, "samples", tot_s)
    s = sum(stall_tot.values())
    print("stall reasons (% of samples):", [(k, round(100 * v / s, 1)) for k, v in stall_tot.most_common(8)])
    for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / tot_inst:5.1f}% inst {100 * v[1] / tot_s:5.1f}% samp  {key[0]}:{key[1]}  {key[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
