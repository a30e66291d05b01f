"""Per source line: instructions, stall samples by reason (wait / long_sb / math / short_sb) from an
`ncu --page source --csv --print-source cuda,sass` dump (first kernel only).

    python tools/ncu_stalls.py gpurun_out/<name>_srcsass.csv [top] [reason]
"""
import collections
import csv
import sys


def main(path, top=40, key="stall_wait"):
    hdr, ix, cur_file, launch = None, None, None, 0
    per = collections.defaultdict(lambda: collections.Counter())
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            launch += 1
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {}
            for i, h in enumerate(hdr):
                ix.setdefault(h, i)
            continue
        if hdr is None or launch > 1:
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        num = lambda name: int(r[ix[name]]) if r[ix[name]] not in ("", "-") else 0
        k = (cur_file, line, r[1].strip()[:80])
        c = per[k]
        c["inst"] += num("Instructions Executed")
        c["samples"] += num("# Samples")
        for name in ("stall_wait", "stall_long_sb", "stall_math", "stall_short_sb", "stall_no_inst", "stall_branch_resolving"):
            c[name] += num(name)
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print({k: v for k, v in tot.items()})
    for k, c in sorted(per.items(), key=lambda kv: -kv[1][key])[:top]:
        print(f"{100 * c[key] / max(tot[key], 1):5.1f}% {key[6:]:8s} {100 * c['inst'] / tot['inst']:5.1f}% inst  wait {c['stall_wait']:6d} long {c['stall_long_sb']:6d} "
              f"math {c['stall_math']:6d}  {k[0]}:{k[1]}  {k[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else "stall_wait")
