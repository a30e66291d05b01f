"""Static SASS statistics per kernel of a built library (no GPU needed).

    python tools/sass_count.py hyperbolic_optics_b200/libho_b200.so [filter]
"""
import collections
import re
import subprocess
import sys


def main(lib, flt=""):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, cnt, ops = None, collections.Counter(), collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r".*(reflect_kernel|tsweep_kernel|fp64_peak_kernel)", r"\1", m.group(1))[:40]
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            cnt[cur] += 1
            ops[cur][m.group(2).split(".")[0]] += 1
    for k, v in sorted(cnt.items()):
        if flt in k:
            o = ops[k]
            print(f"{k:42s} {v:6d} instr  fp64 {sum(o[x] for x in ('DFMA', 'DMUL', 'DADD', 'DSETP')):5d}  "
                  f"LDL {o['LDL']:4d} STL {o['STL']:4d}  CALL {o['CALL']:3d} BRA {o['BRA']:4d} MUFU {o['MUFU']:3d}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
