"""Opcode histogram (+ stall attribution) from `ncu --page source --csv` (SASS view)."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path))]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[hi + 1:]:
        if r and r[0] in ("Address", "Kernel Name"):
            break  # next launch section
        if len(r) == len(hdr):
            data.append(r)
    num = lambda x: int(x) if x not in ("", "-") else 0
    tot = sum(num(r[ix["# Samples"]]) for r in data)
    print("SASS instructions:", len(data), "total samples", tot)
    op_inst, op_samp = collections.Counter(), collections.Counter()
    stall_by_op = collections.defaultdict(collections.Counter)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in data:
        toks = r[ix["Source"]].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        op_inst[op] += num(r[ix["Instructions Executed"]])
        op_samp[op] += num(r[ix["# Samples"]])
        for c in stall_cols:
            stall_by_op[op][c] += num(r[ix[c]])
    ti = sum(op_inst.values())
    print("%-10s %8s %8s  top stalls (%% of that op's samples)" % ("op", "inst%", "samp%"))
    for op, n in op_inst.most_common(24):
        st = stall_by_op[op].most_common(3)
        print("%-10s %7.1f%% %7.1f%%  %s" % (op, 100 * n / ti, 100 * op_samp[op] / tot,
                                             [(k[6:], round(100 * v / max(1, op_samp[op]))) for k, v in st]))


if __name__ == "__main__":
    main(sys.argv[1])
