"""Summarise an `ncu --page source --csv` dump: executed warp-instructions and stall samples by SASS opcode."""
import csv
import collections
import sys


def main(path, top=18):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci, cs, cn = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cstall = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    by_op_inst, by_op_samp = collections.Counter(), collections.Counter()
    stall_tot = collections.Counter()
    tot_i = tot_s = 0
    for r in rows[hi + 1:]:
        if len(r) <= cn:
            continue
        toks = r[ci].strip().split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.split(".")[0]
        try:
            ni, ns = int(r[cn]), int(r[cs])
        except ValueError:
            continue
        by_op_inst[op] += ni
        by_op_samp[op] += ns
        tot_i += ni
        tot_s += ns
        for c in cstall:
            try:
                stall_tot[hdr[c]] += int(r[c])
            except ValueError:
                pass
    print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
    print("opcode            inst%   samples%")
    for op, n in by_op_inst.most_common(top):
        print(f"  {op:14s} {100*n/tot_i:6.2f}  {100*by_op_samp[op]/max(tot_s,1):6.2f}")
    if stall_tot:
        print("stall reasons:", ", ".join(f"{k[6:]}={100*v/max(sum(stall_tot.values()),1):.1f}%" for k, v in stall_tot.most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1])
