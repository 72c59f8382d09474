"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time per kernel family."""
import collections
import csv
import re
import sys


DETAIL = False


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    kn, mv, mn, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        full = r[kn]
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", full))[:72]
        if DETAIL and ("elementwise" in name or name.strip().endswith("at::") or "reduce_kernel" in name or "Kernel2" in name):
            m = re.search(r"(\w+Functor\w*|\w+_kernel_cuda\w*|\w+Op\b|layer_norm\w*|LayerNorm\w*|cat\w*|CatArray\w*|softmax\w*|nll\w*|\w*[Cc]opy\w*)", full)
            name = (name + " :: " + (m.group(1) if m else full[len(name):len(name) + 60]))[:110]
        v = float(r[mv].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "ms": v, "msecond": v, "s": v * 1e3, "second": v * 1e3}.get(r[mu], v / 1e6)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T:.3f} ms over {sum(cnt.values())} launches (cold-cache, serialised: compare shares, not absolutes)")
    native = sum(v for k, v in tot.items() if "smb::" in k)
    print(f"native (smb::) kernels: {native:.3f} ms = {100 * native / T:.1f}%")
    for k, v in tot.most_common(top):
        print(f"{v:9.3f} ms {100 * v / T:5.1f}%  x{cnt[k]:4d}  {k}")


if __name__ == "__main__":
    DETAIL = len(sys.argv) > 3 and sys.argv[3] == "detail"
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
