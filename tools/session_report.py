"""Summarise one tools/gpu_session.sh run: every bench line and microbench table under gpurun_out/<tag>_* in two tables, with the
difference of each A/B arm to the default arm, so the opt-in switches can be judged at a glance.

    python tools/session_report.py r2a [gpurun_out]
"""
import glob
import json
import os
import sys


def load_line(path):
    try:
        txt = open(path).read().strip().splitlines()
    except OSError:
        return None
    for line in reversed(txt):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                continue
    return None


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "s"
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
    benches = {}
    for p in sorted(glob.glob(os.path.join(out, f"{tag}_bench*.json"))):
        d = load_line(p)
        if d and "ms_per_step" in d:
            benches[os.path.basename(p)[len(tag) + 1:-5]] = d
    base = benches.get("bench")
    ops = sorted({k for d in benches.values() for k in (d.get("native_ms_per_step") or {})})
    print(f"{'arm':28s} {'ms/step':>9s} {'delta':>8s} {'patches/s':>10s} {'e2e':>8s} {'fwd frac':>9s}  " + " ".join(f"{o[:11]:>11s}" for o in ops))
    for name, d in benches.items():
        delta = (d["ms_per_step"] - base["ms_per_step"]) if base else float("nan")
        e2e = (d.get("e2e") or {}).get("value")
        roof = (d.get("roofline") or {}).get("frac")
        nat = d.get("native_ms_per_step") or {}
        print(f"{name:28s} {d['ms_per_step']:9.2f} {delta:+8.2f} {d['value']:10.2f} {e2e if e2e is None else round(e2e, 2)!s:>8s} "
              f"{roof if roof is None else round(roof, 3)!s:>9s}  " + " ".join(f"{nat.get(o, float('nan')):11.2f}" for o in ops))
        reasons = (d.get("clocks") or {}).get("reasons")
        if reasons and any(r != "sw_power_cap" for r in reasons):
            print(f"{'':28s} !! clocks: {d['clocks']}")
    mbs = {}
    for p in sorted(glob.glob(os.path.join(out, f"{tag}_mb*.json"))):
        try:
            mbs[os.path.basename(p)[len(tag) + 1:-5]] = json.load(open(p))
        except (OSError, ValueError):
            pass
    if mbs:
        keys = ("scan_fwd_ms", "scan_bwd_ms", "conv_fwd_ms", "conv_bwd_ms")
        print()
        print(f"{'microbench':24s} {'dtype':>5s} {'dim':>4s} {'L':>7s} " + " ".join(f"{k:>12s}" for k in keys) + f" {'fwd frac':>9s} {'ref fwd':>8s} {'ref bwd':>8s}")
        for name, d in mbs.items():
            for r in d.get("rows", []):
                print(f"{name:24s} {r.get('dtype', ''):>5s} {r.get('dim', 0):4d} {r.get('L', 0):7d} "
                      + " ".join(f"{r.get(k, float('nan')):12.4f}" for k in keys)
                      + f" {r.get('scan_fwd_frac', float('nan')):9.3f} {r.get('ref_scan_fwd_ms', float('nan')):8.3f} {r.get('ref_scan_bwd_ms', float('nan')):8.3f}")
    for p in sorted(glob.glob(os.path.join(out, f"{tag}_pytest*.log"))):
        tail = [l for l in open(p).read().splitlines() if l.strip()][-2:]
        print(f"\n{os.path.basename(p)}: " + " | ".join(tail))


if __name__ == "__main__":
    main()
