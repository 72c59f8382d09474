"""Build tuning variants of the native library for A/B runs on the GPU box (they travel with the gpurun snapshot; select one with
SMB_LIB=segmamba_b200/variants/lib_<name>.so).  Each variant is the same source with -D overrides of the build-time knobs in
csrc/scan_internal.h, scan_fwd.cu, scan_bwd.cu, conv1d.cu, instnorm.cu; the default library is not touched.  (Tried and dropped
after looking at ptxas output: a 128-register cap on the forward main pass gains no occupancy because 4 CTAs do not fit in shared
memory; 2-warp CTAs only make the compiler spend more registers.)  A variant library is ~26 MB: build them right before the gpurun
call that uses them and delete them afterwards.

    python tools/build_variants.py [name ...]        (no names: all)      -> prints registers / spills of the affected kernels
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "segmamba_b200", "csrc")
OUTD = os.path.join(ROOT, "segmamba_b200", "variants")
SRCS = ["capi.cu", "scan_fwd.cu", "scan_fwd_v2.cu", "scan_bwd.cu", "scan_bwd_v2.cu", "scan_bwd_r3v2.cu", "conv1d.cu", "conv1d_v2.cu", "instnorm.cu", "layernorm.cu"]
NVFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler",
           "-fvisibility=hidden", "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++", "-Xptxas", "-v"]

VARIANTS = {
    # instance norm: one more resident CTA per SM for the apply / stats kernels (register caps 85 -> 64, 128 -> 85; <= 260 B spills)
    "in_minb4": (["-DSMB_IN_FWD_MINB=4", "-DSMB_IN_BWD_MINB0=4", "-DSMB_IN_BWD_MINB2=3"], ["instnorm.cu"], r"in_(apply|stats)"),
    # instance norm: the apply passes walk each CTA's row range top-down (what the preceding statistics pass left in L2 comes first)
    "in_rev": (["-DSMB_IN_APPLY_REVERSE=1"], ["instnorm.cu"], r"in_apply"),
    "in_rev_stats": (["-DSMB_IN_STATS_REVERSE=1"], ["instnorm.cu"], r"in_stats"),
    # R3 with a register cap for 3 CTAs per SM (80 registers, ~0.5 KB of spills; its 55 KB of shared memory allow it)
    "r3_minb3": (["-DSMB_R3_MINB=3"], ["scan_bwd.cu"], r"scan_bwd_main_kernel"),
    # forward scan and backward R1 with 2 / 4 of the 8 state pairs on the packed-FMA polynomial ex2 (slower on the lockstep kernels, DESIGN.md 3.1;
    # meant for the software-pipelined kernels, SMB_FWD_V2=1, whose MUFU utilisation should be higher)
    "poly11": (["-DSMB_POLY_MASK=0x11"], ["scan_fwd.cu", "scan_fwd_v2.cu", "scan_bwd.cu", "scan_bwd_v2.cu"], r"scan_(fwd_(agg|main)|bwd_ragg)"),
    "poly33": (["-DSMB_POLY_MASK=0x33"], ["scan_fwd.cu", "scan_fwd_v2.cu", "scan_bwd.cu", "scan_bwd_v2.cu"], r"scan_(fwd_(agg|main)|bwd_ragg)"),
    # conv1d with 256-thread CTAs
    "conv256": (["-DSMB_CONV_THREADS=256"], ["conv1d.cu"], r"conv1d_(fwd|bwd)_kernel"),
}


def build(name):
    defs, touched, pat = VARIANTS[name]
    odir = os.path.join(OUTD, name)
    os.makedirs(odir, exist_ok=True)
    base_odir = CSRC                                              # untouched translation units reuse the default objects
    objs, logs = [], {}

    def compile_one(src):
        obj = os.path.join(odir, src.replace(".cu", ".o"))
        r = subprocess.run(["/usr/local/cuda/bin/nvcc"] + NVFLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-3000:])
        return src, obj, r.stderr
    with ThreadPoolExecutor(max_workers=4) as ex:
        for src, obj, log in ex.map(compile_one, touched):
            logs[src] = log
    for src in SRCS:
        objs.append(os.path.join(odir, src.replace(".cu", ".o")) if src in touched else os.path.join(base_odir, src.replace(".cu", ".o")))
    out = os.path.join(OUTD, f"lib_{name}.so")
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-shared", "-o", out] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart"], check=True)
    # registers / spills of the kernels the variant is about (bf16 instantiations)
    rows = []
    for log in logs.values():
        cur = None
        for line in log.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = m.group(1)
            m2 = re.search(r"Used (\d+) registers", line)
            sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if sp and cur:
                spill = (int(sp.group(1)), int(sp.group(2)))
            if m2 and cur and re.search(pat, cur) and "bfloat16" in cur:
                rows.append((re.sub(r"^_ZN3smb\d+", "", cur)[:70], int(m2.group(1)), spill))
    return out, rows


if __name__ == "__main__":
    subprocess.run(["make", "-s", "-C", CSRC, "-j8", "all"], check=True)     # default objects must exist
    for name in (sys.argv[1:] or list(VARIANTS)):
        out, rows = build(name)
        print(f"== {name}: {os.path.relpath(out, ROOT)}   ({' '.join(VARIANTS[name][0])})")
        for k, regs, spill in sorted(set(rows)):
            print(f"   {k:70s} {regs:4d} regs  spill {spill[0]}/{spill[1]} B")
