"""Compare the SASS of two builds of the library kernel by kernel (addresses and encodings stripped): which kernels are
bit-identical in instruction stream, which changed, which are new.  Used to show that adding opt-in kernels left the
hardware-verified default kernels untouched.      usage: python tools/sass_diff.py old.so new.so"""
import hashlib
import re
import subprocess
import sys


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append(m.group(1).strip())
    return {k: (len(v), hashlib.sha1("\n".join(v).encode()).hexdigest()[:12]) for k, v in out.items()}


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = [k for k in a if k in b and a[k] == b[k]]
    changed = [k for k in a if k in b and a[k] != b[k]]
    print(f"{len(same)} kernels identical, {len(changed)} changed, {len(set(b) - set(a))} new, {len(set(a) - set(b))} removed")
    for k in changed:
        print("  CHANGED", k[:110], a[k], "->", b[k])
    for k in sorted(set(a) - set(b)):
        print("  REMOVED", k[:110])
    fam = {}
    for k in set(b) - set(a):
        f = re.search(r"\d+([a-z][a-z0-9_]*?_kernel)I", k) or re.search(r"\d+([a-z][a-z0-9_]*?_kernel)", k)
        fam[f.group(1) if f else k] = fam.get(f.group(1) if f else k, 0) + 1
    for f, n in sorted(fam.items()):
        print(f"  NEW {f} x{n}")


if __name__ == "__main__":
    main()
