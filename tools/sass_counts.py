"""Per-kernel SASS mnemonic counts of libsd3d.so (evidence that the TMA / tcgen05 / mbarrier paths are what was built).
usage: python tools/sass_counts.py <kernel-name-substring> [<mnemonic-prefix> ...]   (run where cuobjdump is on PATH)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = ("UTMALDG", "UTMASTG", "UBLKCP", "UTCHMMA", "UTCBAR", "UTCATOM", "LDTM", "SYNCS", "LDG", "STG", "LDS", "STS",
           "REDUX", "FFMA2", "FMUL2", "FADD2", "RED", "ATOM")


def main():
    pat = sys.argv[1]
    keys = tuple(sys.argv[2:]) or DEFAULT
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "segdino3d_b200", "libsd3d.so")], capture_output=True,
                          text=True, check=True).stdout
    filt = subprocess.run(["c++filt"], input=sass, capture_output=True, text=True).stdout
    name, counts, total = None, None, 0
    out = []
    def flush():
        if name and pat in name:
            out.append((name, total, dict(sorted(counts.items()))))
    for line in filt.splitlines():
        m = re.match(r"\s*Function : (.*)", line)
        if m:
            flush()
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            total += 1
            op = m.group(1)
            if op.startswith(keys):
                counts[op] += 1
    flush()
    for n, t, c in out:
        print(n)
        print(f"    instructions: {t};  " + ", ".join(f"{k}={v}" for k, v in c.items()))


if __name__ == "__main__":
    main()
