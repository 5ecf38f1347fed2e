#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA load / store), HMMA (legacy mma.sync), RED / ATOM (reductions).
Usage: sass_summary.py [lib.so] > profiles/sass_summary.txt    (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scat_b200", "_lib", "libscat_b200.so")
PATS = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"),
                                ("UTMASTG", r"\bUTMASTG\b"), ("HMMA", r"\bHMMA\b"), ("RED", r"\bRED\b"), ("SYNCS", r"\bSYNCS\b")])


def summarize(lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    rows, cur = collections.OrderedDict(), None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            rows[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, p in PATS.items():
            if re.search(p, ln):
                rows[cur][k] += 1
    return rows


def demangle(names):
    try:
        r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return [re.sub(r"\(.*", "", re.sub(r"scat::\(anonymous namespace\)::|void ", "", x)) for x in r]
    except Exception:
        return names


if __name__ == "__main__":
    rows = summarize()
    names = demangle(list(rows))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): occurrences per kernel")
    print(f"{'kernel':78s} " + " ".join(f"{k:>8s}" for k in PATS))
    for (fn, c), nm in zip(rows.items(), names):
        print(f"{nm[:78]:78s} " + " ".join(f"{c.get(k, 0):8d}" for k in PATS))
