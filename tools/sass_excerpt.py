#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libhafgpu.so, the count and first occurrence of the tensor-core / TMEM / TMA /
cluster / atomic instructions (cuobjdump -sass; runs without a GPU).

  python tools/sass_excerpt.py > profiles/rNN_sass_excerpt.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "haf_grasping_b200", "lib", "libhafgpu.so")
PAT = re.compile(r"\b(UTCHMMA[.\w]*|UTCQMMA[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|UCGABAR\w*|"
                 r"SYNCS[.\w]*|ELECT|MUFU\.EX2|DMMA[.\w]*|HMMA[.\w]*|LDGSTS[.\w]*|REDG[.\w]*|ATOMG[.\w]*|RED\.[.\w]*)")
KEEP = ("svm_rbf_tc", "guard_dmma", "guard_fma", "bin_maxz_cloud", "features_tc")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print("# SASS evidence (cuobjdump -sass haf_grasping_b200/lib/libhafgpu.so, built by haf_grasping_b200/build.py for sm_100a)")
    print("# per kernel: count of the tensor-core / TMEM / TMA / cluster / atomic instructions, then the first occurrence of each\n")
    fn, stats = None, None

    def flush():
        if fn and stats and any(k in fn for k in KEEP):
            print("## " + fn)
            for op, (n, first) in sorted(stats.items(), key=lambda kv: -kv[1][0]):
                print("  %-34s x%-5d %s" % (op, n, first))
            print()

    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            flush()
            fn, stats = m.group(1), collections.OrderedDict()
            continue
        if fn is None or "/*" not in ln:
            continue
        m = PAT.search(ln)
        if m:
            op = m.group(1)
            n, first = stats.get(op, (0, ln.strip()[:150]))
            stats[op] = (n + 1, first)
    flush()


if __name__ == "__main__":
    sys.exit(main())
