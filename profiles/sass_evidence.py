"""Counts the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md, "What proves a Blackwell-native kernel")
per kernel of libgngf_sm100.so.  CPU-only: cuobjdump on the cross-compiled library.
usage: python profiles/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "collision_handling_in_instantngp_b200",
                   "libgngf_sm100.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "REDG",
            "ATOMG", "HMMA", "HGMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        for pat in PATTERNS:
            if re.search(r"\b" + pat + r"\b|\b" + pat + r"\.", line):
                cnt[cur][pat] += 1
    print("# SASS evidence per kernel of libgngf_sm100.so (sm_100a): tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR,")
    print("# tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG, mbarrier = SYNCS, cp.async = LDGSTS, red.global = REDG; legacy")
    print("# tensor paths (HMMA / HGMMA) would show up here too -- there are none.")
    names = subprocess.run(["c++filt"], input="\n".join(sorted(cnt)), capture_output=True, text=True).stdout.splitlines()
    for mangled, name in zip(sorted(cnt), names):
        short = re.sub(r"\(.*", "", name)[:72]
        print(f"{short:72s} " + "  ".join(f"{k}={v}" for k, v in sorted(cnt[mangled].items())))


if __name__ == "__main__":
    main()
