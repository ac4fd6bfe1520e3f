"""Blackwell-specific SASS mnemonics per kernel of the shipped library (cuobjdump -sass): the tcgen05 / TMA / TMEM evidence.
usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "reppo_b200", "libreppo_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTMALDG(?:\.\dD)?(?:\.MULTICAST)?|UTMASTG|UTMAREDG|LDTM|STTM|UTCBAR(?:\.MULTICAST)?|"
                 r"UCGABAR_[A-Z]+|ACQBULK|LDGMC[A-Z_.0-9x]*|STGMC[A-Z_.0-9x]*|REDGMC[A-Z_.0-9x]*)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = PAT.search(line) if cur else None
        if m:
            cnt[cur][m.group(1)] += 1
    print("# cuobjdump -sass reppo_b200/libreppo_b200.so (sm_100a): Blackwell-specific mnemonics per kernel (instruction counts)")
    print("# UTC*MMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
    print("# UCGABAR = cluster barrier, ACQBULK = griddepcontrol.wait (programmatic dependent launch), LDGMC / STGMC = multimem.ld_reduce / multimem.st (NVSwitch multicast)")
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    for mangled, name in sorted(zip(cnt, names), key=lambda x: x[1]):
        c = cnt[mangled]
        if not any(k.startswith(("UTC", "UTMA", "LDTM", "LDGMC", "STGMC")) for k in c):
            continue
        short = re.sub(r"\(.*$", "", name.replace("(reppo::", "<").replace("(int)", "")) if "<" not in name else name[:name.rfind(">") + 1]
        print(f"{short[:140]}\n    " + ", ".join(f"{a}={b}" for a, b in sorted(c.items())))


if __name__ == "__main__":
    main()
