"""Blackwell-native instruction census of the built library -> profiles/r02_sass_tcgen05.md
    python tools/sass_summary.py     (runs on the CPU box: cuobjdump -sass of ha2g_b200/csrc/libha2g_b200.so)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ha2g_b200", "csrc", "libha2g_b200.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "HMMA"]


def demangle(n):
    s = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    s = s.replace("(anonymous namespace)::", "")
    m = re.search(r"([A-Za-z_]\w*(?:<[^()]*>)?)\(", s)
    return m.group(1) if m else s


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    pat = re.compile(r"\b(" + "|".join(COLS + ["UTCQMMA", "UTMALDG", "UTMASTG", "HGMMA"]) + r")\b")
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
        elif cur:
            for t in pat.findall(line):
                per[cur][t] += 1
    tot = collections.Counter()
    lines = ["# r02 -- Blackwell-native instructions in the built library (`cuobjdump -sass ha2g_b200/csrc/libha2g_b200.so`)", "",
             "SASS mnemonic -> PTX: `UTCHMMA` = `tcgen05.mma.kind::f16/tf32`, `LDTM`/`STTM` = `tcgen05.ld`/`st`, `UTCBAR` = "
             "`tcgen05.commit`, `UBLKCP` = `cp.async.bulk` (1-D bulk copies: the packed operand layouts make every tile a "
             "contiguous run, so no tensor map is needed -- hence no `UTMALDG`), `SYNCS` = mbarrier arrive / expect_tx / "
             "try_wait, `HMMA` = legacy `mma.sync` (none).", "",
             "| kernel | " + " | ".join(COLS) + " |", "|---|" + "---:|" * len(COLS)]
    for k, c in per.items():
        tot.update(c)
        if c["UTCHMMA"] or c["LDTM"] or c["UBLKCP"] or c["STTM"]:
            lines.append("| `" + demangle(k) + "` | " + " | ".join(str(c[t]) for t in COLS) + " |")
    lines.append("| **library total** | " + " | ".join(str(tot[t]) for t in COLS) + " |")
    open(os.path.join(ROOT, "profiles", "r02_sass_tcgen05.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[6:]))


if __name__ == "__main__":
    main()
