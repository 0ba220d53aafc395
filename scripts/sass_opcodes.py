"""Per-kernel counts of the SASS opcodes that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP (TMA / bulk copies), HMMA (legacy mma.sync), MUFU.EX2, and the
ELECT/R2UR.BROADCAST waterfall count.  Reads the built libdrag_b200.so with cuobjdump (no GPU needed).
    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ai-dial-rag_b200", "dial_rag_b200", "_lib", "libdrag_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU.EX2", "MUFU.TANH",
       "R2UR.BROADCAST", "SYNCS", "USETMAXREG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cut = name.rfind(">(")
            name = name[: cut + 1] if cut >= 0 else re.sub(r"\(.*", "", name)
            per[name] = {}
            continue
        if name is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for key in OPS:
            if op == key or (op.startswith(key + ".") and key != "UTCHMMA") or (key == "UTCHMMA" and op.startswith("UTCHMMA") and ".2CTA" not in op) \
                    or (key == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                per[name][key] = per[name].get(key, 0) + 1
    print(f"# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (sm_100a), cuobjdump -sass")
    print("# " + " ".join(f"{k:>14s}" for k in OPS) + "  kernel")
    tot = {k: 0 for k in OPS}
    for name in sorted(per):
        if not any(per[name].values()):
            continue
        print("  " + " ".join(f"{per[name].get(k, 0):14d}" for k in OPS) + "  " + name)
        for k in OPS:
            tot[k] += per[name].get(k, 0)
    print("  " + " ".join(f"{tot[k]:14d}" for k in OPS) + "  TOTAL")


if __name__ == "__main__":
    main()
