#!/usr/bin/env python3
"""Instruction histogram of the consumer loop of psd_score_ws_kernel<F=HSV> in a built library
(no GPU needed):  python tools/sass_loop_stats.py [path/to/lib.so] [--list]
Classes follow the pipe model measured in profiles/r01*_pipes*.txt (alu half / fma half / wide / other)."""
import collections
import re
import subprocess
import sys

FUNC = "_ZN3psd19psd_score_ws_kernelILj1EEEvNS_9ScoreArgsE"
LINE = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Za-z0-9_.]+)")
ALU = ("PRMT", "LOP3", "SHF", "VIMNMX", "VHMNMX", "HSET2", "FMNMX3", "VABSDIFF", "SEL", "ISETP", "PLOP3", "LEA",
       "VIADD", "MOV", "FSEL", "FSETP", "I2IP", "F2FP", "SGXT", "BMSK")
FMA = ("IMAD", "IDP", "HFMA2", "HADD2", "HMUL2")
WIDE = ("FFMA", "FADD", "FMUL", "IADD3", "FMNMX", "HMNMX2")


def classify(op):
    if op.startswith("IMAD.MOV") or op.startswith("IMAD.IADD") or op.startswith(FMA):
        return "fma"
    if op.startswith("FMNMX3"):
        return "alu"
    if op.startswith(WIDE):
        return "wide"
    if op.startswith(ALU):
        return "alu"
    return "other"


def loop_rows(lib):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", FUNC, lib], capture_output=True, text=True).stdout
    rows = [(int(m.group(1), 16), m.group(2), l) for l in out.splitlines() if (m := LINE.match(l))]
    idx = {a: i for i, (a, _, _) in enumerate(rows)}
    bodies = []
    for i, (a, op, l) in enumerate(rows):
        m = re.search(r"BRA\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", l)
        if op.startswith("BRA") and m and int(m.group(1), 16) < a and int(m.group(1), 16) in idx:
            bodies.append(rows[idx[int(m.group(1), 16)]:i + 1])
    # the consumer loop is the innermost loop around the most LDS.128 (3 per frame, 2 or 4 frames per body)
    def n_lds(b):
        return sum(op == "LDS.128" for _, op, _ in b)
    cands = [b for b in bodies if n_lds(b) >= 6 and n_lds(b) % 3 == 0]
    best = max(n_lds(b) for b in cands if n_lds(b) <= 12)
    return min((b for b in cands if n_lds(b) == best), key=len), len(rows)


def main():
    lib = next((a for a in sys.argv[1:] if not a.startswith("-")), "pyscenedetect_b200/libpsd_b200.so")
    body, total = loop_rows(lib)
    frames = sum(op == "LDS.128" for _, op, _ in body) // 3
    hist = collections.Counter(op for _, op, _ in body)
    cls = collections.Counter(classify(op) for _, op, _ in body)
    print(f"{lib}: kernel {total} instr; consumer loop {len(body)} instr for {frames} frames = "
          f"{len(body) / frames:.1f} per frame = {len(body) / frames / 16:.2f} per pixel")
    print("  classes per frame:", {k: round(v / frames, 1) for k, v in cls.items()})
    print("  " + "  ".join(f"{op}:{n / frames:g}" for op, n in hist.most_common()))
    if "--list" in sys.argv:
        for _, _, l in body:
            print(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).rstrip())


if __name__ == "__main__":
    main()
