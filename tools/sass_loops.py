#!/usr/bin/env python
"""Static opcode histogram of the loops of one kernel in a built .so (no GPU needed).
usage: tools/sass_loops.py lib.so <mangled-name-substring>   -> per backward-branch loop: address range, size, opcode mix"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; funcs = {}
for l in names.splitlines():
    m = re.match(r"\s+Function : (\S+)", l)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name: continue
    print("==", name, len(ins), "instructions")
    for addr, txt in ins:
        m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < addr:
            lo = int(m.group(1), 16)
            body = [t for a, t in ins if lo <= a <= addr]
            if len(body) < 40: continue
            h = collections.Counter()
            for t in body:
                parts = t.split(); op = parts[1] if parts[0].startswith("@") else parts[0]
                fam = op.split(".")[0]
                if op.startswith("IMAD.MOV"): fam = "IMAD.MOV"
                h[fam] += 1
            dp = sum(h[x] for x in ("DADD", "DMUL", "DFMA", "DSETP"))
            print(f"  loop {lo:#06x}-{addr:#06x}: {len(body)} instr, fp64 {dp}; " + " ".join(f"{k}:{v}" for k, v in h.most_common(22)))
