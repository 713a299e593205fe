"""Where are the local-memory (spill) accesses of a compiled object: inside loops or only in prologues/epilogues?
python tools/sass_spills.py /tmp/mrc_regs.o <kernel-mangled-substring>   (after tools/regs.py built the object)
Prints, per function region between RETs, the number of STL/LDL instructions inside backward-branch loops and outside."""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in funcs.items():
    if pat not in name:
        continue
    loops = []
    for addr, txt in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    # regions: split at CALL targets is not available; use RET boundaries as function ends
    starts = [ins[0][0]] + [a + 16 for a, t in ins if t.startswith("RET") and "@" not in t]
    def region(a):
        r = 0
        for i, s0 in enumerate(starts):
            if a >= s0: r = i
        return r
    stats = {}
    for addr, txt in ins:
        if "STL" in txt or "LDL" in txt:
            inl = any(lo <= addr <= hi for lo, hi in loops)
            depth = sum(1 for lo, hi in loops if lo <= addr <= hi)
            k = region(addr)
            stats.setdefault(k, [0, 0, 0])
            stats[k][0 if not inl else 1] += 1
            stats[k][2] = max(stats[k][2], depth)
    print(name)
    for k in sorted(stats):
        has256 = any("ENL2.256" in t or "LDG.E.EF" in t for a, t in ins if starts[k] <= a < (starts[k + 1] if k + 1 < len(starts) else 1 << 40))
        print(f"  region {k} @0x{starts[k]:x}{' (edge stream)' if has256 else ''}: outside loops {stats[k][0]}, inside loops {stats[k][1]} (max depth {stats[k][2]})")
