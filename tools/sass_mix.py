#!/usr/bin/env python
"""Instruction mix of the dominant kernels from `cuobjdump -sass` of the built library (no GPU needed):
per kernel, the counts of the memory / atomic / synchronisation mnemonics the DESIGN argues with.
usage: sass_mix.py [libiaspgemm.so] > profiles/rNN_sass_mix.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "ia_spgemm_b200/libiaspgemm.so"
WANT = ("k_dia_mul_dia", "k_num_tiny", "k_ell_mul_ell", "k_num_global2", "k_num_hash_cta", "k_esc_warp", "k_sym_gwin", "k_num_gwin", "k_consume", "k_row_ub_thread")
KEYS = ("LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "UBLKCP", "UTMALDG", "UTMACMDFLUSH", "SYNCS", "FENCE", "DEPBAR", "BAR", "SHFL", "VOTE", "POPC", "DFMA", "DADD", "DMUL", "NANOSLEEP", "LDL", "STL")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, mix, total, seen = None, None, 0, {}
def flush():
    if cur and mix is not None:
        key = next(w for w in WANT if w in cur)
        if key not in seen or total > seen[key][1]:
            seen[key] = (cur, total, dict(mix))
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name = m.group(1)
        cur = name if any(w in name for w in WANT) else None
        mix, total = collections.Counter(), 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        total += 1
        base = op.split(".")[0]
        if base in KEYS:
            mix[op if base in ("ATOMS", "ATOMG", "RED", "UBLKCP", "SYNCS", "FENCE", "LDG", "STG") else base] += 1
flush()
print("# cuobjdump -sass %s: largest instantiation of each kernel (SASS instructions, then selected mnemonics)" % lib)
for key in WANT:
    if key not in seen:
        continue
    name, total, mix = seen[key]
    print("\n%s  (%d SASS instructions)\n  %s" % (key, total, name[:150]))
    print("  " + "  ".join("%s:%d" % kv for kv in sorted(mix.items())))
