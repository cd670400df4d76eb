"""Opcode sequence of the hottest loop body of a kernel, from cuobjdump (no GPU needed):
   python tools/sass_loop.py <object> <function substring> <first opcode marker count...>
Prints the opcode stream between the first and last MUFU.EX2 of the function's main loop as a compact string:
M = MUFU, f = FFMA2/FADD2/FMUL2 (packed fp32), p = F2FP, s = STS, t = LDTM, . = anything else."""
import subprocess
import sys

obj, fn = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(out) if "Function :" in l and fn in l)
end = next((i for i in range(start + 1, len(out)) if "Function :" in out[i]), len(out))
ops = []
for l in out[start:end]:
    l = l.strip()
    if not l.startswith("/*") or ";" not in l:
        continue
    body = l.split("*/", 1)[1].strip()
    if not body or body.startswith("/*"):
        continue
    t = body.split()
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    ops.append(op)
sym = {"MUFU": "M", "FFMA2": "f", "FADD2": "f", "FMUL2": "f", "F2FP": "p", "STS": "s", "LDTM": "t", "SYNCS": "B", "BRA": "b"}
s = "".join(sym.get(o.split(".")[0], ".") for o in ops)
print(len(ops), "instructions")
for i in range(0, len(s), 120):
    print(f"{i:5d} {s[i:i + 120]}")
