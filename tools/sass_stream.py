"""Dev helper: print the opcode stream of one kernel from an object file as one character per SASS instruction
(F packed FMA/MUL, a packed add, s scalar FP, L/S shared load/store, g/G global, m MOV, | CTA barrier, w warp barrier)."""
import subprocess, sys
obj, fun = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
cls = {'FFMA2': 'F', 'FMUL2': 'F', 'FADD2': 'a', 'FADD': 's', 'FMUL': 's', 'FFMA': 's', 'LDS': 'L', 'STS': 'S', 'MOV': 'm',
       'BAR': '|', 'WARPSYNC': 'w', 'SYNCS': 'Y', 'BRA': 'b', 'STG': 'G', 'LDG': 'g', 'MUFU': 'u', 'LDL': 'l', 'STL': 'x'}
ops = []
for ln in out.splitlines():
    p = ln.split()
    if len(p) > 1 and p[0].startswith('/*') and len(p[0]) == 8:
        op = p[1] if not p[1].startswith('@') else p[2]
        ops.append(op.split('.')[0].rstrip(';'))
line = ''.join(cls.get(o, '.') for o in ops)
print(len(ops), "instructions")
for i in range(0, len(line), 150):
    print(line[i:i + 150])
