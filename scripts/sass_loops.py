"""Instruction mix of the loops of one kernel, from `cuobjdump -sass` (no GPU needed).

    python scripts/sass_loops.py <lib.so> <kernel-name-substring> [--min 40] [--dump N]

A loop = a backward branch; its body = the address range [target, branch].  Prints, per loop (innermost first is up to
the reader: ranges nest), the opcode histogram grouped into fp64 / shared-memory / global-memory / integer+move /
control.  Used for profiles/r2_refl_toa_v5_sass.txt ("commit the SASS of the consume loop").
"""
import collections
import re
import subprocess
import sys


def kernel_sass(lib, name):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    out, on = [], False
    for line in txt:
        if "Function :" in line:
            on = name in line
            if on:
                out.append(line)
            continue
        if on:
            out.append(line)
    return out


INS = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")


def parse(lines):
    ins = []
    for l in lines:
        m = INS.search(l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def group(op):
    o = op.split()[0] if not op.startswith("@") else op.split()[1]
    o = o.split(".")[0]
    if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "MUFU", "F2F", "I2F", "F2I", "DFMA2"):
        return "fp64" if o != "MUFU" else "mufu"
    if o in ("LDS", "STS", "LDSM"):
        return "shared"
    if o in ("LDG", "STG", "LD", "ST", "LDC", "LDCU", "ULDC", "CCTL", "PREFETCH"):
        return "global/const"
    if o in ("BRA", "BAR", "EXIT", "CALL", "RET", "BSSY", "BSYNC", "WARPSYNC", "NOP", "YIELD", "BREAK"):
        return "control"
    return "int/move/pred"


def main():
    lib, name = sys.argv[1], sys.argv[2]
    minlen = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 40
    dump = int(sys.argv[sys.argv.index("--dump") + 1]) if "--dump" in sys.argv else -1
    lines = kernel_sass(lib, name)
    ins = parse(lines)
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print(lines[0].strip())
    print("instructions:", len(ins))
    loops = []
    for i, (a, op) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(?:P\d,\s*)?(0x[0-9a-f]+)", op)
        if m:
            t = int(m.group(1), 16)
            if t <= a and t in addr_index:
                loops.append((addr_index[t], i))
    for n, (b, e) in enumerate(loops):
        body = ins[b:e + 1]
        if len(body) < minlen:
            continue
        g = collections.Counter(group(op) for _, op in body)
        ops = collections.Counter((op.split()[1] if op.startswith("@") else op.split()[0]).split(".")[0] for _, op in body)
        print(f"\nloop {n}: 0x{ins[b][0]:04x} .. 0x{ins[e][0]:04x}  {len(body)} instructions  " +
              "  ".join(f"{k} {v}" for k, v in sorted(g.items(), key=lambda kv: -kv[1])))
        print("   " + "  ".join(f"{k} {v}" for k, v in ops.most_common(24)))
        if n == dump:
            for a, op in body:
                print(f"      /*{a:04x}*/ {op}")


if __name__ == "__main__":
    main()
