#!/usr/bin/env python
"""Instruction mix of a kernel's loops from its SASS (cuobjdump -sass of an object file or library).

  python profiles/tools/sass_mix.py gbnet_b200/build/sweep_fast.o k_fast_sILi0ELb1ELb0E [--loops] [--dump]

Finds backward branches (loop back-edges), and for every loop body (the instructions between the
branch target and the branch) prints the opcode histogram grouped by pipe class.  Static counts: a
loop body with predicated-off or branched-over regions over-counts what a thread executes, so the
listing also shows forward branches inside the body.
"""
import collections
import re
import subprocess
import sys

CLASSES = [
    ("fp64", r"^(DMUL|DADD|DFMA|DSETP|DMNMX)"),
    ("fp32", r"^(FMUL|FADD|FFMA|FSETP|FSEL|FMNMX|MUFU)"),
    ("int-mul", r"^(IMAD|IMUL)"),
    ("int-alu", r"^(IADD|IADD3|LOP3|LOP|SHF|SHL|SHR|LEA|ISETP|SEL|ICMP|POPC|FLO|BREV|PRMT|VIADD|VIMNMX|IABS|SGXT|BMSK|PLOP3|P2R|R2P)"),
    ("cvt", r"^(I2F|F2I|F2F|I2FP|F2FP|I2I)"),
    ("mov", r"^(MOV|UMOV|S2R|S2UR|CS2R|R2UR|UIADD3|ULDC|ULOP3|USHF|UIMAD|ULEA|USEL|UISETP|UPLOP3|UFLO|UPOPC)"),
    ("ld-global", r"^(LDG|LD\b|LD\.)"),
    ("st-global", r"^(STG|ST\b|ST\.|RED|ATOMG|ATOM)"),
    ("ld-shared", r"^(LDS|LDSM)"),
    ("st-shared", r"^(STS)"),
    ("ld-const", r"^(LDC)"),
    ("local", r"^(LDL|STL)"),
    ("branch", r"^(BRA|BRX|BSSY|BSYNC|EXIT|RET|CALL|WARPSYNC|BAR|BREAK|YIELD|NANOSLEEP|DEPBAR|ERRBAR|MEMBAR|CCTL)"),
    ("shuffle/vote", r"^(SHFL|VOTE|VOTEU|MATCH|REDUX)"),
]


def classify(op):
    for name, rx in CLASSES:
        if re.match(rx, op):
            return name
    return "other"


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", line)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    return funcs


def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0] if t else "?"


def main():
    path, pat = sys.argv[1], sys.argv[2]
    dump = "--dump" in sys.argv
    funcs = functions(path)
    for name, ins in funcs.items():
        if pat not in name:
            continue
        print(f"== {name}: {len(ins)} instructions")
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)*`?\(?\.?L?_?x?_?\w*\)?\s*0x([0-9a-f]+)", t) or re.search(r"BRA.*0x([0-9a-f]+)", t)
            if m and opcode(t).startswith("BRA"):
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_index:
                    loops.append((addr_index[tgt], i))
        tot = collections.Counter(classify(opcode(t)) for _, t in ins)
        print("   whole kernel:", dict(tot))
        for (b, e) in sorted(loops, key=lambda x: x[0] - x[1]):
            body = ins[b:e + 1]
            hist = collections.Counter(classify(opcode(t)) for _, t in body)
            ops = collections.Counter(opcode(t).split(".")[0] for _, t in body)
            fwd = sum(1 for a, t in body if opcode(t).startswith("BRA")) - 1
            print(f"-- loop 0x{ins[b][0]:x} .. 0x{ins[e][0]:x}: {len(body)} instructions, {fwd} inner branches")
            print("   by class:", ", ".join(f"{k} {v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])))
            print("   by opcode:", ", ".join(f"{k} {v}" for k, v in ops.most_common(40)))
            if dump:
                for a, t in body:
                    print(f"      {a:05x}  {t}")


if __name__ == "__main__":
    main()
