#!/usr/bin/env python3
"""Static SASS helper: per-function opcode histogram and the backward-branch loops with their opcode mix.
usage: sass_loops.py <file.sass | lib.so> <function-substring> [--loop N]"""
import re, sys, subprocess, collections
def load(path):
    if path.endswith(".so") or path.endswith(".cubin"):
        return subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    return open(path).read()
def funcs(txt):
    out = {}; cur = None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m: cur = m.group(1); out[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", ln)
        if m and cur is not None:
            ins = m.group(2); ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            out[cur].append((int(m.group(1), 16), ins))
    return out
def op(ins): return ins.split()[0]
def main():
    txt = load(sys.argv[1]); F = funcs(txt)
    for name, body in F.items():
        if sys.argv[2] not in name: continue
        print("==", name, len(body), "instructions")
        addr = {a: i for i, (a, _) in enumerate(body)}
        loops = []
        for i, (a, ins) in enumerate(body):
            m = re.search(r"\bBRA\S*\s+.*?(0x[0-9a-f]+)", ins)
            if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] <= i:
                loops.append((addr[int(m.group(1), 16)], i))
        for k, (s, e) in enumerate(loops):
            h = collections.Counter(op(x) for _, x in body[s:e + 1])
            print(" loop %d: [%x..%x] %d instr; top: %s" % (k, body[s][0], body[e][0], e - s + 1, ", ".join("%s %d" % kv for kv in h.most_common(14))))
        if "--loop" in sys.argv:
            k = int(sys.argv[sys.argv.index("--loop") + 1]); s, e = loops[k]
            for a, x in body[s:e + 1]: print("  %05x  %s" % (a, x))
if __name__ == "__main__": main()
