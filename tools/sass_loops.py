"""Static view of a kernel's loops: for every backward branch in the SASS of the kernels matching a regex, the number of
instructions in the loop body and a histogram of its opcodes (what an instruction diet is checked against before any GPU time
is spent).  usage: python tools/sass_loops.py <lib.so> <kernel regex> [min body size]"""
import collections
import re
import subprocess
import sys


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, body


def main():
    lib, pat = sys.argv[1], re.compile(sys.argv[2])
    min_body = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    for name, body in kernels(lib):
        if not pat.search(name):
            continue
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print("%s: %d instructions" % (demangled[:110], len(body)))
        addr_index = {a: i for i, (a, _) in enumerate(body)}
        for i, (a, ins) in enumerate(body):
            m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", ins)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt >= a or tgt not in addr_index:
                continue
            j = addr_index[tgt]
            n = i - j + 1
            if n < min_body:
                continue
            ops = collections.Counter()
            for _, s in body[j:i + 1]:
                s = re.sub(r"^@!?U?P\d+\s+", "", s)
                ops[s.split()[0].split(".")[0]] += 1
            print("  loop [%d..%d] %d instructions: %s" % (j, i, n, ", ".join("%s %d" % kv for kv in ops.most_common(14))))


if __name__ == "__main__":
    main()
