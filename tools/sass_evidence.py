"""Per-kernel SASS instruction counts of libttn_b200.so (`cuobjdump -sass`, demangled with `cu++filt`): the mnemonics that identify
the hardware path of each kernel.  usage: python tools/sass_evidence.py > profiles/sass_evidence_rNN.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tensortrainnumerics.jl_b200", "libttn_b200.so")
COLS = ["DMMA", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "DFMA", "BAR.SYNC", "UCGABAR", "SHFL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = dict.fromkeys(COLS, 0)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for c in COLS:
            if op == c or op.startswith(c + "."):
                funcs[cur][c] += 1
    names = list(funcs)
    dem = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass libttn_b200.so (sm_100a), per kernel: instruction counts that identify the hardware path")
    print("#   DMMA = FP64 tensor MMA (mma.sync.m8n8k4.f64), LDGSTS = cp.async, UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier ops,")
    print("#   UTMALDG = tensor-map TMA (not used: the bulk form moves the contiguous tile rows), UCGABAR = cluster barrier")
    print(f"# {'kernel':84s}" + "".join(f"{c:>10s}" for c in COLS))
    only = sys.argv[1:] or None
    for n, d in zip(names, dem):
        d = re.sub(r"^void\s+", "", d)
        d = re.sub(r"\((int|bool|unsigned int)\)", "", d)
        d = re.sub(r"\(.*", "", d).replace("ttn::(anonymous namespace)::", "").replace("<unnamed>::", "").replace("ttn::", "")
        if only and not any(o in d for o in only):
            continue
        c = funcs[n]
        if not any(c.values()):
            continue
        print(f"{d[:84]:84s}" + "".join(f"{c[k]:10d}" for k in COLS))


if __name__ == "__main__":
    main()
