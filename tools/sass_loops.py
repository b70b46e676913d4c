"""Static view of a kernel's loops from its SASS (no GPU needed): every backward branch = one loop, printed with its
instruction count and opcode mix.  The blend kernels are instruction-issue-bound, so the instruction count of the
per-record loop is the figure of merit when comparing source variants before spending GPU time.
    python tools/sass_loops.py topo4d_b200/_build/gs_blend.o 'blend_bwd_kernelILi2'
"""
import collections
import re
import subprocess
import sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, res = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return res


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    for name, ins in kernels(obj).items():
        if not re.search(pat, name):
            continue
        print(name, "instructions:", len(ins))
        addr_idx = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, txt) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", txt)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_idx:
                    j = addr_idx[tgt]
                    body = ins[j:i + 1]
                    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
                    print(f"  loop {tgt:#06x}..{a:#06x}: {len(body):4d} instr  " + " ".join(f"{k}={v}" for k, v in ops.most_common(12)))


if __name__ == "__main__":
    main()
