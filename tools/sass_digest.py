"""Per-kernel SASS digest of libwbk.so: instruction counts that characterise each kernel (FP64 math, global / shared
memory, asynchronous copies, barriers, warp collectives, tensor-core / TMA opcodes).

  python tools/sass_digest.py > profiles/<round>_sass_digest.md        (cuobjdump only, no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wavebreaking_b200", "libwbk.so")
GROUPS = collections.OrderedDict([
    ("FP64 (DADD/DMUL/DFMA/DSETP)", r"^(DADD|DMUL|DFMA|DSETP)"),
    ("FP32 math", r"^(FADD|FMUL|FFMA|FSETP|MUFU)"),
    ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS/STS", r"^(LDS|STS)"),
    ("LDGSTS (cp.async)", r"^LDGSTS"), ("UTMALDG/UTMASTG/UBLKCP (TMA)", r"^(UTMALDG|UTMASTG|UBLKCP)"),
    ("UTC*MMA / HMMA (tensor)", r"^(UTC|HMMA|IMMA|DMMA)"),
    ("ATOM/RED", r"^(ATOM|ATOMG|ATOMS|RED)"), ("BAR", r"^BAR"), ("SHFL", r"^SHFL"), ("VOTE", r"^VOTE"),
])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for g, pat in GROUPS.items():
                if re.match(pat, op):
                    kernels[cur][g] += 1
    print("| kernel | SASS instr | " + " | ".join(GROUPS) + " |")
    print("|---|---|" + "---|" * len(GROUPS))
    for k, c in kernels.items():
        if c["total"] == 0:
            continue
        print("| `{}` | {} | ".format(k[:90], c["total"]) + " | ".join(str(c[g]) for g in GROUPS) + " |")


if __name__ == "__main__":
    main()
