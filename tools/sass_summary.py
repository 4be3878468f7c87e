"""SASS evidence of the hot kernels -> profiles/rN_sass_hot_kernels.txt

    python tools/sass_summary.py profiles/r2_sass_hot_kernels.txt

Per kernel of libudal.so (cuobjdump -sass, demangled): instruction count, the tcgen05 / TMA / bulk-copy / packed-math
mnemonics, and the first tensor-core MMA group as it appears in the SASS."""
import collections
import re
import subprocess
import sys

LIB = "uncertainty-detection-autolabeling_b200/libudal.so"
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HFMA2", "FFMA2", "FADD2", "FMUL2", "MUFU", "DFMA"]
HOT = re.compile(r"heads_dw|heads_l1|heads_wide|nms_epoch|decode_stream|decode_moments_f32|decode_moments_kernel<10, true")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    cur, body = None, collections.defaultdict(list)
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            body[cur].append(line)
    dem = subprocess.run(["c++filt"], input="\n".join(body), capture_output=True, text=True).stdout.splitlines()
    for k, d in zip(body, dem):
        names[k] = d
    lines = ["SASS evidence of the hot kernels (cuobjdump -sass %s, sm_100a; tools/sass_summary.py)." % LIB,
             "Per kernel: instruction count and the tcgen05 / TMA / packed-math mnemonics (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,",
             "UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier,",
             "HFMA2 = packed fp16 FMA, FFMA2 / FADD2 / FMUL2 = packed fp32), then the first tensor-core MMA group of the kernel.", ""]
    for k in sorted(body, key=lambda k: names[k]):
        if not HOT.search(names[k]):
            continue
        ops = collections.Counter()
        for l in body[k]:
            m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
            if m:
                ops[m.group(1)] += 1
        lines.append("== " + names[k])
        lines.append("   instructions %d  " % len(body[k]) + "  ".join("%s %d" % (x, ops[x]) for x in KEYS if ops[x]))
        first = next((i for i, l in enumerate(body[k]) if "UTCHMMA" in l), None)
        if first is not None:
            for l in body[k][max(0, first - 1):first + 5]:
                lines.append("      " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.strip()))
        lines.append("")
    open(sys.argv[1], "w").write("\n".join(lines))
    print(len(lines), "lines ->", sys.argv[1])


if __name__ == "__main__":
    main()
