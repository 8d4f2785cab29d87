"""SASS listings of the hot kernels from the in-tree objects (binf_b200/csrc/_obj/*.o) -> profiles/sass/.
usage: python profiles/experiments/dump_sass.py   (after python -m binf_b200.build)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OBJ = os.path.join(ROOT, "binf_b200", "csrc", "_obj")
OUT = os.path.join(ROOT, "profiles", "sass")
KERNELS = [
    ("chromatin.o", "chrom_kernel<2, 2, false, 4, false>", "r2_chrom_kernel_2_2_0_4_0.sass",
     "the benchmarked chromatin kernel (n = 1000: 2 roles per chain, 2 steps per role and stage, 4-stage ring, no excluded volume)"),
    ("chromatin.o", "chrom_kernel<16, 1, false, 3, false>", "r2_chrom_kernel_16_1_0_3_0.sass",
     "the n = 5000 chromatin kernel (16 roles per chain, 3-stage ring)"),
    ("poly.o", "poly_hmc_kernel<4, 4, 2, true>", "r2_poly_hmc_kernel_4_4_2_ur.sass",
     "the polynomial kernel, uniform-row mapping (K = 4, 4 warps per set, chain pairs)"),
]


def dump(obj, pattern, out, title):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0]
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if pattern not in dem:
            continue
        lines, hist = [], collections.Counter()
        for l in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                lines.append("/*%s*/  %s" % (m.group(1), m.group(2).strip()))
                t = m.group(2).strip().split()
                hist[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += 1
        with open(os.path.join(OUT, out), "w") as fh:
            fh.write("// %s\n// cuobjdump -sass of %s (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo); "
                     "encodings stripped\n" % (title, dem[:160]))
            fh.write("// %d instructions; opcode histogram: %s\n" % (
                len(lines), ", ".join("%s %d" % kv for kv in hist.most_common(24))))
            fh.write("\n".join(lines) + "\n")
        print(out, len(lines))
        return
    print("not found:", pattern)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for k in KERNELS:
        dump(*k)
