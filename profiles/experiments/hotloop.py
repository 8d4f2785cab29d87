"""Print the non-FP instructions of the chromatin kernel's non-energy stage loop from a built library.
usage: python profiles/experiments/hotloop.py [lib.so] [kernel mangled-name substring]"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "binf_b200/libbinf_b200.so"
kern = sys.argv[2] if len(sys.argv) > 2 else "chrom_kernelILi2ELi2"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins, on = [], False
for l in out.split("\n"):
    if "Function :" in l:
        on = kern in l
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
fp = re.compile(r"(@!?U?P\d\s+)?(FFMA2|FADD2|FMUL2|MUFU|FFMA|FADD|FMUL)\b")
rel = [i for i, (a, t) in enumerate(ins) if "ATOMS" in t]
tw = [i for i, (a, t) in enumerate(ins) if "TRYWAIT" in t]
end = rel[1]
begin = max(i for i in tw if i < rel[1] - 1500) if len(sys.argv) <= 3 else int(sys.argv[3], 16)
# extend to the loop's back edge
for i in range(end, len(ins)):
    m = re.search(r"BRA(?:\.U)?\s+(?:U?P\d, )?(0x[0-9a-f]+)", ins[i][1])
    if m and int(m.group(1), 16) <= ins[begin][0]:
        end = i
        break
n = 0
for a, t in ins[begin - 12:end + 1]:
    if not fp.match(t):
        print("%05x %s" % (a, t))
        n += 1
print("non-FP instructions listed:", n, " total in range:", end - begin + 13)
