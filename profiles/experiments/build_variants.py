"""Build experiment variants of libbinf_b200.so into build/variants/ (git-ignored, travels to the GPU box).

    python profiles/experiments/build_variants.py name=-DFOO=1,-DBAR=2 other= ...

Each variant recompiles chromatin.cu with the extra nvcc flags and links it with the regular objects of
the other translation units.  `head=<git rev>` builds chromatin.cu/pair_block.cuh as of that revision.
Used with profiles/experiments/run_variants.sh for A/B timing on one box.
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from binf_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "build", "variants")


def main():
    os.makedirs(OUT, exist_ok=True)
    B.build()
    others = [os.path.join(B.OBJ, s.replace(".cu", ".o")) for s in B.SOURCES if s != "chromatin.cu"]
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        src_dir = B.CSRC
        tmp = None
        extra = [f for f in flags.split(",") if f]
        if name.startswith("head") and extra and not extra[0].startswith("-"):
            rev = extra.pop(0)
            tmp = tempfile.mkdtemp()
            for f in os.listdir(B.CSRC):
                if f.endswith((".cu", ".cuh", ".h")):
                    blob = subprocess.run(["git", "show", "%s:binf_b200/csrc/%s" % (rev, f)], cwd=ROOT,
                                          capture_output=True)
                    if blob.returncode == 0:
                        open(os.path.join(tmp, f), "wb").write(blob.stdout)
            src_dir = tmp
        obj = os.path.join(OUT, "chromatin_%s.o" % name)
        cmd = [B.NVCC] + B.FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", B.CSRC] + extra + [
            "-c", os.path.join(src_dir, "chromatin.cu"), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout, r.stderr)
            raise SystemExit("variant %s failed" % name)
        regs = [l for l in r.stderr.split("\n") if "Used" in l or "spill" in l]
        lib = os.path.join(OUT, "lib_%s.so" % name)
        subprocess.run([B.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, obj] + others,
                       check=True)
        os.remove(obj)
        print(name, extra, "|", " ; ".join(x.strip() for x in regs[-2:]))


if __name__ == "__main__":
    main()
