"""Build libbinf_b200.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

    python -m binf_b200.build            # incremental
    python -m binf_b200.build --force

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libbinf_b200.so")
SOURCES = ["capi.cu", "poly.cu", "chromatin.cu", "misc.cu", "sink.cu", "rwmc.cu", "generic.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(OBJ, "generic_src.inc")] + [
        os.path.join(os.path.dirname(HERE), "include", "binf_b200.h")]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and not _stale(obj, [path] + _deps()):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(obj + ".ptxas.log", "w") as fh:
        fh.write(r.stderr)
    return obj, r.stderr if verbose else ""


def _generic_source_bundle():
    """csrc/_obj/generic_src.inc: the NVRTC source of the generic-model kernels as two C++ string
    constants -- HEAD (prelude, common.cuh, the kernel-argument structs of internal.h) and TAIL
    (generic_kernel.cuh); the user's device code is spliced in between at run time."""
    import re
    common = open(os.path.join(CSRC, "common.cuh")).read()
    common = "\n".join(l for l in common.split("\n") if not l.startswith("#include") and not l.startswith("#pragma once"))
    internal = open(os.path.join(CSRC, "internal.h")).read()
    args = internal[internal.index("// BEGIN_KERNEL_ARGS"):internal.index("// END_KERNEL_ARGS")]
    header = open(os.path.join(os.path.dirname(HERE), "include", "binf_b200.h")).read()
    defines = "\n".join(re.findall(r"^#define BINFB_(?:FLAG|GIBBS)_\w+ .*?$", header, flags=re.M))
    defines = re.sub(r"/\*.*", "", defines)
    prelude = ("typedef unsigned char uint8_t;\ntypedef int int32_t;\ntypedef unsigned int uint32_t;\n"
               "typedef long long int64_t;\ntypedef unsigned long long uint64_t;\n" + defines + "\n")
    pack = open(os.path.join(CSRC, "generic_pack.cuh")).read()
    head = prelude + common + "\nnamespace binfb {\n" + args + "\n}\nusing namespace binfb;\n" + pack
    tail = open(os.path.join(CSRC, "generic_kernel.cuh")).read()

    def lit(name, text):
        chunks = [text[i:i + 12000] for i in range(0, len(text), 12000)]
        body = "\n".join('R"BINFBSRC(' + c + ')BINFBSRC"' for c in chunks)
        return "static const char *const %s =\n%s;\n" % (name, body)
    out = lit("GEN_SRC_HEAD", head) + lit("GEN_SRC_TAIL", tail)
    path = os.path.join(OBJ, "generic_src.inc")
    if not os.path.exists(path) or open(path).read() != out:
        with open(path, "w") as fh:
            fh.write(out)
    return path


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    _generic_source_bundle()
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    objs = [o for o, _ in results]
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
