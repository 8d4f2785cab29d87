"""Import the UNMODIFIED reference package straight from /root/reference, in memory.

TEST INFRASTRUCTURE ONLY (never imported by the product package `binf_b200`).

The reference (simeoncarstens/binf) is Python 2.7 and depends on CSB, neither of which
exists in this image.  This module makes `import binf` work under Python 3 *in this
container* without copying a single reference source file into the repository:

  * `oracle/csb_standin/` is put on sys.path (the ~150-line CSB surface the reference
    touches);
  * a meta-path finder reads each `binf/**.py` source from /root/reference, applies the
    three py2->py3 token fixes listed in SURVEY.md 8(c) to the text in memory, and
    compiles it.  Nothing is written to disk.

Token fixes (file:line in /root/reference):
  binf/pdf/posteriors.py:182   `.iteritems()`  -> `.items()`
  binf/__init__.py:141         `.viewkeys()`   -> `.keys()`
  binf/example/samplers.py:18-20  `filter(...)[0]` -> `list(filter(...))[0]`, drop `print prior`

/root/reference does not exist on the GPU box; `available()` says whether this route
can be used.  It is used by oracle/make_golden.py (to generate tests/golden/*.npz) and by
the CPU-only tests that re-check the golden vectors when the reference is present.
"""
import importlib.abc
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get("BINF_REFERENCE_ROOT", "/root/reference")
_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csb_standin")

# modules that are py2-only demos and are not imported by anything on the path
_SKIP = {"binf.pdf.parameters", "binf.pdf.example", "binf.example.plots"}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "binf", "__init__.py"))


def _patch(modname, text):
    if modname == "binf.pdf.posteriors":
        text = text.replace(".iteritems()", ".items()")
    elif modname == "binf":
        text = text.replace(".viewkeys()", ".keys()")
    elif modname == "binf.example.samplers":
        text = text.replace("prior = filter(lambda p: 'precision' in p.variables,\n"
                            "                       self.pdf.priors.values())[0]",
                            "prior = list(filter(lambda p: 'precision' in p.variables,\n"
                            "                       self.pdf.priors.values()))[0]")
        text = text.replace("        print prior\n", "")
    return text


class _RefLoader(importlib.abc.Loader):
    def __init__(self, modname, path, is_pkg):
        self.modname, self.path, self.is_pkg = modname, path, is_pkg

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(self.path, "r") as fh:
            text = fh.read()
        code = compile(_patch(self.modname, text), self.path, "exec")
        exec(code, module.__dict__)


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != "binf" and not fullname.startswith("binf."):
            return None
        if fullname in _SKIP:
            return None
        rel = fullname.split(".")
        base = os.path.join(REFERENCE_ROOT, *rel)
        if os.path.isdir(base) and os.path.isfile(os.path.join(base, "__init__.py")):
            loader = _RefLoader(fullname, os.path.join(base, "__init__.py"), True)
            spec = importlib.util.spec_from_loader(fullname, loader, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        if os.path.isfile(base + ".py"):
            loader = _RefLoader(fullname, base + ".py", False)
            return importlib.util.spec_from_loader(fullname, loader)
        return None


_installed = False


def install():
    """Make `import binf` resolve to the reference.  Raises if it is not present."""
    global _installed
    if not available():
        raise ImportError("reference tree not found at %s" % REFERENCE_ROOT)
    # another package may have been registered under the name `binf` (binf_b200.install_as_binf)
    stale = [k for k, m in sys.modules.items() if (k == "binf" or k.startswith("binf."))
             and not str(getattr(m, "__file__", "")).startswith(REFERENCE_ROOT)]
    for k in stale:
        del sys.modules[k]
    if not _installed:
        if _STANDIN not in sys.path:
            sys.path.insert(0, _STANDIN)
        sys.meta_path.insert(0, _RefFinder())
        _installed = True
    import binf  # noqa: F401
    return sys.modules["binf"]
