"""
TEST INFRASTRUCTURE — runs the reference's own Python sources, unmodified and in place, under Python 3.

The reference's Python layer (``tredparse/{utils,meta,bam_parser,models,tred}.py`` and ``src/ssw_wrap.py``)
is Python 2 only and depends on pysam, neither of which exists in this image.  This module loads those
files **from where they lie under /root/reference** (nothing is copied into the repository), applies the
small, mechanical Python 2 → 3 transform listed below, and executes them as the package ``reftredparse``:

    ref = refshim.load(libssw="/path/to/libssw-compatible.so")
    ref.models.IntegratedCaller(...).call()        # the reference's own likelihood code
    ref.ssw.Aligner(ref_seq=...).align(read, ...)   # the reference's own ctypes binding, bound to `libssw`
    ref.tred.run((samplekey, bam, repo, ...))       # the reference's own per-sample loop

so that golden vectors (``tests/golden/make_ref_fixtures.py``) and parity tests are anchored on the
reference itself instead of on a restatement of it.

The transform (``py3_source`` + ``_Py2Semantics``) — everything else runs verbatim:

 text level (the files do not parse as Python 3 otherwise)
  T1  ``print >> f, a, b`` / ``print >> f`` / ``print x``   →  ``print(a, b, file=f)`` / ``print(file=f)`` / ``print(x)``
      (bam_parser.py:145-149, tred.py:307,309,366,369,448,534)
  T2  ``xrange`` → ``range``; ``d.iteritems()`` → ``d.items()``; ``x.next()`` → ``next(x)``
      (models.py:71-72, bam_parser.py:91,345, tred.py:420, ssw_wrap.py:353, utils.py:205)
  T3  ``string.maketrans`` → ``str.maketrans`` (bam_parser.py:32); ``unicode`` → ``str`` (utils.py:209)
  T4  ``gzip.open(f, "w")`` → ``"wt"`` (tred.py:365: the VCF is written with ``print``)
 AST level (semantics Python 3 changed silently)
  A1  every ``a / b`` becomes ``_py2div(a, b)``: floor division when both operands are integers (Python /
      numpy ints or integer arrays), true division otherwise — exactly Python 2's classic division.
      This covers models.py:156,161,288-289,312,353-362,410, meta.py:119, bam_parser.py:133,
      ssw_wrap.py:199 without hand-picking the sites.
  A2  ``range(...)`` returns a list, as in Python 2 (models.py:252-253 concatenates it to a list).
  A3  implicit relative imports (``from bam_parser import …``, ``from utils import …``, ``from ssw import
      Aligner``: models.py:27-28, bam_parser.py:25-26) and ``import pysam`` are redirected to the modules
      of this package / to ``oracle.pysam_stub`` (a pysam look-alike over the repo's pure-Python BAM reader).

``ssw_wrap.py`` finds ``libssw.so`` next to its ``__file__`` (ssw_wrap.py:20-31).  ``load(libssw=…)`` points
that ``__file__`` into a scratch directory holding a symlink named ``libssw.so`` to the requested library —
``oracle/_ref/libssw_ref.so`` (the reference's ``ssw.c`` compiled as is) or ``libtredsw.so`` (the GPU
library: INTEGRATION.md level 0, an unmodified ``Aligner`` on the drop-in).

/root/reference does not exist on the GPU box.  ``build()`` (run here by ``__graft_entry__.build``)
therefore also writes the *compiled code objects* of the transformed modules to ``oracle/_ref/refpy/*.bin`` and
the data tables they read (catalogue, alt regions, step / stutter model) to ``oracle/_ref/refpy/data/`` —
git-ignored build artefacts that travel with the snapshot like ``libssw_ref.so``; ``load()`` falls back to them
when the source tree is absent, so ``bench.py``'s reference arm runs the reference's own code there too.
"""
import ast
import builtins
import marshal
import os
import re
import shutil
import sys
import tempfile
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("TREDPARSE_REFERENCE", "/root/reference")
CACHE = os.path.join(HERE, "_ref", "refpy")
PKG = "reftredparse"
# module name in the package -> path below the reference root
MODULES = {
    "utils": "tredparse/utils.py",
    "ssw": "src/ssw_wrap.py",
    "meta": "tredparse/meta.py",
    "bam_parser": "tredparse/bam_parser.py",
    "models": "tredparse/models.py",
    "tred": "tredparse/tred.py",
}
_SIBLINGS = set(MODULES)


def available():
    return os.path.isdir(os.path.join(REFERENCE, "tredparse"))


def cached(name="ssw"):
    return os.path.exists(os.path.join(CACHE, name + ".bin"))


def usable():
    """The whole reference package can be loaded: the source tree is mounted, or build() left the code objects and
    the reference's data tables under oracle/_ref/refpy/."""
    return available() or (all(cached(m) for m in MODULES) and os.path.isdir(os.path.join(CACHE, "data")))


# ------------------------------------------------------------------------------------------------
# Python-2 semantics injected into every module
# ------------------------------------------------------------------------------------------------
def _is_int(x):
    if isinstance(x, (bool, int, np.integer)):
        return True
    return isinstance(x, np.ndarray) and x.dtype.kind in "iub"


def _py2div(a, b):
    """Python 2 classic division."""
    if _is_int(a) and _is_int(b):
        return a // b
    return a / b


def _py2range(*args):
    return list(builtins.range(*args))


# ------------------------------------------------------------------------------------------------
# the transform
# ------------------------------------------------------------------------------------------------
_TEXT_RULES = [
    # T1
    (re.compile(r"^([ \t]*)print[ \t]*>>[ \t]*([^,\n]+?),[ \t]*([^\n]*)$", re.M), r"\1print(\3, file=\2)"),
    (re.compile(r"^([ \t]*)print[ \t]*>>[ \t]*([^,\n]+?)[ \t]*$", re.M), r"\1print(file=\2)"),
    (re.compile(r"^([ \t]*)print[ \t]+(?!\()([^\n]+)$", re.M), r"\1print(\2)"),
    # T2
    (re.compile(r"\bxrange\b"), "range"),
    (re.compile(r"\.iteritems\(\)"), ".items()"),
    (re.compile(r"\b([A-Za-z_][A-Za-z_0-9]*)\.next\(\)"), r"next(\1)"),
    # T3
    (re.compile(r"\bstring\.maketrans\b"), "str.maketrans"),
    (re.compile(r"\bunicode\b"), "str"),
    # T4
    (re.compile(r"gzip\.open\(([^,\n]+),[ \t]*\"w\"\)"), r'gzip.open(\1, "wt")'),
]


def _join_print_continuations(text):
    """A print statement continued with a backslash becomes one line (blank lines keep the numbering)."""
    out, lines, i = [], text.split("\n"), 0
    while i < len(lines):
        line = lines[i]
        pad = 0
        if re.match(r"[ \t]*print\b", line):
            while line.endswith("\\") and i + 1 < len(lines):
                i += 1
                pad += 1
                line = line[:-1] + " " + lines[i].strip()
        out.append(line)
        out.extend([""] * pad)
        i += 1
    return "\n".join(out)


def py3_source(text):
    text = _join_print_continuations(text)
    for rx, rep in _TEXT_RULES:
        text = rx.sub(rep, text)
    return text


class _Py2Semantics(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(
                ast.Call(func=ast.Name(id="_py2div", ctx=ast.Load()), args=[node.left, node.right], keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div) and isinstance(node.target, ast.Name):
            load = ast.Name(id=node.target.id, ctx=ast.Load())
            return ast.copy_location(ast.Assign(
                targets=[node.target],
                value=ast.Call(func=ast.Name(id="_py2div", ctx=ast.Load()), args=[load, node.value], keywords=[])), node)
        return node

    def visit_ImportFrom(self, node):
        if node.level == 0 and node.module in _SIBLINGS:
            node.module = PKG + "." + node.module
        return node

    def visit_Import(self, node):
        if len(node.names) == 1 and node.names[0].name == "pysam":
            return ast.copy_location(ast.ImportFrom(
                module="oracle", names=[ast.alias(name="pysam_stub", asname="pysam")], level=0), node)
        return node


def compile_module(name):
    path = os.path.join(REFERENCE, MODULES[name])
    with open(path) as fp, warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)       # a '\ ' in a reference docstring
        tree = ast.parse(py3_source(fp.read()), filename=path)
    tree = ast.fix_missing_locations(_Py2Semantics().visit(tree))
    return compile(tree, path, "exec")


def build():
    """Compile every module from the reference tree into oracle/_ref/refpy/*.bin (code objects)."""
    if not available():
        return False
    os.makedirs(CACHE, exist_ok=True)
    for name in MODULES:
        with open(os.path.join(CACHE, name + ".bin"), "wb") as fp:
            marshal.dump(compile_module(name), fp)
    with open(os.path.join(REFERENCE, "tredparse", "__init__.py")) as fp:
        ver = re.search(r'__version__\s*=\s*"([^"]+)"', fp.read()).group(1)
    with open(os.path.join(CACHE, "VERSION"), "w") as fp:
        fp.write(ver + "\n")
    # the tables the reference's modules read at run time (catalogue, alt regions, step / stutter model, chrY regions)
    os.makedirs(os.path.join(CACHE, "data"), exist_ok=True)
    src = os.path.join(REFERENCE, "tredparse", "data")
    for name in os.listdir(src):
        if name.endswith((".csv", ".stepmodel", ".stuttermodel", ".gc")):
            shutil.copyfile(os.path.join(src, name), os.path.join(CACHE, "data", name))
    return True


def _code(name):
    if available():
        return compile_module(name)
    with open(os.path.join(CACHE, name + ".bin"), "rb") as fp:
        return marshal.load(fp)


# ------------------------------------------------------------------------------------------------
# loader
# ------------------------------------------------------------------------------------------------
class Reference(types.SimpleNamespace):
    pass


_loaded = {}


def load(libssw=None, modules=None):
    """Load the reference package bound to ``libssw`` (default: oracle/_ref/libssw_ref.so).
    Returns a namespace with one attribute per module.  Each distinct ``libssw`` gets its own package
    instance (``reftredparse``, ``reftredparse_1`` …) because ssw_wrap binds the library at class-body time."""
    global PKG
    if libssw is None:
        libssw = os.path.join(HERE, "_ref", "libssw_ref.so")
    libssw = os.path.abspath(libssw)
    if not os.path.exists(libssw):
        raise IOError("libssw-compatible library not found: {}".format(libssw))
    if modules is None:
        modules = list(MODULES) if usable() else ["ssw"]
    key = (libssw, tuple(modules))
    if key in _loaded:
        return _loaded[key]

    # the full package keeps the name `reftredparse` (the cached code objects import from it by that name); an
    # ssw-only instance bound to another library (the drop-in tests) imports nothing and takes a numbered name
    full = any(m != "ssw" for m in modules)
    pkgname = "reftredparse" if full and "reftredparse" not in sys.modules else "reftredparse_{}".format(len(_loaded) + 1)
    scratch = tempfile.mkdtemp(prefix="refssw_")
    os.symlink(libssw, os.path.join(scratch, "libssw.so"))

    pkg = types.ModuleType(pkgname)
    pkg.__path__ = []
    pkg.__package__ = pkgname
    if available():
        with open(os.path.join(REFERENCE, "tredparse", "__init__.py")) as fp:
            pkg.__version__ = re.search(r'__version__\s*=\s*"([^"]+)"', fp.read()).group(1)
    else:
        with open(os.path.join(CACHE, "VERSION")) as fp:
            pkg.__version__ = fp.read().strip()
    sys.modules[pkgname] = pkg

    ns = Reference(package=pkg, libssw=libssw)
    saved = PKG
    PKG = pkgname  # consulted by _Py2Semantics.visit_ImportFrom while compiling from source
    try:
        for name in modules:
            mod = types.ModuleType(pkgname + "." + name)
            mod.__package__ = pkgname
            if name == "ssw":
                mod.__file__ = os.path.join(scratch, "ssw_wrap.py")
            elif available():
                mod.__file__ = os.path.join(REFERENCE, MODULES[name])
            else:
                mod.__file__ = os.path.join(CACHE, name + ".py")       # utils.datafile() -> oracle/_ref/refpy/data
            mod.__dict__.update(_py2div=_py2div, range=_py2range)
            sys.modules[mod.__name__] = mod
            code = _code(name)
            if pkgname != "reftredparse" and name != "ssw" and not available():
                raise RuntimeError("cached code objects import from `reftredparse` only")
            exec(code, mod.__dict__)
            setattr(pkg, name, mod)
            setattr(ns, name, mod)
    finally:
        PKG = saved
    _loaded[key] = ns
    return ns


if __name__ == "__main__":
    print("built" if build() else "reference tree not present; nothing built")
