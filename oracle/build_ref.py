"""Build recipe for oracle/_ref: the reference's own training engines, byte-compiled where they lie.  TEST INFRASTRUCTURE.

    python oracle/build_ref.py            # build container only (needs /root/reference); writes oracle/_ref/*.bin

The reference is pure Python, so "compiling it from its own source files" (the only way oracle/_ref may be produced) is
``py_compile``: /root/reference/{bestengine,ViT_engine,ViT_model}.py -> sourceless CPython 3.12 bytecode.  No reference
source text enters the repository: oracle/_ref/ is git-ignored, holds binaries only and travels to the GPU box with the
snapshot (same image, same interpreter), where tests/test_gpu_reference_engine.py imports the bytecode and runs the
reference's OWN ``train_model`` / ``validate_model`` / ``test_model`` / ``visualize_sample_images`` on the drop-in loaders --
the consumer contract of SURVEY.md 8(a20) executed by the consumer itself.  The product never imports this.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
MODULES = ("bestengine", "ViT_engine", "ViT_model")
EXT = ".bin"        # CPython bytecode; not named .pyc because snapshot tools commonly drop *.pyc


def build(quiet: bool = False) -> list:
    """Compile the engines when /root/reference is present; returns the files written ([] elsewhere)."""
    if not os.path.isdir(REF):
        return []
    os.makedirs(OUT, exist_ok=True)
    done = []
    for m in MODULES:
        src = os.path.join(REF, m + ".py")
        if os.path.exists(src):
            dst = os.path.join(OUT, m + EXT)
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/{m}.py", doraise=True, optimize=0)
            done.append(dst)
    with open(os.path.join(OUT, "PYTHON_TAG"), "w") as f:
        f.write(sys.implementation.cache_tag + "\n")
    if not quiet:
        print("oracle/_ref:", ", ".join(os.path.basename(d) for d in done))
    return done


def available(module: str) -> bool:
    tag = os.path.join(OUT, "PYTHON_TAG")
    return os.path.exists(os.path.join(OUT, module + EXT)) and os.path.exists(tag) and \
        open(tag).read().strip() == sys.implementation.cache_tag


def load(module: str, name: str | None = None):
    """Import a byte-compiled reference module (executes its top level: callers install stand-ins for absent packages
    first; the reference's ``main()`` calls are under ``if __name__ == '__main__'``)."""
    path = os.path.join(OUT, module + EXT)
    loader = importlib.machinery.SourcelessFileLoader(name or module, path)
    spec = importlib.util.spec_from_loader(name or module, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build()
