#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: byte-compile the UNMODIFIED reference sources of the hot path, where they lie under
/root/reference, into marshalled code objects (build artefacts: git-ignored, shipped to the GPU box with the tree
like the package's own ``.so``).  TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/build_ref.py          # needs /root/reference (the build container); a no-op message elsewhere

Compiled: src/neural_decoder/model.py (GRUDecoder) and src/neural_decoder/augmentations.py (GaussianSmoothing), the
only reference files the path `GRUDecoder.forward` needs.  neural_decoder_trainer.py is NOT compiled: it imports hydra /
edit_distance, which this image lacks; its loss / optimiser lines (trainer:139-141, 163-169, 194-218, 242, 251-259) are
restated in oracle/torch_port.py:train_step, which drives either module.  No reference source text enters the repo.
"""
from __future__ import annotations

import importlib.util
import marshal
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/neural_decoder"
OUT = os.path.join(HERE, "_ref", "neural_decoder")
FILES = ("model", "augmentations")
MAGIC = b"NSDREF1" + importlib.util.MAGIC_NUMBER          # interpreter-specific bytecode: refuse another version's


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle/build_ref.py: {SRC} not present (GPU box?) -- keeping whatever oracle/_ref holds")
        return False
    os.makedirs(OUT, exist_ok=True)
    for f in FILES:
        # compiled code object of the unmodified source file, marshalled (".bin": sync tools tend to drop *.pyc)
        with open(os.path.join(SRC, f + ".py"), "rb") as fh:
            code = compile(fh.read(), f"/root/reference/src/neural_decoder/{f}.py", "exec", dont_inherit=True, optimize=0)
        with open(os.path.join(OUT, f + ".bin"), "wb") as fh:
            fh.write(MAGIC + marshal.dumps(code))
    if verbose:
        print(f"oracle/build_ref.py: compiled {', '.join(FILES)} -> {OUT}")
    return True


_loaded = None


def load_reference_decoder():
    """The reference's own ``GRUDecoder`` class, executed from the code objects in oracle/_ref (None when they have not
    been built or were built by another interpreter version)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    try:
        codes = {}
        for f in FILES:
            with open(os.path.join(OUT, f + ".bin"), "rb") as fh:
                raw = fh.read()
            if raw[:len(MAGIC)] != MAGIC:
                return None
            codes[f] = marshal.loads(raw[len(MAGIC):])
        pkg = types.ModuleType("neural_decoder")
        pkg.__path__ = [OUT]
        sys.modules.setdefault("neural_decoder", pkg)
        for f in ("augmentations", "model"):            # model does `from .augmentations import GaussianSmoothing`
            name = "neural_decoder." + f
            if name not in sys.modules:
                mod = types.ModuleType(name)
                mod.__package__ = "neural_decoder"
                mod.__file__ = codes[f].co_filename
                sys.modules[name] = mod
                exec(codes[f], mod.__dict__)
                setattr(sys.modules["neural_decoder"], f, mod)
        _loaded = sys.modules["neural_decoder.model"].GRUDecoder
        return _loaded
    except Exception as e:                               # noqa: BLE001 -- a missing / stale artefact means "use the port"
        print(f"oracle/build_ref.py: oracle/_ref unusable ({type(e).__name__}: {e}); falling back to the port", file=sys.stderr)
        return None


if __name__ == "__main__":
    build()
