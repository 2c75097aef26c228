#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: byte-compile the UNMODIFIED reference sources of the hot path, where they lie under
/root/reference, into sourceless ``.pyc`` modules (build artefacts: git-ignored, shipped to the GPU box with the tree
like the package's own ``.so``).  TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/build_ref.py          # needs /root/reference (the build container); a no-op message elsewhere

Compiled: src/neural_decoder/model.py (GRUDecoder) and src/neural_decoder/augmentations.py (GaussianSmoothing), the
only reference files the path `GRUDecoder.forward` needs.  neural_decoder_trainer.py is NOT compiled: it imports hydra /
edit_distance, which this image lacks; its loss / optimiser lines (trainer:139-141, 163-169, 194-218, 242, 251-259) are
restated in oracle/torch_port.py:train_step, which drives either module.  No reference source text enters the repo.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/neural_decoder"
OUT = os.path.join(HERE, "_ref", "neural_decoder")
FILES = ("model", "augmentations")


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle/build_ref.py: {SRC} not present (GPU box?) -- keeping whatever oracle/_ref holds")
        return False
    os.makedirs(OUT, exist_ok=True)
    for f in FILES:
        py_compile.compile(os.path.join(SRC, f + ".py"), cfile=os.path.join(OUT, f + ".pyc"), doraise=True, optimize=0)
    if verbose:
        print(f"oracle/build_ref.py: compiled {', '.join(FILES)} -> {OUT}")
    return True


def load_reference_decoder():
    """The reference's own ``GRUDecoder`` class from oracle/_ref (None when it has not been built)."""
    if not all(os.path.exists(os.path.join(OUT, f + ".pyc")) for f in FILES):
        return None
    root = os.path.join(HERE, "_ref")
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        from neural_decoder.model import GRUDecoder      # sourceless import of the compiled reference module
        return GRUDecoder
    except Exception:
        return None


if __name__ == "__main__":
    build()
