"""Compatibility shim: the verbatim-reference harness lives in ``oracle/ref_verbatim.py`` (it now also serves the
GPU-box tests and bench.py's reference arm through the bytecode ``oracle/build_ref.py`` compiles into
``oracle/_ref/``).  ``tools/make_golden.py`` keeps importing it under this name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_verbatim import install_stubs, load_reference, run_curvature, run_dk_dn  # noqa: E402,F401
