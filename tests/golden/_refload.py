"""Import the UNMODIFIED reference modules from /root/reference (build container only).

Used by tests/golden/make_golden.py to generate the committed fixtures.  Nothing
that runs on the GPU box imports this file: /root/reference does not exist there.

The reference needs two packages that are not installed here:
``Bio`` (only ``SeqIO.parse(path, "fasta")`` is used) and ``igraph`` (only a handful
of ``Graph`` methods).  Both are replaced by the minimal stand-ins of tests/cli_shims (the
ones the unmodified ``metacoag`` script runs with in tests/test_cli.py); the reference sources
themselves are imported as they are.
"""
import importlib
import os
import sys

REFERENCE_ROOT = "/root/reference"

_SHIMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cli_shims")
if _SHIMS not in sys.path:
    sys.path.append(_SHIMS)      # after site-packages: real BioPython / igraph win when installed

import igraph as _igraph  # noqa: E402
from Bio import SeqIO as _seqio  # noqa: E402,F401

ShimGraph = _igraph.Graph


def load_reference():
    """Returns the reference's (feature_utils, matching_utils, label_prop_utils)."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    fu = importlib.import_module("metacoag_utils.feature_utils")
    mu = importlib.import_module("metacoag_utils.matching_utils")
    lp = importlib.import_module("metacoag_utils.label_prop_utils")
    assert fu.__file__.startswith(REFERENCE_ROOT), fu.__file__
    return fu, mu, lp
