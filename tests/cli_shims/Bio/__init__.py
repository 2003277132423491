"""TEST INFRASTRUCTURE: stand-in for BioPython where it is not installed.  MetaCoAG v1.1.5 uses
``SeqIO.parse(path, "fasta")`` only (feature_utils.py:131,182; graph_utils.py:129;
metacoag:358,1236)."""
from . import SeqIO  # noqa: F401
