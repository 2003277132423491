"""TEST INFRASTRUCTURE, picked up by the interpreter that runs the UNMODIFIED ``metacoag``
script when this directory is on PYTHONPATH (tests/test_cli.py, tests/golden/make_cli.py).

MCG_CLI_SWAP=1   the module swap of INTEGRATION.md section 1: ``metacoag_utils.feature_utils``,
                 ``.matching_utils`` and ``.label_prop_utils`` resolve to the metacoag_b200
                 drop-ins (and graph_utils' two component helpers to theirs); everything else of
                 the script and of metacoag_utils stays the reference's.
MCG_CLI_BACKEND=oracle   (CPU containers only) the drop-ins compute through the pinned CPU
                 oracle instead of libmcg_b200: checks the host layer of the drop-in against
                 the real CLI where there is no GPU.  The product has no such switch.
"""
import os
import sys

if os.environ.get("MCG_CLI_SWAP") == "1":
    root = os.environ["MCG_REPO_ROOT"]
    if root not in sys.path:
        sys.path.insert(0, root)
    import metacoag_utils
    import metacoag_utils.graph_utils as ref_graph_utils
    import metacoag_b200.feature_utils
    import metacoag_b200.graph_utils
    import metacoag_b200.label_prop_utils
    import metacoag_b200.matching_utils

    for name in ("feature_utils", "matching_utils", "label_prop_utils"):
        mod = sys.modules["metacoag_b200." + name]
        sys.modules["metacoag_utils." + name] = mod
        setattr(metacoag_utils, name, mod)
    ref_graph_utils.get_isolated = metacoag_b200.graph_utils.get_isolated
    ref_graph_utils.get_non_isolated = metacoag_b200.graph_utils.get_non_isolated

    if os.environ.get("MCG_CLI_BACKEND") == "oracle":
        sys.path.insert(0, os.path.join(root, "tests"))
        from metacoag_b200 import _backend
        from oracle_backend import OracleBackend
        _backend.set_backend(OracleBackend())
    print("[mcg] module swap active: metacoag_utils.{feature,matching,label_prop}_utils -> "
          "metacoag_b200 (backend: %s)" % (os.environ.get("MCG_CLI_BACKEND") or "libmcg_b200"),
          file=sys.stderr)
