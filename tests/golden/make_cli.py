#!/usr/bin/env python3
"""Generate tests/golden/cli.json with the UNMODIFIED reference CLI (build container).

For every flavour of tests/cli_data.py the synthetic assembly is written to a temporary directory
and ``/root/reference/metacoag`` runs on it with the reference's own metacoag_utils (igraph / Bio
from tests/cli_shims, nothing of this repo on the path).  Stored: the SHA-256 of the inputs, the
contig -> bin assignment of ``contig_to_bin.tsv`` and the sorted record counts of the per-bin
FASTA files (what the reference's own acceptance test asserts, tests/test_metacoag.py:72-79).
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import cli_data  # noqa: E402
import cli_run  # noqa: E402


def main():
    found = cli_run.find_cli()
    assert found and found[0].startswith("/root/reference"), "run this in the build container"
    golden = {}
    for flavour in cli_data.FLAVOURS:
        with tempfile.TemporaryDirectory() as tmp:
            files = cli_data.write_dataset(os.path.join(tmp, "data"), flavour)
            out = os.path.join(tmp, "out")
            os.makedirs(out)
            rc, log = cli_run.run_cli(cli_data.cli_args(flavour, files, out), swap=False)
            assert rc == 0, log
            assign, counts = cli_data.read_result(out)
            golden[flavour] = {"sha256": files["sha256"], "n_contigs": files["n_contigs"],
                               "assign": assign, "bin_record_counts": counts}
            print(flavour, "bins:", counts, "binned contigs:", len(assign))
    with open(os.path.join(HERE, "cli.json"), "w") as fh:
        json.dump(golden, fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
