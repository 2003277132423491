"""TEST INFRASTRUCTURE: run the UNMODIFIED ``metacoag`` script of MetaCoAG v1.1.5 in a subprocess.

The script comes from ``/root/reference`` (build container) or from ``baseline/_ref`` -- the
reference installed once, unmodified, with
``pip install --no-index --no-deps --target baseline/_ref /root/reference`` (git-ignored; it
travels to the GPU box with the snapshot, ``/root/reference`` does not).  ``tests/cli_shims`` on
PYTHONPATH supplies igraph / Bio where they are not installed and, with ``swap=True``, the module
swap of INTEGRATION.md section 1.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "tests", "cli_shims")


def find_cli():
    """(script path, directory that holds metacoag_utils) or None."""
    for script, pkg in (("/root/reference/metacoag", "/root/reference"),
                        (os.path.join(ROOT, "baseline", "_ref", "bin", "metacoag"),
                         os.path.join(ROOT, "baseline", "_ref"))):
        if os.path.isfile(script) and os.path.isdir(os.path.join(pkg, "metacoag_utils")):
            return script, pkg
    return None


def run_cli(args, swap, backend=None, timeout=900):
    found = find_cli()
    if found is None:
        raise FileNotFoundError("no metacoag script (neither /root/reference nor baseline/_ref)")
    script, pkg = found
    env = dict(os.environ)
    # site-packages first: the shims only fill in for packages that are not installed
    env["PYTHONPATH"] = os.pathsep.join([pkg, SHIMS] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env["MCG_REPO_ROOT"] = ROOT
    env["MCG_CLI_SWAP"] = "1" if swap else "0"
    env.pop("MCG_CLI_BACKEND", None)
    if backend:
        env["MCG_CLI_BACKEND"] = backend
    proc = subprocess.run([sys.executable, script] + list(args), env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=timeout)
    return proc.returncode, proc.stdout
