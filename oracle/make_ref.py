"""Build oracle/_ref/: a runnable copy of the UNMODIFIED Python reference + its shipped data, so that the GPU box
(which has no /root/reference) can time the reference itself on its host cores (SURVEY 8d i/ii, bench.py --impl
reference -> `reference_python`).  Run in the build container:

    python oracle/make_ref.py

oracle/_ref/ is git-ignored (never part of this repository's history) but travels with the gpurun snapshot, like
tests/golden/_real/.  The copy is byte-identical to /root/reference except for the ONE pandas-3 copy-on-write shim
documented in SURVEY 8c (`.values` -> `.values.copy()` at simulator.py:315; the array mutated there is not used
afterwards) and a no-op `matplotlib` stub package (the module imports it at top level; it is absent here).
TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing in the product package reads oracle/_ref."""
import os
import shutil
import sys

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(REF_ROOT):
        print("no /root/reference here: nothing to do")
        return 1
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for sub in ("simulator", "objects", "config", "preprocessing", "data"):
        shutil.copytree(os.path.join(REF_ROOT, sub), os.path.join(OUT, sub))
    for dirpath, _, files in os.walk(OUT):
        os.chmod(dirpath, 0o755)
        for f in files:
            os.chmod(os.path.join(dirpath, f), 0o644)
    p = os.path.join(OUT, "simulator", "simulator.py")
    src = open(p, encoding="utf-8").read()
    old = "NodeID = self.Node['NodeID'].values\n        for i in range(len(NodeID)):"
    new = "NodeID = self.Node['NodeID'].values.copy()\n        for i in range(len(NodeID)):"
    assert src.count(old) == 1, "reference changed: shim anchor not found"
    open(p, "w", encoding="utf-8").write(src.replace(old, new))
    mp = os.path.join(OUT, "matplotlib")
    os.makedirs(mp)
    stub = "def __getattr__(name):\n    def _noop(*a, **k):\n        return None\n    return _noop\n"
    open(os.path.join(mp, "__init__.py"), "w").write(stub)
    open(os.path.join(mp, "pyplot.py"), "w").write(stub)
    size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(OUT) for f in fs)
    print(f"oracle/_ref ready: {size / 1e6:.1f} MB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
