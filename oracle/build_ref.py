"""Recipe for oracle/_ref: the reference's OWN hot path, compiled where it lies.

    python oracle/build_ref.py            (also run by `make -C oracle` and __graft_entry__.build())

The reference's hot path is four pure-Python modules: the engine (src/LanPaint/lanpaint.py, earlystop.py, types.py;
they import only torch and each other) and the node layer above it (nodes.py; it imports ComfyUI, for which
`minicomfy` stands in on the GPU box exactly as it does for lanpaint_b200's nodes).  This compiles them -- from
/root/reference, which exists only in the build container -- to CPython bytecode under oracle/_ref/LanPaint/*.pyc.
Nothing is copied into the repository: oracle/_ref/ is git-ignored (built artefact, like the .so files) but travels
to the GPU box, where `bench.py --impl reference` and the `cpu_baseline` leg import it as the sourceless package
`LanPaint` and time the unmodified reference (`cpu_baseline.kind == "reference"`): the same
`LanPaint_KSampler.sample(...)` call the GPU arm makes, LATENT dict in, LATENT dict out.  Without oracle/_ref both
fall back to the oracle port (`kind == "port"`).

TEST / BENCH INFRASTRUCTURE ONLY: nothing under lanpaint_b200/ imports it.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LANPAINT_REFERENCE", "/root/reference/src/LanPaint")
DST = os.path.join(HERE, "_ref", "LanPaint")
MODULES = ("lanpaint", "earlystop", "types", "nodes")


def build() -> bool:
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    for name in MODULES:
        py_compile.compile(os.path.join(SRC, name + ".py"), cfile=os.path.join(DST, name + ".pyc"), doraise=True)
    with open(os.path.join(DST, "BUILD_INFO"), "w") as f:
        f.write(f"compiled from {SRC} by oracle/build_ref.py with python {sys.version.split()[0]}\n")
    return True


def load():
    """-> the reference's LanPaint class, or None when oracle/_ref has not been built (or does not fit this
    interpreter)."""
    root = os.path.join(HERE, "_ref")
    if not os.path.exists(os.path.join(DST, "lanpaint.pyc")):
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        from LanPaint.lanpaint import LanPaint  # noqa: sourceless import
        return LanPaint
    except Exception:
        return None


def load_nodes():
    """-> the reference's node module (`LanPaint.nodes`: LanPaint_KSampler & co.), imported over minicomfy, or None
    when oracle/_ref has not been built / does not fit this interpreter.  minicomfy must be importable."""
    root = os.path.join(HERE, "_ref")
    if not os.path.exists(os.path.join(DST, "nodes.pyc")):
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        import minicomfy
        minicomfy.install()
        import importlib
        return importlib.import_module("LanPaint.nodes")
    except Exception:
        return None


if __name__ == "__main__":
    print("oracle/_ref built" if build() else f"{SRC} not present: oracle/_ref left as it is")
