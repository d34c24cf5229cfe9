"""pytest plugin (test infrastructure; used only where /root/reference exists): wherever the REFERENCE's own test files
import the reference's node module (`src.LanPaint.nodes`, or `LanPaint.src.LanPaint.nodes` when the checkout is
imported as the ComfyUI custom-node package), hand them `lanpaint_b200/comfy_nodes.py` instead -- executed under
that name, inside whatever ComfyUI stubs the reference test installed.  Every other `src.LanPaint.*` module still
resolves to the reference's files.  See tests/test_reference_suite_on_b200_nodes.py."""
import importlib.abc
import importlib.util
import os
import sys
import types

ROOT = os.environ["B200_ROOT"]
REF = os.environ.get("LANPAINT_REFERENCE_ROOT", "/root/reference")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TARGET = os.path.join(ROOT, "lanpaint_b200", "comfy_nodes.py")
NODE_MODULES = ("src.LanPaint.nodes", "LanPaint.src.LanPaint.nodes")
PACKAGES = {"src": os.path.join(REF, "src"), "src.LanPaint": os.path.join(REF, "src", "LanPaint"),
            "LanPaint": None, "LanPaint.src": None, "LanPaint.src.LanPaint": None}
LOADED = []


class _NodesLoader(importlib.abc.Loader):
    def create_module(self, spec):
        return None

    def exec_module(self, module):
        module.__package__ = "lanpaint_b200"      # `from .engine import ...` resolves against the real package
        module.__file__ = TARGET
        with open(TARGET) as f:
            exec(compile(f.read(), TARGET, "exec"), module.__dict__)
        LOADED.append(module.__name__)


class _PackageLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        path = PACKAGES[spec.name]
        m.__path__ = [path] if path else []
        return m

    def exec_module(self, module):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        if name in NODE_MODULES:
            return importlib.util.spec_from_loader(name, _NodesLoader(), origin=TARGET)
        if name in PACKAGES:
            return importlib.util.spec_from_loader(name, _PackageLoader(), is_package=True)
        return None


sys.meta_path.insert(0, _Finder())


def pytest_terminal_summary(terminalreporter):
    terminalreporter.write_line(f"b200-alias: node module loaded from {TARGET} x{len(LOADED)}")
