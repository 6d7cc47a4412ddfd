"""The C-ABI library loads and exports exactly the symbols include/b200ssl.h declares (no GPU needed)."""
import ctypes
import os
import re

from cv_ssl_mis_b200 import _lib


def header_functions():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    declared = header_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_lib.SIGNATURES) == declared


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in b200ssl.h but not exported"
    assert _lib.load().b200_abi_version() == _lib.ABI_VERSION


def test_no_torch_types_in_the_abi():
    src = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)      # declarations only
    assert "torch" not in src.lower() and "at::" not in src and "Tensor" not in src


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(_lib.__file__))
    pkg = os.path.join(root, "cv_ssl_mis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text or "import oracle" not in text and "from oracle" not in text, f
