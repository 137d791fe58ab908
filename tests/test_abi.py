"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import os
import re

from ic_gan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "icgan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|const char\*)\s+(icgan_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/icgan_b200.h but missing from libicgan_b200.so"
    assert lib.icgan_version() >= 100
    assert lib.icgan_last_error() is not None


def test_python_binding_covers_the_header():
    names = set(_header_functions()) - {"icgan_last_error", "icgan_version"}
    assert names == set(_lib.SIGNATURES), (names ^ set(_lib.SIGNATURES))


def test_argument_counts_match_header():
    text = open(os.path.join(ROOT, "include", "icgan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\bint\s+" + name + r"\s*\((.*?)\);", text, flags=re.S)
        assert m, name
        n_args = len([a for a in m.group(1).split(",") if a.strip()])
        assert n_args == len(argtypes), f"{name}: header has {n_args} args, binding {len(argtypes)}"
