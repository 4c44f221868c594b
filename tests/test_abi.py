"""CPU: the C-ABI library loads without a GPU, exports every symbol include/swegl_b200.h declares, the ctypes
mirrors have the C layout, and the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from swegl_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swegl_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swegl_b200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/swegl_b200.h but not exported"
    bound = {n for n, _, _ in _abi.SYMBOLS}
    assert set(names) == bound, f"ctypes bindings and header differ: {set(names) ^ bound}"
    assert lib.swegl_b200_abi_version() == _abi.ABI_VERSION == 4


def test_struct_layouts_match_the_header():
    structs = {"swegl_b200_primitive": _abi.Primitive, "swegl_b200_material": _abi.Material,
               "swegl_b200_texture": _abi.Texture, "swegl_b200_scene_desc": _abi.SceneDesc,
               "swegl_b200_frame_desc": _abi.FrameDesc, "swegl_b200_viewport_desc": _abi.ViewportDesc,
               "swegl_b200_stats": _abi.Stats}
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "swegl_b200.h"\nint main(void){\n'
    for cname, st in structs.items():
        prog += f'printf("{cname} %zu\\n", sizeof({cname}));\n'
        for f, _ in st._fields_:
            prog += f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));\n'
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")], text=True)
    got = dict(line.split() for line in out.strip().splitlines())
    for cname, st in structs.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for f, _ in st._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(st, f).offset, f"{cname}.{f}"


def test_no_cpu_fallback():
    """without a CUDA device the product refuses to run instead of silently using a CPU path"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from swegl_b200.renderer import Renderer, SweglB200Error
    with pytest.raises(SweglB200Error):
        Renderer(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under swegl_b200/ may import or load it"""
    pkg = os.path.join(ROOT, "swegl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "libswegl_ref" not in text, f
                if f.endswith(".py"):
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
