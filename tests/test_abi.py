"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/shark_b200.h declares.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


@pytest.fixture(scope="module")
def lib_path():
    from shark_b200 import build
    return build.build()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "shark_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(shk_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(lib_path):
    L = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), n


def test_binding_matches_header(lib_path):
    from shark_b200 import capi
    assert sorted(capi.EXPORTED) == _declared()
    L = capi.load()
    assert L.shk_abi_version() == 5


def test_struct_sizes():
    from shark_b200 import capi
    # must match the C layouts in include/shark_b200.h (x86-64 SysV)
    assert ctypes.sizeof(capi.Params) == 88
    assert ctypes.sizeof(capi.IndexInfo) == 96
    assert ctypes.sizeof(capi.Assoc) == 8
    assert ctypes.sizeof(capi.ChunkResult) == 112
    assert ctypes.sizeof(capi.IndexViews) == 64 + 64 + 96
    assert ctypes.sizeof(capi.ShardMem) == 24 + 192 + 8 + 16


def test_index_info_layout_matches_header(tmp_path):
    """Field by field: the ctypes view of shk_index_info against the header as gcc lays it out (the last field was
    `reserved` until round 2 and is `plain_front` now - same offset, same ABI version)."""
    import subprocess
    from shark_b200 import capi
    fields = [f[0] for f in capi.IndexInfo._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "shark_b200.h"\nint main(void) {\n'
                   '  printf("%zu", sizeof(shk_index_info));\n' +
                   "".join('  printf(" %%zu", offsetof(shk_index_info, %s));\n' % f for f in fields) +
                   '  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(capi.IndexInfo)
    assert out[1:] == [getattr(capi.IndexInfo, f).offset for f in fields]
    assert fields[-1] == "plain_front"


def test_no_cpu_fallback(lib_path):
    """Without a device shk_create must fail loudly (SHK_E_CUDA), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from shark_b200 import capi
    from shark_b200.engine import Shark
    with pytest.raises(capi.SharkError) as ei:
        Shark(k=17, bf_bits=1 << 20)
    assert ei.value.code == -2


def test_product_does_not_touch_oracle():
    """The product tree never imports, links or executes anything under oracle/."""
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "shark_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                if re.search(r"(import\s+oracle|from\s+oracle|oracle/|shko_|libshark_oracle)", txt):
                    # mentions in comments of the form "nothing here touches oracle/" are allowed
                    code = re.sub(r"#.*|//.*", "", txt)
                    code = re.sub(r'""".*?"""', "", code, flags=re.S)
                    if re.search(r"(import\s+oracle|from\s+oracle|shko_|libshark_oracle)", code):
                        bad.append(f)
    assert not bad, bad
