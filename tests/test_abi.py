"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every
symbol include/glass_b200.h declares; the ctypes structs match the header's field order.  No compute
call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from glass_text_spotting_b200 import build, lib
    build.build()
    return lib.load()


def _header():
    return open(os.path.join(ROOT, "include", "glass_b200.h")).read()


def _declared_symbols():
    return sorted(set(re.findall(r"^\s*(?:const char\*|int64_t|int)\s+(glass_\w+)\s*\(", _header(), flags=re.M)))


def test_header_symbols_all_exported(built_lib):
    from glass_text_spotting_b200 import lib
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in glass_b200.h but not exported"
    assert sorted(lib.SYMBOLS) == declared, "lib.SYMBOLS out of sync with the header"


def test_abi_version_and_error_string(built_lib):
    assert built_lib.glass_abi_version() == 2
    assert isinstance(built_lib.glass_last_error(), bytes)
    assert built_lib.glass_launch_count() == 0


def _struct_fields(name):
    body = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\}\s*" + name + ";", _header(), flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(",")
        first = names[0].split()[-1]
        for nm in [first] + [x.strip() for x in names[1:]]:
            fields.append(re.sub(r"[\*\s]|\[.*\]", "", nm))
    return fields


@pytest.mark.parametrize("cname,pyname", [("GlassConvGemmParams", "ConvGemmParams"),
                                           ("GlassRoiAlignParams", "RoiAlignParams"),
                                           ("GlassImageRoiAlignParams", "ImageRoiAlignParams"),
                                           ("GlassRpnTopkParams", "RpnTopkParams"),
                                           ("GlassNmsParams", "NmsParams"),
                                           ("GlassGcAttentionParams", "GcAttentionParams"),
                                           ("GlassAsterParams", "AsterParams"),
                                           ("GlassPostprocessParams", "PostprocessParams")])
def test_ctypes_structs_match_header(cname, pyname):
    from glass_text_spotting_b200 import lib
    assert [f[0] for f in getattr(lib, pyname)._fields_] == _struct_fields(cname)


def test_validation_errors_do_not_need_a_gpu(built_lib):
    """Shape validation happens on the host before any launch: bad params -> negative rc + message."""
    from glass_text_spotting_b200 import lib
    p = lib.ConvGemmParams()
    assert built_lib.glass_conv_gemm(ctypes.byref(p), None) < 0
    assert b"a_hi" in built_lib.glass_last_error()
    assert built_lib.glass_launch_count() == 0


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    from glass_text_spotting_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.load()


def test_ops_reject_cpu_tensors(built_lib):
    import torch
    from glass_text_spotting_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.Act.from_nchw(torch.zeros(1, 3, 4, 4))


def test_integration_doc_stub_matches_the_binding():
    """INTEGRATION.md shows the ctypes stub a reference maintainer would add; its struct must be the header's."""
    from glass_text_spotting_b200 import lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blk = doc[doc.index("class _RoiAlignParams"):doc.index("def roi_align_rotated_forward")]
    assert re.findall(r'\("(\w+)"', blk) == [f[0] for f in lib.RoiAlignParams._fields_]
    assert re.findall(r'\("(\w+)"', blk) == _struct_fields("GlassRoiAlignParams")
    for sym in ("glass_roi_align_rotated", "glass_nms_rotated_all", "glass_box_iou_rotated", "glass_postprocess_merge"):
        assert sym in doc and sym in _declared_symbols()
