"""CPU checks of the C-ABI boundary: the library builds, loads without a GPU, exports every symbol that
include/ltgan.h declares, and the ctypes table in _lib.py agrees with the header's parameter lists."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    text = open(os.path.join(ROOT, "include", "ltgan.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:const char\*|int)\s+(ltg_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        decls[name] = plist
    return decls


def _kind(param):
    if "*" in param:
        return "ptr"
    for t, k in (("uint64_t", "u64"), ("int64_t", "i64"), ("uint32_t", "u32"), ("float", "f32"), ("int", "i32")):
        if re.search(r"\b%s\b" % t, param):
            return k
    raise AssertionError("unparsed parameter: " + param)


_CT = {ctypes.c_void_p: "ptr", ctypes.c_int: "i32", ctypes.c_int64: "i64", ctypes.c_uint64: "u64", ctypes.c_uint32: "u32",
       ctypes.c_float: "f32"}


def test_header_declares_entry_points():
    decls = _header_decls()
    assert len(decls) >= 20
    for required in ("ltg_gemm_bf16", "ltg_enc_gather_fwd", "ltg_dec_logits_fwd", "ltg_adam", "ltg_sample_pairs", "ltg_topk_metrics"):
        assert required in decls


def test_library_exports_every_declared_symbol(pkg, lib):
    decls = _header_decls()
    raw = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in decls:
        assert hasattr(raw, name), "libltgan.so does not export %s" % name


def test_ctypes_table_matches_header(pkg):
    decls = _header_decls()
    sigs = pkg._lib.SIGNATURES
    assert set(decls) == set(sigs), set(decls) ^ set(sigs)
    for name, plist in decls.items():
        _, args = sigs[name]
        assert len(args) == len(plist), "%s: header has %d params, ctypes table %d" % (name, len(plist), len(args))
        for i, (p, a) in enumerate(zip(plist, args)):
            assert _kind(p) == _CT[a], "%s param %d (%s): header %s vs ctypes %s" % (name, i, p, _kind(p), _CT[a])


def test_no_compute_without_gpu_but_version_works(lib):
    assert lib.ltg_version() >= 100
    # error path is reachable and descriptive on a box without a GPU; on a GPU box init succeeds
    rc = lib.ltg_init()
    assert rc in (0, -2, -3)
    if rc != 0:
        assert len(lib.ltg_last_error()) > 0
