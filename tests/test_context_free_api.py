"""Context-free part of the C API against the compiled reference, on the CPU: default parameter structs (returned by value --
every field, recursively), the language table, and the by-ref / free pairs.  (reference src/whisper.cpp:3606-3622, 5912-6034,
3976-4019)."""
import ctypes as C

import pytest

import open_whisper_kit_b200 as pkg
from open_whisper_kit_b200 import capi
from oracle import reflib


def _fields(obj, prefix=""):
    out = {}
    for name, typ in obj._fields_:
        v = getattr(obj, name)
        if isinstance(v, C.Structure):
            out.update(_fields(v, prefix + name + "."))
        elif isinstance(v, C._Pointer):
            out[prefix + name] = C.cast(v, C.c_void_p).value          # address (None = NULL)
        else:
            out[prefix + name] = v
    return out


@pytest.fixture(scope="module")
def libs():
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    return pkg.load(), ref


@pytest.mark.parametrize("strategy", [capi.GREEDY, capi.BEAM_SEARCH])
def test_full_default_params_equal_reference(libs, strategy):
    ours, ref = libs
    a, b = _fields(ours.whisper_full_default_params(strategy)), _fields(ref.whisper_full_default_params(strategy))
    assert a.keys() == b.keys() and len(a) > 60
    for k in a:
        assert a[k] == b[k] or (a[k] != a[k] and b[k] != b[k]), k
    pa = ours.whisper_full_default_params_by_ref(strategy)
    pb = ref.whisper_full_default_params_by_ref(strategy)
    assert _fields(C.cast(pa, C.POINTER(capi.whisper_full_params)).contents) == _fields(C.cast(pb, C.POINTER(capi.whisper_full_params)).contents)
    ours.whisper_free_params(pa)
    ref.whisper_free_params(pb)


def test_context_default_params_equal_reference(libs):
    ours, ref = libs
    a, b = _fields(ours.whisper_context_default_params()), _fields(ref.whisper_context_default_params())
    assert a.keys() == b.keys()
    for k in a:
        if k == "use_gpu":
            continue            # the reference's default depends on how it was built; ours is the GPU path by definition
        assert a[k] == b[k], k


def test_language_table_equals_reference(libs):
    ours, ref = libs
    n = ref.whisper_lang_max_id()
    assert ours.whisper_lang_max_id() == n and n >= 99
    for i in range(n + 1):
        code, full = ref.whisper_lang_str(i), ref.whisper_lang_str_full(i)
        assert ours.whisper_lang_str(i) == code and ours.whisper_lang_str_full(i) == full
        assert ours.whisper_lang_id(code) == ref.whisper_lang_id(code) == i
        assert ours.whisper_lang_id(full) == ref.whisper_lang_id(full)
    for bad in (b"xx", b"", b"klingon", b"EN"):
        assert ours.whisper_lang_id(bad) == ref.whisper_lang_id(bad)
    assert ours.whisper_lang_str(n + 1) == ref.whisper_lang_str(n + 1) and ours.whisper_lang_str(-1) == ref.whisper_lang_str(-1)
