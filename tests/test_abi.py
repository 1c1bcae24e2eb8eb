"""The drop-in boundary: include/whisper.h must be ABI-identical to the reference's header.

A C program compiled against our header prints sizeof / offsetof of every by-value struct; the numbers must equal
the ctypes mirror used by the tests and, when /root/reference is present (build container), the same program
compiled against the reference's own include/whisper.h + ggml headers."""
import ctypes as C
import os
import re
import subprocess

import pytest

from open_whisper_kit_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STRUCTS = {
    "whisper_context_params": capi.whisper_context_params,
    "whisper_full_params": capi.whisper_full_params,
    "whisper_token_data": capi.whisper_token_data,
    "whisper_vad_params": capi.whisper_vad_params,
    "whisper_timings": capi.whisper_timings,
}


def _flat_fields(st, prefix=""):
    out = []
    for name, typ in st._fields_:
        if issubclass(typ, C.Structure) and name in ("greedy", "beam_search", "vad_params", "dtw_aheads"):
            for sub, off in _flat_fields(typ, prefix + name + "."):
                out.append((sub, getattr(st, name).offset + off))
        else:
            out.append((prefix + name, getattr(st, name).offset))
    return out


def _probe_source():
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "whisper.h"', 'int main(void) {']
    for sname, st in STRUCTS.items():
        lines.append(f'  printf("{sname} %zu\\n", sizeof(struct {sname}));')
        for f, _ in _flat_fields(st):
            lines.append(f'  printf("{sname}.{f} %zu\\n", offsetof(struct {sname}, {f}));')
    lines += ['  printf("enum %d %d %d\\n", (int) WHISPER_SAMPLING_BEAM_SEARCH, (int) WHISPER_AHEADS_LARGE_V3_TURBO, (int) WHISPER_GRETYPE_CHAR_ALT);',
              '  return 0; }']
    return "\n".join(lines)


def _run_probe(tmp_path, includes, tag):
    src = tmp_path / f"probe_{tag}.c"
    exe = tmp_path / f"probe_{tag}"
    src.write_text(_probe_source())
    cmd = ["gcc", "-std=c11", "-o", str(exe), str(src)] + [f"-I{i}" for i in includes]
    subprocess.run(cmd, check=True, capture_output=True)
    return subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout


def test_header_layout_matches_ctypes(tmp_path):
    out = _run_probe(tmp_path, [os.path.join(ROOT, "include")], "ours")
    got = dict(line.rsplit(" ", 1) for line in out.strip().splitlines() if not line.startswith("enum"))
    for sname, st in STRUCTS.items():
        assert int(got[sname]) == C.sizeof(st), sname
        for f, off in _flat_fields(st):
            assert int(got[f"{sname}.{f}"]) == off, f"{sname}.{f}"


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="reference tree only exists in the build container")
def test_header_layout_matches_reference_header(tmp_path):
    ours = _run_probe(tmp_path, [os.path.join(ROOT, "include")], "ours")
    ref = _run_probe(tmp_path, ["/root/reference/include", "/root/reference/ggml/include"], "ref")
    assert ours == ref


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="reference tree only exists in the build container")
def test_every_reference_api_symbol_is_declared_and_bound():
    """Every WHISPER_API prototype of the reference header has a same-named entry in our header and in capi.PROTOTYPES."""
    ref = open("/root/reference/include/whisper.h").read()
    ours = open(os.path.join(ROOT, "include", "whisper.h")).read()
    names = set(re.findall(r"\b(whisper_[a-z0-9_]+)\s*\(", ref))
    names = {n for n in names if re.search(r"WHISPER_API[^;]*\b" + n + r"\s*\(", ref, re.S) or n == "whisper_token_count"}
    assert len(names) > 100
    missing_header = [n for n in sorted(names) if not re.search(r"\b" + n + r"\s*\(", ours)]
    missing_ctypes = [n for n in sorted(names) if n not in capi.PROTOTYPES]
    assert not missing_header, missing_header
    assert not missing_ctypes, missing_ctypes
