"""Pins the CPU oracle (oracle/mel_oracle.c) and the test tooling against the real reference.

  * golden vectors produced by the UNMODIFIED reference build (tests/golden/golden_tensors.npz, make_golden.py);
  * the reference library itself when oracle/_ref is present (build container / GPU box);
  * the reference's own fixtures when /root/reference is present (mel filters of models/for-tests-*.bin).
"""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from open_whisper_kit_b200 import api, modelgen
from oracle import mel_oracle, reflib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_tensors.npz"))
FP = C.POINTER(C.c_float)


def test_mel_oracle_matches_reference_golden():
    pcm = api.read_wav_f32(os.path.join(HERE, "golden", "jfk.wav"))
    mel, n_org = mel_oracle.log_mel(pcm, modelgen.mel_filters(80))
    assert mel.shape == (80, 4100) and n_org == 1099
    g = GOLD["tiny.en/f1/fa0/mel_sub"]
    d = np.abs(mel[:, :1100:5] - g)
    # restated algorithm vs the compiled reference: same operations, different FMA contraction by the compiler
    assert d.max() <= 3e-5 and (d > 1e-6).mean() < 0.02
    assert abs(mel.astype(np.float64).sum() - GOLD["tiny.en/f1/fa0/mel_sum"][0]) < 1e-2


def test_mel_oracle_geometry_and_edge_cases():
    filt = modelgen.mel_filters(80)
    for n in (401, 1600, 16000, 100003, 480000):
        mel, n_org = mel_oracle.log_mel(modelgen.synth_pcm(n, stream=n % 7), filt)
        assert mel.shape == (80, (n + 480000) // 160) and n_org == 1 + (n + 200 - 400) // 160
        assert np.isfinite(mel).all()
    sil, _ = mel_oracle.log_mel(np.zeros(16000, np.float32), filt)
    assert np.allclose(sil, (-10.0 + 4.0) / 4.0)            # log10(1e-10) everywhere, clamp is a no-op


def test_mel_oracle_close_to_exact_float64():
    pcm = modelgen.synth_pcm(16000, stream=2)
    filt = modelgen.mel_filters(128)
    exact = mel_oracle.log_mel_f64(pcm, filt)
    mel, _ = mel_oracle.log_mel(pcm, filt)
    assert np.abs(mel - exact).max() < 2e-5


def test_mel_oracle_vs_live_reference():
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    for n_mel, pcm in ((80, modelgen.synth_pcm(64000, stream=1)), (128, modelgen.synth_pcm(33333, stream=4))):
        filt = modelgen.mel_filters(n_mel)
        n_len = (len(pcm) + 480000) // 160
        out = np.empty((n_mel, n_len), np.float32)
        a, b = C.c_int(), C.c_int()
        rc = ref.ref_log_mel(pcm.ctypes.data_as(FP), len(pcm), n_mel, filt.ctypes.data_as(FP), 2, out.ctypes.data_as(FP),
                             out.size, C.byref(a), C.byref(b))
        assert rc == 0 and a.value == n_len
        mel, n_org = mel_oracle.log_mel(pcm, filt)
        assert n_org == b.value
        assert np.abs(mel - out).max() <= 3e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference fixtures only exist in the build container")
def test_mel_filter_generator_matches_reference_fixture():
    b = open("/root/reference/models/for-tests-ggml-tiny.en.bin", "rb").read()
    n_mel, n_fft = struct.unpack_from("<2i", b, 48)
    fx = np.frombuffer(b, dtype=np.float32, count=n_mel * n_fft, offset=56).reshape(n_mel, n_fft)
    assert (n_mel, n_fft) == (80, 201)
    assert np.abs(fx - modelgen.mel_filters(80)).max() < 1e-8


def test_model_file_round_trip(tmp_path):
    """The generator writes exactly the container the reference reader expects (src/whisper.cpp:1485-1947)."""
    p = tmp_path / "m.bin"
    modelgen.write_model(str(p), "micro.en", seed=5)
    b = p.read_bytes()
    assert struct.unpack_from("<I", b, 0)[0] == 0x67676D6C
    hp = struct.unpack_from("<11i", b, 4)
    assert hp == (51864, 1500, 128, 2, 2, 448, 128, 2, 2, 80, 1)
    off = 48
    n_mel, n_fft = struct.unpack_from("<2i", b, off)
    off += 8 + n_mel * n_fft * 4
    n_tok, = struct.unpack_from("<i", b, off)
    off += 4
    toks = []
    for _ in range(n_tok):
        ln, = struct.unpack_from("<I", b, off)
        toks.append(b[off + 4: off + 4 + ln])
        off += 4 + ln
    assert n_tok == 50257 and toks[220] == b" " and len(set(toks)) == n_tok
    names = {}
    while off < len(b):
        n_dims, ln, ttype = struct.unpack_from("<3i", b, off)
        off += 12
        ne = struct.unpack_from(f"<{n_dims}i", b, off)
        off += 4 * n_dims
        name = b[off: off + ln].decode()
        off += ln
        off += int(np.prod(ne)) * (2 if ttype == 1 else 4)
        names[name] = (ne, ttype)
    assert off == len(b)
    assert len(names) == 7 + 15 * 2 + 4 + 24 * 2
    assert names["encoder.conv1.weight"] == ((3, 80, 128), 1) and names["encoder.conv1.bias"] == ((1, 128), 0)
    assert names["decoder.token_embedding.weight"] == ((128, 51864), 1)
    assert names["decoder.blocks.1.cross_attn.key.weight"][0] == (128, 128)
    # seeded: bit-identical on regeneration
    q = tmp_path / "m2.bin"
    modelgen.write_model(str(q), "micro.en", seed=5)
    assert q.read_bytes() == b


def test_golden_token_fixture_is_consistent():
    import json
    g = json.load(open(os.path.join(HERE, "golden", "golden_tokens.json")))
    assert {"tiny.en/jfk/ts/fa0", "base.en/synth16/ts/fa0", "base.en/synth16/nots/fa0"} <= set(g)
    nots = g["base.en/synth16/nots/fa0"]
    assert len(nots["segments"]) == 16 and all(len(s[2]) == 220 for s in nots["segments"])
    # the per-step top-2 gaps of the reference: near-ties far below any 16-bit implementation's logit error exist
    gaps = np.concatenate([np.array(w["gaps"]) for w in nots["steps"]])
    assert gaps.size == 3520 and gaps.min() < 1e-3 and np.median(gaps) > 0.05


# ---- ggml block quantisation (modelgen.quantize_blocks / dequantize_blocks) ----------------------------------------
QUANT_GOLD = os.path.join(HERE, "golden", "golden_quant.npz")


def _quant_inputs():
    rng = np.random.default_rng(20260118)
    x = (rng.standard_normal(32 * 257) * 0.07).astype(np.float32)
    x[:32] = 0.0                                   # an all-zero block (d = 0)
    x[32:64] = np.float32(0.125)                   # a constant block (q4_1 / q5_1: max == min)
    x[64] = np.float32(-3.0)                       # one large negative / positive outlier per block
    x[96 + 17] = np.float32(2.5)
    return x


@pytest.mark.parametrize("qtype", sorted(modelgen.QUANT_TYPES))
def test_block_quantisers_match_reference_golden(qtype):
    """The numpy restatement writes the same block bytes and expands them to the same floats as the reference's
    quantize_row_*_ref / dequantize_row_* (golden vectors made by tests/golden/make_golden_quant.py from oracle/_ref)."""
    g = np.load(QUANT_GOLD)
    x = _quant_inputs()
    raw = modelgen.quantize_blocks(x, qtype)
    assert raw == g[f"{qtype}/raw"].tobytes()
    y = modelgen.dequantize_blocks(raw, qtype)
    assert np.array_equal(y.view(np.uint32), g[f"{qtype}/deq"].view(np.uint32))


@pytest.mark.parametrize("qtype", sorted(modelgen.QUANT_TYPES))
def test_block_quantisers_vs_live_reference(qtype):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(32 * 4096) * rng.uniform(0.001, 2.0)).astype(np.float32)
    bb = modelgen.QUANT_TYPES[qtype][2]
    raw = np.zeros(len(x) // 32 * bb, np.uint8)
    getattr(ref, f"quantize_row_{qtype}_ref")(x.ctypes.data_as(FP), raw.ctypes.data_as(C.c_void_p), C.c_int64(len(x)))
    assert modelgen.quantize_blocks(x, qtype) == raw.tobytes()
    y = np.empty_like(x)
    getattr(ref, f"dequantize_row_{qtype}")(raw.ctypes.data_as(C.c_void_p), y.ctypes.data_as(FP), C.c_int64(len(x)))
    assert np.array_equal(modelgen.dequantize_blocks(raw.tobytes(), qtype).view(np.uint32), y.view(np.uint32))


# ---- expansion of quantised blocks: numpy restatement, the product's loader (host-only hook) and the reference ----------
_SCALE_COLS = {"q4_0": (0,), "q4_1": (0, 2), "q5_0": (0,), "q5_1": (0, 2), "q8_0": (0,),
               "q2_K": (80, 82), "q3_K": (108,), "q4_K": (0, 2), "q5_K": (0, 2), "q6_K": (208,)}
ALL_QTYPES = sorted(modelgen.QUANT_TYPES) + sorted(modelgen.KQUANT_TYPES)


def _random_blocks(qtype, n_blocks=257):
    """Seeded random block bytes (every bit pattern of the packed fields occurs) with small finite f16 scales."""
    k = qtype in modelgen.KQUANT_TYPES
    bb = (modelgen.KQUANT_TYPES if k else modelgen.QUANT_TYPES)[qtype][2]
    rng = np.random.default_rng(sum(map(ord, qtype)))
    raw = rng.integers(0, 256, (n_blocks, bb), dtype=np.uint8)
    for c in _SCALE_COLS[qtype]:
        raw[:, c:c + 2] = rng.uniform(-0.03, 0.03, n_blocks).astype(np.float16).view(np.uint8).reshape(n_blocks, 2)
    return raw, (256 if k else 32)


def _numpy_expand16(raw, qtype):
    fn = modelgen.dequantize_kblocks if qtype in modelgen.KQUANT_TYPES else modelgen.dequantize_blocks
    return fn(raw.tobytes(), qtype).astype(np.float16)


@pytest.mark.parametrize("qtype", ALL_QTYPES)
def test_block_expansion_matches_reference_golden(qtype):
    raw, _ = _random_blocks(qtype)
    g = np.load(QUANT_GOLD)[f"{qtype}/rand_deq16"]
    assert np.array_equal(_numpy_expand16(raw, qtype).view(np.uint16), g.view(np.uint16))


@pytest.mark.parametrize("qtype", ALL_QTYPES)
def test_loader_block_expansion_is_the_oracles(qtype):
    """csrc/model.cu's load-time expansion (exported host-only hook, no device needed) == the pinned numpy restatement."""
    import open_whisper_kit_b200 as pkg
    lib = pkg.load()
    raw, n_el = _random_blocks(qtype)
    tt = (modelgen.KQUANT_TYPES if qtype in modelgen.KQUANT_TYPES else modelgen.QUANT_TYPES)[qtype][0]
    out = np.zeros(raw.shape[0] * n_el, np.uint16)
    n = lib.whisper_b200_dequantize_blocks(tt, raw.ctypes.data_as(C.c_void_p), raw.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint16)))
    assert n == out.size
    assert np.array_equal(out, _numpy_expand16(raw, qtype).view(np.uint16))
    assert lib.whisper_b200_dequantize_blocks(9, raw.ctypes.data_as(C.c_void_p), 1, out.ctypes.data_as(C.POINTER(C.c_uint16))) == -1


@pytest.mark.parametrize("qtype", sorted(modelgen.KQUANT_TYPES))
def test_kquant_expansion_vs_live_reference(qtype):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    raw, n_el = _random_blocks(qtype, n_blocks=1500)
    y = np.empty(raw.shape[0] * n_el, np.float32)
    getattr(ref, f"dequantize_row_{qtype}")(raw.ctypes.data_as(C.c_void_p), y.ctypes.data_as(FP), C.c_int64(y.size))
    mine = modelgen.dequantize_kblocks(raw.tobytes(), qtype)
    assert np.array_equal(mine.view(np.uint32), y.view(np.uint32))
