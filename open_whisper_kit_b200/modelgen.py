"""Random-init GGML model files of a named Whisper architecture + synthetic PCM (test / bench tooling).

There is no network for real checkpoints, so parity and throughput are measured on seeded
random-init weights written in the reference's legacy GGML container
(reference: src/whisper.cpp:1485-1956 reader, models/convert-pt-to-ggml.py:268-337 writer;
tensor names src/whisper-arch.h:42-109).  The same file is fed to the reference CPU build and to
our library.

Initialisation is variance preserving: sigma = 1/sqrt(fan_in) for GEMM/conv
weights, 0.5x on residual-branch output projections, LayerNorm gamma = 1 + N(0, 0.02), biases
N(0, 0.02), sinusoidal encoder positions (as real checkpoints have), so the residual stream stays
O(1) and attention is not uniform.
"""
import struct

import numpy as np

GGML_MAGIC = 0x67676D6C
N_FFT_BINS = 201
SAMPLE_RATE = 16000
WINDOW_SAMPLES = 30 * SAMPLE_RATE

# name -> (n_vocab, n_audio_ctx, d, n_head, n_audio_layer, n_text_ctx, n_text_layer, n_mels)
ARCHS = {
    "tiny.en": (51864, 1500, 384, 6, 4, 448, 4, 80),
    "tiny": (51865, 1500, 384, 6, 4, 448, 4, 80),
    "base.en": (51864, 1500, 512, 8, 6, 448, 6, 80),
    "base": (51865, 1500, 512, 8, 6, 448, 6, 80),
    "small.en": (51864, 1500, 768, 12, 12, 448, 12, 80),
    "medium.en": (51864, 1500, 1024, 16, 24, 448, 24, 80),
    "large-v2": (51865, 1500, 1280, 20, 32, 448, 32, 80),
    "large-v3": (51866, 1500, 1280, 20, 32, 448, 32, 128),
    "large-v3-turbo": (51866, 1500, 1280, 20, 32, 448, 4, 128),
    # a 2-layer toy with the tiny geometry for fast CPU tests of the full pipeline
    "micro.en": (51864, 1500, 128, 2, 2, 448, 2, 80),
}


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filters(n_mels, sr=SAMPLE_RATE, n_fft=400):
    """Slaney-scale, area-normalised triangular filterbank [n_mels][201] (what the model files carry,
    reference src/whisper.cpp:1576-1586; reproduces the 80-bin fixture to ~2e-9)."""
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def _gpt2_byte_order():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    rest = [b for b in range(256) if b not in bs]
    return bs + rest


NON_SPEECH = ["\"", "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^", "_", "`",
              "{", "|", "}", "~", "<<", ">>", "<<<", ">>>", "--", "---", "-(", "-[", "('", "(\"", "((", "))",
              "(((", ")))", "[[", "]]", "{{", "}}"]


def synth_vocab(multilingual, seed=99):
    """50257 byte strings shaped like the GPT-2 BPE table: ids 0..255 are the single bytes in GPT-2 order
    (so " " is id 220, which the suppress_blank rule looks up, reference src/whisper.cpp:6219), followed by
    the multi-character non-speech strings the suppress_nst rule searches for and seeded pseudo-words."""
    toks = [bytes([b]) for b in _gpt2_byte_order()]
    seen = set(toks)
    for t in NON_SPEECH + ["-", "'"]:
        for cand in (t.encode(), b" " + t.encode()):
            if cand not in seen:
                seen.add(cand)
                toks.append(cand)
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"etaoinshrdlucmfwypvbgkqjxz", dtype=np.uint8)
    while len(toks) < 50256:
        n = int(rng.integers(2, 8))
        w = bytes(rng.choice(letters, size=n).tolist())
        if rng.random() < 0.5:
            w = b" " + w
        if w not in seen:
            seen.add(w)
            toks.append(w)
    toks.append(b"" if multilingual else b"<|endoftext|>")
    return toks


def sinusoids(length, channels, max_timescale=10000.0):
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2))
    t = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def tensor_specs(arch):
    """Yield (name, torch-order shape, kind) for every tensor in file order (Appendix A of SURVEY.md)."""
    n_vocab, n_actx, d, n_head, n_al, n_tctx, n_tl, n_mels = ARCHS[arch]
    out = []
    out.append(("encoder.positional_embedding", (n_actx, d), "enc_pos"))
    out.append(("encoder.conv1.weight", (d, n_mels, 3), "conv"))
    out.append(("encoder.conv1.bias", (d, 1), "bias2d"))
    out.append(("encoder.conv2.weight", (d, d, 3), "conv"))
    out.append(("encoder.conv2.bias", (d, 1), "bias2d"))

    def block(prefix, cross):
        b = []
        b.append((prefix + ".attn_ln.weight", (d,), "ln_w"))
        b.append((prefix + ".attn_ln.bias", (d,), "bias"))
        b.append((prefix + ".attn.query.weight", (d, d), "w"))
        b.append((prefix + ".attn.query.bias", (d,), "bias"))
        b.append((prefix + ".attn.key.weight", (d, d), "w"))
        b.append((prefix + ".attn.value.weight", (d, d), "w"))
        b.append((prefix + ".attn.value.bias", (d,), "bias"))
        b.append((prefix + ".attn.out.weight", (d, d), "w_out"))
        b.append((prefix + ".attn.out.bias", (d,), "bias"))
        if cross:
            b.append((prefix + ".cross_attn_ln.weight", (d,), "ln_w"))
            b.append((prefix + ".cross_attn_ln.bias", (d,), "bias"))
            b.append((prefix + ".cross_attn.query.weight", (d, d), "w"))
            b.append((prefix + ".cross_attn.query.bias", (d,), "bias"))
            b.append((prefix + ".cross_attn.key.weight", (d, d), "w"))
            b.append((prefix + ".cross_attn.value.weight", (d, d), "w"))
            b.append((prefix + ".cross_attn.value.bias", (d,), "bias"))
            b.append((prefix + ".cross_attn.out.weight", (d, d), "w_out"))
            b.append((prefix + ".cross_attn.out.bias", (d,), "bias"))
        b.append((prefix + ".mlp_ln.weight", (d,), "ln_w"))
        b.append((prefix + ".mlp_ln.bias", (d,), "bias"))
        b.append((prefix + ".mlp.0.weight", (4 * d, d), "w"))
        b.append((prefix + ".mlp.0.bias", (4 * d,), "bias"))
        b.append((prefix + ".mlp.2.weight", (d, 4 * d), "w_out"))
        b.append((prefix + ".mlp.2.bias", (d,), "bias"))
        return b

    for i in range(n_al):
        out += block(f"encoder.blocks.{i}", False)
    out.append(("encoder.ln_post.weight", (d,), "ln_w"))
    out.append(("encoder.ln_post.bias", (d,), "bias"))
    out.append(("decoder.positional_embedding", (n_tctx, d), "dec_pos"))
    out.append(("decoder.token_embedding.weight", (n_vocab, d), "tok_emb"))
    for i in range(n_tl):
        out += block(f"decoder.blocks.{i}", True)
    out.append(("decoder.ln.weight", (d,), "ln_w"))
    out.append(("decoder.ln.bias", (d,), "bias"))
    return out


def _init(kind, shape, rng, tok_emb_gain):
    """Seeded variance-preserving initialisation of one tensor (see module docstring)."""
    if kind == "enc_pos":
        return sinusoids(shape[0], shape[1])
    if kind == "dec_pos":
        # large enough that the decoder state depends on the position: with a small table a random-init decoder
        # falls into a one-token fixed point after a few steps, which would make token parity a weak test
        return (1.0 * rng.standard_normal(shape, dtype=np.float32))
    if kind == "ln_w":
        return (1.0 + 0.02 * rng.standard_normal(shape, dtype=np.float32)).astype(np.float32)
    if kind in ("bias", "bias2d"):
        return (0.02 * rng.standard_normal(shape, dtype=np.float32))
    if kind == "conv":
        fan_in = shape[1] * shape[2]
        return (1.4 / np.sqrt(fan_in)) * rng.standard_normal(shape, dtype=np.float32)
    if kind == "w":
        return (1.0 / np.sqrt(shape[1])) * rng.standard_normal(shape, dtype=np.float32)
    if kind == "w_out":
        return (0.5 / np.sqrt(shape[1])) * rng.standard_normal(shape, dtype=np.float32)
    if kind == "tok_emb":
        return (tok_emb_gain / np.sqrt(shape[1])) * rng.standard_normal(shape, dtype=np.float32)
    raise ValueError(kind)


# ---- ggml 32-element block formats (what `examples/quantize` writes) ----------------------------------------------
# ggml_type of a record / ggml_ftype of the file header (reference ggml/include/ggml.h), bytes per block
QUANT_TYPES = {"q4_0": (2, 2, 18), "q4_1": (3, 3, 20), "q5_0": (6, 8, 22), "q5_1": (7, 9, 24), "q8_0": (8, 7, 34)}
GGML_QNT_VERSION = 2


def quantize_blocks(x, qtype):
    """numpy restatement of quantize_row_{q4_0,q4_1,q5_0,q5_1,q8_0}_ref (reference ggml/src/ggml-quants.c:36-222):
    x f32, size a multiple of 32 -> raw block bytes.  float32 arithmetic, C truncation where the reference casts."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 32)
    nb = x.shape[0]
    f32 = np.float32
    if qtype == "q8_0":
        amax = np.abs(x).max(axis=1)
        d = (amax / f32(127)).astype(f32)
        idv = np.where(d != 0, f32(1) / np.where(d != 0, d, f32(1)), f32(0)).astype(f32)
        v = x * idv[:, None]
        q = (np.sign(v) * np.floor(np.abs(v) + f32(0.5))).astype(np.int8)          # roundf: half away from zero
        out = np.empty((nb, 34), np.uint8)
        out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(nb, 2)
        out[:, 2:] = q.view(np.uint8)
        return out.tobytes()
    sym = qtype in ("q4_0", "q5_0")
    bits = 4 if qtype in ("q4_0", "q4_1") else 5
    top = (1 << bits) - 1
    if sym:
        idx = np.abs(x).argmax(axis=1)                       # first element with the largest magnitude
        mx = x[np.arange(nb), idx]
        d = (mx / f32(-(1 << (bits - 1)))).astype(f32)
        mn = np.zeros(nb, f32)
        off = f32((1 << (bits - 1)) + 0.5)
    else:
        mn = x.min(axis=1)
        d = ((x.max(axis=1) - mn) / f32(top)).astype(f32)
        off = f32(0.5)
    idv = np.where(d != 0, f32(1) / np.where(d != 0, d, f32(1)), f32(0)).astype(f32)
    # the reference build contracts `t * id + off` into one fused multiply-add (gcc -O3 -mfma): the float64 product of two
    # float32 values is exact, so rounding the float64 sum once more to float32 reproduces it
    t = (x - mn[:, None]).astype(f32)
    v = (t.astype(np.float64) * idv[:, None].astype(np.float64) + np.float64(off)).astype(f32)
    if qtype == "q5_1":
        q = np.trunc(v).astype(np.int64).astype(np.uint8)                         # (uint8_t) cast, no clamp
    else:
        q = np.minimum(top, np.trunc(v).astype(np.int64).astype(np.int8)).astype(np.uint8)      # MIN(top, (int8_t) v)
    lo, hi = q[:, :16], q[:, 16:]
    qs = ((lo & 0x0F) | ((hi & 0x0F) << 4)).astype(np.uint8)
    parts = [d.astype(np.float16).view(np.uint8).reshape(nb, 2)]
    if not sym:
        parts.append(mn.astype(np.float16).view(np.uint8).reshape(nb, 2))
    if bits == 5:
        sh = np.arange(16, dtype=np.uint32)
        qh = (((lo.astype(np.uint32) & 0x10) >> 4) << sh).sum(axis=1, dtype=np.uint32) | \
             (((hi.astype(np.uint32) & 0x10) >> 4) << (sh + 16)).sum(axis=1, dtype=np.uint32)
        parts.append(qh.astype("<u4").view(np.uint8).reshape(nb, 4))
    parts.append(qs)
    return np.concatenate(parts, axis=1).tobytes()


def dequantize_blocks(raw, qtype):
    """numpy restatement of dequantize_row_* (reference ggml/src/ggml-quants.c:307-415): raw block bytes -> f32."""
    bb = QUANT_TYPES[qtype][2]
    b = np.frombuffer(raw, np.uint8).reshape(-1, bb)
    nb = b.shape[0]
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32).reshape(nb)
    if qtype == "q8_0":
        return (b[:, 2:].copy().view(np.int8).astype(np.float32) * d[:, None]).reshape(-1)
    sym = qtype in ("q4_0", "q5_0")
    five = qtype in ("q5_0", "q5_1")
    o = 2
    m = np.zeros(nb, np.float32)
    if not sym:
        m = b[:, 2:4].copy().view(np.float16).astype(np.float32).reshape(nb)
        o = 4
    lo = (b[:, o + (4 if five else 0):] & 0x0F).astype(np.int32)
    hi = (b[:, o + (4 if five else 0):] >> 4).astype(np.int32)
    if five:
        qh = b[:, o:o + 4].copy().view("<u4").reshape(nb).astype(np.uint32)
        j = np.arange(16, dtype=np.uint32)
        lo |= (((qh[:, None] >> j) << 4) & 0x10).astype(np.int32)
        hi |= ((qh[:, None] >> (j + 12)) & 0x10).astype(np.int32)
    if sym:
        lo -= 16 if five else 8
        hi -= 16 if five else 8
    y = np.concatenate([lo, hi], axis=1).astype(np.float32) * d[:, None]
    if not sym:
        y = y + m[:, None]           # (symmetric formats: no `+ 0`, the reference keeps the sign of -8 * 0)
    return y.astype(np.float32).reshape(-1)


# ---- 256-element super-block formats ("K-quants") -------------------------------------------------------------------
# ggml_type == ggml_ftype for these (reference ggml/include/ggml.h:400-404, 450-454), bytes per block
KQUANT_TYPES = {"q2_K": (10, 10, 84), "q3_K": (11, 11, 110), "q4_K": (12, 12, 144), "q5_K": (13, 13, 176), "q6_K": (14, 14, 210)}


def _scale_min_k4(sc):
    """sc: [nb][12] uint8 -> (scales [nb][8], mins [nb][8]) as in get_scale_min_k4 (ggml/src/ggml-quants.c:703-710)."""
    sc = sc.astype(np.int32)
    d = np.empty((sc.shape[0], 8), np.int32)
    m = np.empty((sc.shape[0], 8), np.int32)
    d[:, :4] = sc[:, 0:4] & 63
    m[:, :4] = sc[:, 4:8] & 63
    d[:, 4:] = (sc[:, 8:12] & 0xF) | ((sc[:, 0:4] >> 6) << 4)
    m[:, 4:] = (sc[:, 8:12] >> 4) | ((sc[:, 4:8] >> 6) << 4)
    return d, m


def dequantize_kblocks(raw, qtype):
    """numpy restatement of dequantize_row_q2_K / q3_K / q4_K / q5_K / q6_K (reference ggml/src/ggml-quants.c:784-814,
    1128-1176, 1352-1374, 1554-1579, 1762-1791): raw block bytes -> f32, same f32 expressions (no fused multiply-add)."""
    bb = KQUANT_TYPES[qtype][2]
    b = np.frombuffer(raw, np.uint8).reshape(-1, bb)
    nb = b.shape[0]
    f32 = np.float32

    def h(col):
        return b[:, col:col + 2].copy().view(np.float16).astype(f32).reshape(nb)

    y = np.empty((nb, 256), f32)
    if qtype == "q2_K":
        scales, q = b[:, 0:16].astype(np.int32), b[:, 16:80].astype(np.int32)
        d, mn = h(80), h(82)
        o, i_s = 0, 0
        for n in range(2):
            for j in range(4):
                for half in range(2):
                    sc = scales[:, i_s]
                    i_s += 1
                    dl = d * (sc & 0xF).astype(f32)
                    ml = mn * (sc >> 4).astype(f32)
                    v = (q[:, 32 * n + 16 * half:32 * n + 16 * half + 16] >> (2 * j)) & 3
                    y[:, o:o + 16] = dl[:, None] * v.astype(f32) - ml[:, None]
                    o += 16
    elif qtype == "q3_K":
        hm, q = b[:, 0:32].astype(np.int32), b[:, 32:96].astype(np.int32)
        d_all = h(108)
        aux = b[:, 96:108].copy().view("<u4").astype(np.uint32)              # [nb][3]
        k1, k2 = np.uint32(0x03030303), np.uint32(0x0f0f0f0f)
        a0, a1, tmp = aux[:, 0], aux[:, 1], aux[:, 2]
        out = np.stack([(a0 & k2) | (((tmp >> 0) & k1) << 4), (a1 & k2) | (((tmp >> 2) & k1) << 4),
                        ((a0 >> 4) & k2) | (((tmp >> 4) & k1) << 4), ((a1 >> 4) & k2) | (((tmp >> 6) & k1) << 4)], axis=1)
        scales = np.ascontiguousarray(out.astype("<u4")).view(np.int8).reshape(nb, 16).astype(np.int32)
        o, i_s = 0, 0
        for n in range(2):
            for j in range(4):
                mbit = 1 << (4 * n + j)
                for half in range(2):
                    dl = d_all * (scales[:, i_s] - 32).astype(f32)
                    i_s += 1
                    sl = slice(16 * half, 16 * half + 16)
                    v = ((q[:, 32 * n:32 * n + 32][:, sl] >> (2 * j)) & 3) - np.where(hm[:, sl] & mbit, 0, 4)
                    y[:, o:o + 16] = dl[:, None] * v.astype(f32)
                    o += 16
    elif qtype in ("q4_K", "q5_K"):
        d, mn = h(0), h(2)
        sc, m = _scale_min_k4(b[:, 4:16])
        five = qtype == "q5_K"
        qh = b[:, 16:48].astype(np.int32) if five else None
        q = b[:, 48:176].astype(np.int32) if five else b[:, 16:144].astype(np.int32)
        for g in range(4):
            ql = q[:, 32 * g:32 * g + 32]
            lo, hi = ql & 0xF, ql >> 4
            if five:
                lo = lo + np.where(qh & (1 << (2 * g)), 16, 0)
                hi = hi + np.where(qh & (2 << (2 * g)), 16, 0)
            d1, m1 = d * sc[:, 2 * g].astype(f32), mn * m[:, 2 * g].astype(f32)
            d2, m2 = d * sc[:, 2 * g + 1].astype(f32), mn * m[:, 2 * g + 1].astype(f32)
            y[:, 64 * g:64 * g + 32] = d1[:, None] * lo.astype(f32) - m1[:, None]
            y[:, 64 * g + 32:64 * g + 64] = d2[:, None] * hi.astype(f32) - m2[:, None]
    else:       # q6_K
        ql, qh = b[:, 0:128].astype(np.int32), b[:, 128:192].astype(np.int32)
        sc = b[:, 192:208].copy().view(np.int8).astype(f32)
        d = h(208)
        for n in range(2):
            l_, h_, s_ = ql[:, 64 * n:64 * n + 64], qh[:, 32 * n:32 * n + 32], sc[:, 8 * n:8 * n + 8]
            qs = [((l_[:, 0:32] & 0xF) | (((h_ >> 0) & 3) << 4)) - 32, ((l_[:, 32:64] & 0xF) | (((h_ >> 2) & 3) << 4)) - 32,
                  ((l_[:, 0:32] >> 4) | (((h_ >> 4) & 3) << 4)) - 32, ((l_[:, 32:64] >> 4) | (((h_ >> 6) & 3) << 4)) - 32]
            for k in range(4):
                scl = np.repeat(s_[:, 2 * k:2 * k + 2], 16, axis=1)            # is = l / 16
                y[:, 128 * n + 32 * k:128 * n + 32 * k + 32] = (d[:, None] * scl) * qs[k].astype(f32)
    return y.reshape(-1)


def write_model(path, arch, seed=1234, ftype=1, tok_emb_gain=1.5, with_tensors=True, qtype=None, dequantized=False,
                quantize_fn=None):
    """Write a GGML whisper model file.  ftype=1: 2-D+ weights as F16 (conv biases / positional
    embeddings / 1-D tensors stay F32, as the reference converter does); ftype=0: everything F32.
    qtype ('q4_0' | 'q4_1' | 'q5_0' | 'q5_1' | 'q8_0'): what the reference's `quantize` tool makes of the ftype=1 file --
    every 2-D tensor except the positional embeddings and conv biases becomes block-quantised, the header ftype becomes
    2000 + ggml_ftype (examples/quantize/quantize.cpp:70-176, examples/common-ggml.cpp:121-140).  dequantized=True writes
    the F16 file whose weights are exactly those blocks expanded again (the teacher file of the load-time expansion test)."""
    if qtype is not None:
        ftype = 1
    kq = qtype in KQUANT_TYPES          # K-quants: blocks come from `quantize_fn` (the reference's quantiser, see tests)
    n_vocab, n_actx, d, n_head, n_al, n_tctx, n_tl, n_mels = ARCHS[arch]
    multilingual = n_vocab >= 51865
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", GGML_MAGIC))
        types = KQUANT_TYPES if kq else QUANT_TYPES
        hdr_ftype = ftype if qtype is None or dequantized else GGML_QNT_VERSION * 1000 + types[qtype][1]
        f.write(struct.pack("<11i", n_vocab, n_actx, d, n_head, n_al, n_tctx, d, n_head, n_tl, n_mels, hdr_ftype))
        filt = mel_filters(n_mels)
        f.write(struct.pack("<2i", n_mels, N_FFT_BINS))
        f.write(filt.tobytes())
        toks = synth_vocab(multilingual)
        f.write(struct.pack("<i", len(toks)))
        for t in toks:
            f.write(struct.pack("<I", len(t)))
            f.write(t)
        if not with_tensors:
            return
        for name, shape, kind in tensor_specs(arch):
            data = _init(kind, shape, rng, tok_emb_gain)
            use_f16 = ftype == 1 and len(shape) >= 2 and kind not in ("enc_pos", "dec_pos", "bias2d")
            ttype = 1 if use_f16 else 0
            payload = data.astype(np.float16 if use_f16 else np.float32).tobytes()
            if qtype is not None and len(shape) == 2 and kind not in ("enc_pos", "dec_pos", "bias2d"):
                src = data.astype(np.float16).astype(np.float32)                              # the tool reads the F16 file
                raw = quantize_fn(src, qtype) if kq else quantize_blocks(src, qtype)
                if dequantized:
                    expand = dequantize_kblocks if kq else dequantize_blocks
                    payload = expand(raw, qtype).astype(np.float16).tobytes()
                else:
                    ttype, payload = types[qtype][0], raw
            nb = name.encode()
            f.write(struct.pack("<3i", len(shape), len(nb), ttype))
            for dim in reversed(shape):  # ggml ne[] order: fastest dimension first
                f.write(struct.pack("<i", dim))
            f.write(nb)
            f.write(payload)


def synth_pcm(n_samples, seed=7, stream=0):
    """0.1*N(0,1) + 0.3*sin(2*pi*440 t)*sin(2*pi*0.5 t), clipped to [-1,1]."""
    rng = np.random.default_rng([seed, stream])
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    x = 0.1 * rng.standard_normal(n_samples) + 0.3 * np.sin(2 * np.pi * (440.0 + 37.0 * stream) * t) * np.sin(
        2 * np.pi * 0.5 * t + 0.1 * stream)
    return np.clip(x, -1.0, 1.0).astype(np.float32)
