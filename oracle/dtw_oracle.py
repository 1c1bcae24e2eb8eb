"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the alignment stage of the reference's DTW token timestamps
(whisper_exp_compute_token_level_timestamps_dtw, median_filter, dtw_and_backtrace: src/whisper.cpp:8712-8998, itself a port of
openai/whisper timing.py).  Only tests/ may import this module; the product never does.

Pinned by tests/test_gpu_dtw.py, which compares the product's t_dtw with the compiled reference end to end; this restatement
checks the host stage alone, on the CPU.
"""
import numpy as np


def dtw_first_positions(probs, n_audio, skip_front, medfilt_width=7):
    """probs [n_heads][n_tokens][T] -> for tokens skip_front .. n_tokens - 2 the first audio position of their run on the path."""
    w = probs[:, :, :n_audio].astype(np.float32)                       # [A][tokens][M]
    # ggml_norm over the token axis with eps 1e-9: sums in double, values in float (src/whisper.cpp:8930)
    mean = (w.astype(np.float64).sum(axis=1, keepdims=True) / w.shape[1]).astype(np.float32)
    z = (w - mean).astype(np.float32)
    var = ((z * z).astype(np.float64).sum(axis=1, keepdims=True) / w.shape[1]).astype(np.float32)
    z = (z * (np.float32(1.0) / np.sqrt(var + np.float32(1e-9)))).astype(np.float32)
    # median filter along time, reflect padding (8805-8840)
    hw = medfilt_width // 2
    idx = np.arange(-hw, hw + 1)[None, :] + np.arange(n_audio)[:, None]
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= n_audio, 2 * (n_audio - 1) - idx, idx)
    med = np.sort(z[:, :, idx], axis=-1)[..., medfilt_width // 2]        # [A][tokens][M]
    cost = -(med.astype(np.float64).sum(axis=0).astype(np.float32) / np.float32(med.shape[0]))
    x = cost[skip_front:-1]                                             # drop the sot sequence and eot
    N, M = x.shape
    D = np.full((N + 1, M + 1), np.inf, np.float32)
    tr = np.full((N + 1, M + 1), -1, np.int32)
    D[0, 0] = 0
    for j in range(1, M + 1):
        for i in range(1, N + 1):
            c0, c1, c2 = D[i - 1, j - 1], D[i - 1, j], D[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            D[i, j] = np.float32(x[i - 1, j - 1] + c)
            tr[i, j] = t
    tr[0, :] = 2
    tr[:, 0] = 1
    i, j = N, M
    path = []
    while i > 0 or j > 0:
        path.append((i - 1, j - 1))
        t = tr[i, j]
        if t == 0:
            i, j = i - 1, j - 1
        elif t == 1:
            i -= 1
        else:
            j -= 1
    first = [-1] * N
    for ti, tj in reversed(path):
        if first[ti] < 0:
            first[ti] = tj
    return first
