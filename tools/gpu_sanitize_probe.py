"""Small end-to-end run for compute-sanitizer (memcheck): tiny.en, jfk.wav, encoder + a few decoder steps, default and chain modes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import open_whisper_kit_b200 as pkg  # noqa: E402
from open_whisper_kit_b200 import api, modelgen  # noqa: E402

lib = pkg.load()
os.makedirs("/tmp/models", exist_ok=True)
path = "/tmp/models/tiny.en-1.bin"
if not os.path.exists(path):
    modelgen.write_model(path, "tiny.en")
pcm = api.read_wav_f32(os.path.join(ROOT, "tests", "golden", "jfk.wav"))
for mode in ("0", "1", "2"):
    os.environ["WHISPER_B200_CHAIN"] = mode
    with api.Whisper(lib, path, flash_attn=True) as w:
        assert w.pcm_to_mel(pcm) == 0 and w.encode(0) == 0
        tok = lib.whisper_token_sot(w.ctx)
        for n_past in range(4):
            rc, lg = w.decode([tok], n_past)
            assert rc == 0 and np.isfinite(lg).all()
            tok = int(lg[:50000].argmax())
        p = w.greedy_params(no_timestamps=True)
        p.max_tokens = 6
        rc, segs = w.full(p, np.concatenate([pcm, pcm])[:16000 * 40], n_processors=2)
        assert rc == 0
    print("mode", mode, "ok", flush=True)
