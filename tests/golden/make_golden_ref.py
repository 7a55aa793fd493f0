"""Golden vectors produced by RUNNING the reference's own Python code (not a restatement).

Everything in the reference that does not need TensorFlow at call time is imported from
``/root/reference`` with stub modules for its absent third-party imports (``unidecode``, ``inflect``,
``librosa``, ``tensorflow``: only touched at import time or by functions we do not call) and executed
on seeded inputs:

* ``text.text_to_sequence2`` / ``sequence_to_text2`` / ``symbols2`` (reference ``text/__init__.py:51-62,
  83-91,106-120``, ``text/symbols.py:20-23`` with ``datasets/normal.json``) on the 21 sentences of
  ``eval.py:8-29``, every line of ``eval.txt`` and brace / unknown-symbol / duplicate-symbol cases;
* ``util.audio``: ``inv_preemphasis``, ``preemphasis``, ``_denormalize``, ``_normalize``, ``_db_to_amp``,
  ``_amp_to_db``, ``_stft_parameters``, ``find_endpoint``, ``save_wav`` (the int16 samples handed to
  ``librosa.output.write_wav``) (``util/audio.py:14-24,55-63,114-148``);
* ``hparams.hparams`` defaults (``hparams.py:5-53``);
* the feeder's batch layout helpers ``_prepare_inputs`` / ``_prepare_targets`` / ``_round_up``
  (``datasets/datafeeder_npy.py:174-195``): the layout ``Tacotron.initialize`` receives when teacher forced.

Writes ``tests/golden/ref_text.json`` and ``tests/golden/ref_audio.npz``.  ``/root/reference`` only exists in
the build container; the fixtures travel, this script documents how they were made.

    python tests/golden/make_golden_ref.py
"""
from __future__ import annotations

import ast
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TACO_REFERENCE", "/root/reference")


class _HParams:
    """stand-in for tf.contrib.training.HParams: attribute bag with values()"""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def values(self):
        return dict(self.__dict__)


def _install_stubs(captured):
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("unidecode", unidecode=lambda s: s)
    mod("inflect", engine=lambda: types.SimpleNamespace(number_to_words=lambda *a, **k: ""))
    tf = mod("tensorflow")
    tf.contrib = types.SimpleNamespace(training=types.SimpleNamespace(HParams=_HParams))

    def write_wav(path, data, sr):
        captured["wav"] = (path, np.array(data), sr)

    lib = mod("librosa", output=types.SimpleNamespace(write_wav=write_wav))
    lib.filters = mod("librosa.filters")
    lib.core = types.SimpleNamespace()
    lib.effects = mod("librosa.effects")


def main():
    captured = {}
    _install_stubs(captured)
    os.chdir(REF)   # text/symbols.py opens ./datasets/normal.json relative to the working directory
    sys.path.insert(0, REF)
    import text as rtext                      # noqa: E402  (the reference's package)
    from text.symbols import symbols2         # noqa: E402
    from hparams import hparams as rhp        # noqa: E402
    from util import audio as raudio          # noqa: E402

    # ---- text -----------------------------------------------------------------------------------------
    tree = ast.parse(open(os.path.join(REF, "eval.py"), encoding="utf-8").read())
    sentences = None
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "sentences":
            sentences = ast.literal_eval(node.value)
    assert sentences and len(sentences) == 21
    eval_txt = [l.rstrip("\n") for l in open(os.path.join(REF, "eval.txt"), encoding="utf-8")]
    extra = ["", " ", "~", "_", "abc XYZ", "你好{n i3 h ao3}世界", "{sh ang4 h ai3}", "x{zz yy}y{a1}", "1234 五六",
             "，。？！", "a{b}c{d e}f", "{unclosed", "tab\there"]
    cases = sentences + eval_txt + extra
    cleaner = ["basic_cleaners"]
    seqs = [rtext.text_to_sequence2(t, cleaner) for t in cases]
    back = [rtext.sequence_to_text2(s[:-1]) for s in seqs]
    dup = {}
    for i, s in enumerate(symbols2):
        dup.setdefault(s, []).append(i)
    dup = {s: ix for s, ix in dup.items() if len(ix) > 1}
    blob = json.dumps(symbols2, ensure_ascii=False).encode("utf-8")
    out_text = {
        "generator": "tests/golden/make_golden_ref.py (reference text/__init__.py, text/symbols.py run as is)",
        "symbols2": symbols2,
        "symbols2_len": len(symbols2),
        "symbols2_distinct": len(set(symbols2)),
        "symbols2_sha256": hashlib.sha256(blob).hexdigest(),
        "duplicates": dup,
        "eos_id": rtext._symbol_to_id["~"],
        "n_eval_py_sentences": len(sentences),
        "cases": cases,
        "sequences": seqs,
        "roundtrip": back,
    }
    with open(os.path.join(HERE, "ref_text.json"), "w", encoding="utf-8") as f:
        json.dump(out_text, f, ensure_ascii=False)

    # ---- audio helpers -----------------------------------------------------------------------------
    rng = np.random.default_rng(20261018)
    x = rng.standard_normal(5000)
    spec = rng.uniform(-0.3, 1.3, (16, 64))            # also outside [0,1]: _denormalize clips
    db = rng.uniform(-120.0, 30.0, (64,))
    amp = np.abs(rng.standard_normal(64)) * 3.0
    amp[:4] = [0.0, 1e-7, 1e-5, 1.0]
    n_fft, hop, win = raudio._stft_parameters()
    # find_endpoint: speech, then > 2 s of silence (sample_rate 20000)
    wav_ep = np.concatenate([0.5 * np.sin(np.arange(30000) * 0.05), np.zeros(70000), 0.4 * np.ones(5000)])
    ep = raudio.find_endpoint(wav_ep)
    ep_none = raudio.find_endpoint(0.5 * np.ones(120000))
    wav_save = (0.3 * rng.standard_normal(4000)).astype(np.float64)
    raudio.save_wav(wav_save.copy(), "/tmp/x.wav")
    pcm = captured["wav"][1]
    wav_quiet = 1e-3 * rng.standard_normal(1000)
    raudio.save_wav(wav_quiet.copy(), "/tmp/y.wav")
    pcm_quiet = captured["wav"][1]
    hp_values = {k: v for k, v in rhp.values().items()}

    # ---- feeder batch layout ------------------------------------------------------------------------
    from datasets import datafeeder_npy as feeder   # noqa: E402
    seq_in = [np.arange(1, n + 1, dtype=np.int32) for n in (5, 9, 3)]
    tgt_in = [rng.uniform(0, 1, (n, 4)).astype(np.float32) for n in (7, 12, 10)]
    feed_inputs = feeder._prepare_inputs(seq_in)
    feed_targets = feeder._prepare_targets(tgt_in, 5)
    round_up = np.array([[x_, m, feeder._round_up(x_, m)] for x_ in (0, 1, 4, 5, 6, 13) for m in (1, 2, 5)], np.int64)

    # ---- vocoder chain: the reference's own numpy helpers around the (TensorFlow) STFT pair, which is the oracle's ----
    # wav = inv_preemphasis(griffin_lim_tf(_db_to_amp(_denormalize(spec) + ref_level_db) ** power))   (util/audio.py:33-46)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import audio_oracle as AO             # noqa: E402  (restates tf.contrib.signal only)
    spec_gl = rng.uniform(-0.1, 1.1, (2, 9, rhp.num_freq)).astype(np.float32)
    mags_ref = raudio._db_to_amp(raudio._denormalize(spec_gl.astype(np.float64)) + rhp.ref_level_db) ** rhp.power
    wav_gl = {}
    for iters in (0, 2):
        wav_gl[iters] = np.stack([raudio.inv_preemphasis(AO.griffin_lim_tf(mags_ref[i], rhp, iters)) for i in range(2)])

    np.savez_compressed(
        os.path.join(HERE, "ref_audio.npz"),
        spec_gl=spec_gl, mags_ref=mags_ref.astype(np.float32), wav_gl0=wav_gl[0], wav_gl2=wav_gl[2],
        x=x, inv_preemphasis=raudio.inv_preemphasis(x), preemphasis=raudio.preemphasis(x),
        spec=spec, denormalize=raudio._denormalize(spec), normalize=raudio._normalize(raudio._denormalize(spec) - 7.0),
        db=db, db_to_amp=raudio._db_to_amp(db), amp=amp, amp_to_db=raudio._amp_to_db(amp),
        stft_parameters=np.array([n_fft, hop, win], np.int64),
        wav_ep=wav_ep.astype(np.float32), endpoint=np.array([raudio.find_endpoint(wav_ep.astype(np.float32)), ep, ep_none], np.int64),
        wav_save=wav_save, pcm=pcm, wav_quiet=wav_quiet, pcm_quiet=pcm_quiet,
        save_sr=np.array([captured["wav"][2]], np.int64),
        hparams_json=np.array(json.dumps(hp_values, sort_keys=True)),
        feed_targets_in_lens=np.array([len(t) for t in tgt_in], np.int64),
        feed_targets_in=np.concatenate(tgt_in), feed_targets=feed_targets, feed_inputs=feed_inputs, round_up=round_up,
    )
    print("symbols2: %d entries (%d distinct), %d duplicated symbols; %d text cases" %
          (len(symbols2), len(set(symbols2)), len(dup), len(cases)))
    print("stft parameters", (n_fft, hop, win), "endpoint", ep, ep_none)


if __name__ == "__main__":
    main()
