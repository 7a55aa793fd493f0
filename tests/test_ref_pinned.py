"""Parity against outputs of the REFERENCE'S OWN CODE (not the oracle restatement).

``tests/golden/ref_text.json`` and ``tests/golden/ref_audio.npz`` were written by
``tests/golden/make_golden_ref.py``, which imports ``/root/reference``'s ``text`` package, ``util/audio.py``,
``hparams.py`` and ``datasets/datafeeder_npy.py`` (third-party imports stubbed) and runs them.  What stays
unpinned against TensorFlow itself is listed in DESIGN.md section 2.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from tacotron_multispeaker_b200.hparams import HParams

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref_text():
    with open(os.path.join(GOLD, "ref_text.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def ref_audio():
    return np.load(os.path.join(GOLD, "ref_audio.npz"))


# ---- text front-end (reference text/__init__.py:51-62,83-91,106-120; text/symbols.py:20-23) ------------------------
def test_symbol_table_matches_reference(ref_text):
    from tacotron_multispeaker_b200 import text
    from tacotron_multispeaker_b200.hparams import NUM_SYMBOLS2
    table = text.load_symbols(ref_text["symbols2"][2:])
    assert table == ref_text["symbols2"]
    assert len(table) == ref_text["symbols2_len"] == NUM_SYMBOLS2 == 7352    # models/tacotron.py:40
    assert len(set(table)) == ref_text["symbols2_distinct"] == 7321
    blob = json.dumps(table, ensure_ascii=False).encode("utf-8")
    assert hashlib.sha256(blob).hexdigest() == ref_text["symbols2_sha256"]
    assert text.EOS_ID == ref_text["eos_id"] == 1
    # later duplicates win in the reference's dict comprehension (text/__init__.py:11)
    for sym, ids in ref_text["duplicates"].items():
        assert text._symbol_to_id2[sym] == ids[-1], sym
    assert text._symbol_to_id2[" "] == 7351 and text._symbol_to_id2["~"] == 7348


def test_text_to_sequence2_matches_reference(ref_text):
    from tacotron_multispeaker_b200 import text
    text.load_symbols(ref_text["symbols2"][2:])
    assert ref_text["n_eval_py_sentences"] == 21
    assert len(ref_text["cases"]) == len(ref_text["sequences"]) >= 150
    for case, want, back in zip(ref_text["cases"], ref_text["sequences"], ref_text["roundtrip"]):
        got = text.text_to_sequence2(case, ["basic_cleaners"])
        assert got == want, case
        assert text.sequence_to_text2(got[:-1]) == back, case
    # ids the synthesizer feeds (synthesizer.py:39 strips the EOS) stay inside the embedding table
    assert max(max(s) for s in ref_text["sequences"]) < 7352


# ---- hparams defaults (reference hparams.py:5-53) -----------------------------------------------------------------
def test_hparams_defaults_match_reference(ref_audio):
    ref = json.loads(str(ref_audio["hparams_json"]))
    ours = HParams().values()
    for name in ("cleaners", "num_mels", "num_freq", "sample_rate", "frame_length_ms", "frame_shift_ms", "preemphasis",
                 "min_level_db", "ref_level_db", "outputs_per_step", "max_iters", "griffin_lim_iters", "power",
                 "embedding_text_channels", "embedding_id_channels", "bucket_len", "eos"):
        assert ours[name] == ref[name], name


# ---- audio helpers (reference util/audio.py:14-24,55-63,114-151) -----------------------------------------------------
def test_audio_host_helpers_match_reference(ref_audio):
    from tacotron_multispeaker_b200 import audio
    hp = HParams()
    assert list(audio._stft_parameters(hp)) == list(ref_audio["stft_parameters"]) == [2048, 250, 1000]
    np.testing.assert_allclose(audio._db_to_amp(ref_audio["db"]), ref_audio["db_to_amp"], rtol=1e-15)
    eps = [audio.find_endpoint(ref_audio["wav_ep"], hp.sample_rate),
           audio.find_endpoint(ref_audio["wav_ep"].astype(np.float64), hp.sample_rate),
           audio.find_endpoint(0.5 * np.ones(120000), hp.sample_rate)]
    assert eps[0] == ref_audio["endpoint"][0] and eps[2] == ref_audio["endpoint"][2]
    # save_wav: the int16 samples the reference hands to librosa (peak normalisation with the 0.01 floor, truncation)
    for wav, pcm in (("wav_save", "pcm"), ("wav_quiet", "pcm_quiet")):
        got = audio.wav_to_int16(ref_audio[wav])
        assert got.dtype == np.int16
        assert np.abs(got.astype(np.int32) - ref_audio[pcm].astype(np.int32)).max() <= 1   # float32 vs float64 scaling
    assert int(ref_audio["save_sr"][0]) == hp.sample_rate


def test_audio_oracle_matches_reference_helpers(ref_audio):
    """the numpy halves of the vocoder oracle against the reference's own functions"""
    from oracle import audio_oracle as AO
    hp = HParams()
    np.testing.assert_allclose(AO.denormalize(ref_audio["spec"], hp), ref_audio["denormalize"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(AO.db_to_amp(ref_audio["db"]), ref_audio["db_to_amp"], rtol=1e-15)
    np.testing.assert_allclose(AO.inv_preemphasis(ref_audio["x"], hp), ref_audio["inv_preemphasis"], rtol=0, atol=1e-10)
    assert list(AO.stft_parameters(hp)) == list(ref_audio["stft_parameters"])
    np.testing.assert_array_equal(AO.save_wav_int16(ref_audio["wav_save"]), ref_audio["pcm"])
    mags = AO.db_to_amp(AO.denormalize(ref_audio["spec_gl"].astype(np.float64), hp) + hp.ref_level_db) ** hp.power
    np.testing.assert_allclose(mags, ref_audio["mags_ref"], rtol=1e-6)
    for iters, key in ((0, "wav_gl0"), (2, "wav_gl2")):
        got = np.stack([AO.synthesize_wav(ref_audio["spec_gl"][i], hp, iters) for i in range(2)])
        np.testing.assert_allclose(got, ref_audio[key], rtol=0, atol=1e-9 * np.abs(ref_audio[key]).max())


def test_feeder_batch_layout_matches_reference(ref_audio):
    """teacher-forced targets are padded to max_len + 1 rounded up to a multiple of r with zeros, ids with 0
    (reference datasets/datafeeder_npy.py:174-195) -- the layout conftest.make_inputs / taco_max_steps assume."""
    from tacotron_multispeaker_b200.engine import pad_targets, pad_inputs
    lens = ref_audio["feed_targets_in_lens"]
    flat = ref_audio["feed_targets_in"]
    tg, o = [], 0
    for n in lens:
        tg.append(flat[o:o + n]); o += n
    np.testing.assert_array_equal(pad_targets(tg, 5), ref_audio["feed_targets"])
    assert ref_audio["feed_targets"].shape[1] % 5 == 0
    np.testing.assert_array_equal(pad_inputs([np.arange(1, n + 1, dtype=np.int32) for n in (5, 9, 3)]), ref_audio["feed_inputs"])
    for x, m, want in ref_audio["round_up"]:
        assert -(-int(x) // int(m)) * int(m) == want


# ---- GPU vocoder against the reference's helpers + the TF-STFT oracle ------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("iters,key,tol", [(0, "wav_gl0", 2e-5), (2, "wav_gl2", 2e-4)])
def test_gpu_vocoder_matches_reference_chain(ref_audio, iters, key, tol):
    """taco_griffin_lim (denormalise -> dB->amp -> **power -> Griffin-Lim -> inverse pre-emphasis on the device) against
    reference util/audio.py:23-24,138-151 run as is around the oracle's restatement of tf.contrib.signal's STFT pair."""
    import torch
    from tacotron_multispeaker_b200.engine import Engine
    hp = HParams()
    eng = Engine(hp, 0)
    try:
        wav = eng.griffin_lim(torch.from_numpy(ref_audio["spec_gl"]).cuda(), iters, inv_preemphasis=True)
        torch.cuda.synchronize()
        want = ref_audio[key]
        err = np.abs(wav.cpu().numpy().astype(np.float64) - want).max() / np.abs(want).max()
        assert err < tol, err
    finally:
        eng.close()
