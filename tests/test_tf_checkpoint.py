"""TF V2 checkpoint (tensor bundle) reader: format constants, table round trips, corruption handling, and the
Synthesizer.load path that discovers id_num from `model/inference/embedding_id` (reference synthesizer.py:23-25)."""
import os
import struct

import numpy as np
import pytest

from tacotron_multispeaker_b200 import tf_checkpoint as T


def test_crc32c_known_answers():
    # RFC 3720 / iSCSI test vectors
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    # incremental == one shot
    assert T.crc32c(b"6789", T.crc32c(b"12345")) == 0xE3069283


def test_crc_mask_formula():
    # tensorflow/core/lib/hash/crc32c.h: ((crc >> 15) | (crc << 17)) + 0xa282ead8
    c = T.crc32c(b"foo")
    m = T.mask_crc(c)
    assert m != c and T.unmask_crc(m) == c
    assert T.mask_crc(0) == 0xA282EAD8
    assert T.unmask_crc(T.mask_crc(0xFFFFFFFF)) == 0xFFFFFFFF


def test_varint_round_trip():
    for v in (0, 1, 127, 128, 300, 2 ** 31 - 1, 2 ** 35 + 7, 2 ** 63 - 1):
        b = T._put_varint(v)
        assert T._get_varint(b, 0) == (v, len(b))
    assert T._put_varint(300) == b"\xac\x02"


def test_table_layout_and_round_trip():
    items = [(b"", b"header")] + [(("k%04d" % i).encode(), os.urandom(i % 50)) for i in range(400)]
    blob = T.build_table(items, block_size=512)          # many data blocks, restart points every 16 keys
    assert struct.unpack("<Q", blob[-8:])[0] == T.TABLE_MAGIC == 0xDB4775248B80FB57
    assert len(blob) > T.FOOTER_LEN
    assert T.read_table(blob) == items
    one = T.build_table(items[:3])                       # single block
    assert T.read_table(one) == items[:3]
    with pytest.raises(ValueError):
        T.build_table([(b"b", b""), (b"a", b"")])        # keys must increase


def test_table_detects_corruption():
    items = [(("key%03d" % i).encode(), b"v" * 20) for i in range(50)]
    blob = bytearray(T.build_table(items, block_size=256))
    with pytest.raises(ValueError):
        T.read_table(bytes(blob[:-1]))                   # magic gone
    bad = bytearray(blob)
    bad[10] ^= 0x40                                      # flip a bit inside the first data block
    with pytest.raises(ValueError):
        T.read_table(bytes(bad))
    assert len(T.read_table(bytes(bad), verify=False)) == 50   # structure still parses without the checksum
    with pytest.raises(ValueError):
        T.read_table(b"short")


def test_snappy_blocks_are_readable():
    # hand-assembled snappy stream: literal "abcd", then a copy (offset 4, length 8) that overlaps its own output
    raw = bytes([12, (4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert T._snappy_uncompress(raw) == b"abcdabcdabcd"


def test_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {
        "model/inference/embedding": rng.standard_normal((37, 16)).astype(np.float32),
        "model/inference/embedding_id": rng.standard_normal((5, 8)).astype(np.float32),
        "model/inference/prenet/dense_1/bias": np.zeros((0,), np.float32),          # empty tensor
        "model/inference/prenet/dense_1/kernel/Adam": rng.standard_normal((3, 3)).astype(np.float32),
        "global_step": np.asarray(1234, np.int64),                                   # scalar
        "some/int32": np.arange(12, dtype=np.int32).reshape(3, 4),
        "some/f64": rng.standard_normal((2, 2, 2)),
    }
    prefix = str(tmp_path / "model.ckpt-1234")
    T.write_checkpoint(prefix, tensors, block_size=128)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    r = T.CheckpointReader(prefix)
    shapes = r.get_variable_to_shape_map()
    assert shapes["model/inference/embedding_id"] == [5, 8] and shapes["global_step"] == []
    assert r.get_variable_to_dtype_map()["global_step"] == "int64"
    assert r.has_tensor("some/int32") and not r.has_tensor("nope")
    for k, v in tensors.items():
        got = r.get_tensor(k, verify=True)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    assert "model/inference/prenet/dense_1/kernel/Adam" not in r.tensors()
    with pytest.raises(KeyError):
        r.get_tensor("nope")
    # a damaged data shard is caught by the per-tensor checksum
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(r._entries["model/inference/embedding"].offset + 5)
        f.write(b"\xff")
    with pytest.raises(ValueError):
        T.CheckpointReader(prefix).get_tensor("model/inference/embedding", verify=True)
    with pytest.raises(FileNotFoundError):
        T.CheckpointReader(str(tmp_path / "missing"))


def test_latest_checkpoint_and_load_weights(tmp_path):
    d = tmp_path / "logs-tacotron"
    d.mkdir()
    w = {"model/inference/embedding_id": np.ones((4, 64), np.float32)}
    T.write_checkpoint(str(d / "model.ckpt-2000"), w)
    (d / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-2000"\nall_model_checkpoint_paths: "model.ckpt-1000"\n'
                                  'all_model_checkpoint_paths: "model.ckpt-2000"\n')
    assert T.latest_checkpoint(str(d)) == str(d / "model.ckpt-2000")
    assert T.latest_checkpoint(str(tmp_path)) is None
    for path in (str(d), str(d / "model.ckpt-2000"), str(d / "model.ckpt-2000.index")):
        got = T.load_weights(path)
        assert list(got) == ["model/inference/embedding_id"] and got["model/inference/embedding_id"].shape == (4, 64)
    np.savez(str(tmp_path / "w.npz"), **w)
    assert T.load_weights(str(tmp_path / "w.npz"))["model/inference/embedding_id"].shape == (4, 64)


def test_full_model_checkpoint_maps_onto_the_weight_specs(tmp_path):
    """A checkpoint with the reference's variable names (plus optimizer slots) canonicalizes to the model's weights."""
    from tacotron_multispeaker_b200.hparams import HParams
    from tacotron_multispeaker_b200.weights import PREFIX, canonicalize, random_init, weight_specs
    hp = HParams(outputs_per_step=5)
    w = random_init(hp, 3, seed=5)
    full = {}
    for k, v in w.items():
        name = k if k.startswith(PREFIX) else PREFIX + k
        full[name] = v
    some = next(n for n in full if n.endswith("dense_1/kernel"))
    full[some + "/Adam"] = np.zeros_like(full[some])
    full[some + "/Adam_1"] = np.zeros_like(full[some])
    full["global_step"] = np.asarray(7, np.int32)
    prefix = str(tmp_path / "model.ckpt-7")
    T.write_checkpoint(prefix, full)
    r = T.CheckpointReader(prefix)
    assert r.get_variable_to_shape_map()[PREFIX + "embedding_id"][0] == 3          # synthesizer.py:25
    got = canonicalize(r.tensors(), hp, 3)
    ref = canonicalize(w, hp, 3)
    assert set(got) == set(weight_specs(hp, 3))
    for k in ref:
        assert np.array_equal(got[k], ref[k])


@pytest.mark.gpu
def test_synthesizer_loads_a_v2_checkpoint(tmp_path):
    """Synthesizer.load(prefix) on a V2 bundle: id_num from the shape of embedding_id, same spectrogram as the same
    weights handed over directly; a single-speaker checkpoint raises KeyError like the reference (synthesizer.py:25)."""
    from tacotron_multispeaker_b200.hparams import HParams
    from tacotron_multispeaker_b200.synthesizer import Synthesizer
    from tacotron_multispeaker_b200.weights import PREFIX, random_init
    hp = HParams(outputs_per_step=5)
    w = random_init(hp, 4, seed=11)
    full = {(k if k.startswith(PREFIX) else PREFIX + k): v for k, v in w.items()}
    d = tmp_path / "logs"
    d.mkdir()
    T.write_checkpoint(str(d / "model.ckpt-10"), full)
    (d / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-10"\n')
    seq = [7110, 7200, 7300, 7150, 7111, 7222]
    a = Synthesizer(hp).load(str(d))                      # directory -> checkpoint state file -> bundle
    assert a.id_num == 4
    a.hparams.max_iters = 6
    lin_a, al_a = a.synthesize_sequence(seq, 2)
    b = Synthesizer(hp).load(None, id_num=4, seed=11)
    b.hparams.max_iters = 6
    lin_b, al_b = b.synthesize_sequence(seq, 2)
    assert lin_a.shape == lin_b.shape and np.array_equal(lin_a, lin_b) and np.array_equal(al_a, al_b)
    single = {k: v for k, v in full.items() if not k.endswith("embedding_id")}
    T.write_checkpoint(str(tmp_path / "single.ckpt"), single)
    with pytest.raises(KeyError):
        Synthesizer(hp).load(str(tmp_path / "single.ckpt"))
