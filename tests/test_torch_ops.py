"""TORCH_LIBRARY(taco_b200, ...): the C ABI as PyTorch custom operators (csrc/torch_ops.cpp, torch_ops.py).
CPU: the operator library loads and registers every operator with its schema.  GPU: the operators return exactly what the
ctypes path (Engine) returns, and refuse tensors the C ABI could not read."""
import numpy as np
import pytest
import torch

from conftest import make_inputs


def test_ops_registered():
    from tacotron_multispeaker_b200 import torch_ops
    torch_ops.load()
    for name in torch_ops.OPS:
        op = getattr(torch.ops.taco_b200, name)
        schema = str(op.default._schema)
        assert schema.startswith("taco_b200::%s(int handle" % name), schema
    assert "Tensor? identities" in str(torch.ops.taco_b200.forward.default._schema)


def test_ops_have_no_cpu_kernel():
    """There is no CPU fallback: a CPU tensor is refused by the dispatcher (no kernel registered for the CPU key)."""
    from tacotron_multispeaker_b200 import torch_ops
    torch_ops.load()
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.taco_b200.postnet(1, torch.zeros(1, 5, 80), 0, 1025)


@pytest.mark.gpu
def test_ops_match_ctypes_path(small_hp, small_weights):
    from tacotron_multispeaker_b200 import torch_ops
    from tacotron_multispeaker_b200.engine import Engine
    hp = small_hp
    eng = Engine(hp, 6)
    eng.load_weights(small_weights)
    try:
        ids, lengths, spk = make_inputs(3, 17, 6, 5)
        ref = eng.forward(ids, lengths, spk)
        dev = torch.device("cuda", 0)
        ids_d, len_d, spk_d = (torch.from_numpy(x).to(dev) for x in (ids, lengths, spk))
        mel, lin, al, steps = torch_ops.forward(eng, ids_d, len_d, spk_d)
        assert steps == ref[3]
        assert torch.equal(mel, ref[0]) and torch.equal(lin, ref[1]) and torch.equal(al, ref[2])
        # stage operators: encoder -> decode -> postnet reproduce the whole path
        memory = torch.ops.taco_b200.encoder(eng.handle, ids_d, len_d, spk_d, 0)
        dec, al2, s2 = torch.ops.taco_b200.decode(eng.handle, memory, None, False, hp.num_mels, hp.outputs_per_step)
        assert s2 == steps and torch.equal(al2, ref[2])
        mel2 = dec.reshape(3, s2 * hp.outputs_per_step, hp.num_mels)
        assert torch.equal(mel2, ref[0])
        lin2 = torch.ops.taco_b200.postnet(eng.handle, mel2.contiguous(), 0, hp.num_freq)
        assert torch.equal(lin2, ref[1])
        # teacher forcing through the operator
        tg = torch.rand(3, 10, hp.num_mels, device=dev)
        ref_t = eng.forward(ids, lengths, spk, mel_targets=tg, teacher_force=True)
        mel_t, lin_t, al_t, s_t = torch_ops.forward(eng, ids_d, len_d, spk_d, mel_targets=tg, teacher_force=True)
        assert s_t == ref_t[3] and torch.equal(mel_t, ref_t[0]) and torch.equal(lin_t, ref_t[1])
        # vocoder operator
        wav = torch.ops.taco_b200.griffin_lim(eng.handle, lin.contiguous(), 3, hp.sample_rate, hp.frame_shift_ms, hp.frame_length_ms,
                                              hp.min_level_db, hp.ref_level_db, hp.power, hp.preemphasis)
        wav_ref = eng.griffin_lim(lin.contiguous(), 3)
        assert torch.equal(wav, wav_ref)
        # dtype / layout checks
        with pytest.raises(RuntimeError):
            torch_ops.forward(eng, ids_d.long(), len_d, spk_d)
        with pytest.raises(RuntimeError):
            torch.ops.taco_b200.postnet(eng.handle, mel2.transpose(1, 2), 0, hp.num_freq)
    finally:
        eng.close()
