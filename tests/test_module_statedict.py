"""CPU: the host-side mirror keeps the reference's nn.Module surface: state_dict key layout, shapes, strict loading
incl. the training-only mel_encoder.* keys, and (when /root/reference is present) identical default initialisation."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import fs2_oracle as O
from helpers import ljspeech_configs


def make():
    from smart_nar_fast_tts_b200 import FastSpeech2Align
    pc, mc = ljspeech_configs(O.STATS_NAN_BINS)
    with np.errstate(invalid="ignore"):
        return FastSpeech2Align(pc, mc), pc, mc


def test_state_dict_layout():
    m, _, _ = make()
    sd = m.state_dict()
    ref = O.make_state_dict(0, include_mel_encoder=True)
    assert set(sd.keys()) == set(ref.keys()), set(sd.keys()) ^ set(ref.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    m.load_state_dict(ref, strict=True)
    # SURVEY.md section 6 parameter counts (nn.Parameters, incl. position tables and bins)
    assert sum(p.numel() for p in m.parameters()) == 41275969
    assert sum(p.numel() for k, p in m.named_parameters() if not k.startswith("mel_encoder.")) == 29385537


def test_forward_signature_matches_reference():
    import inspect
    m, _, _ = make()
    params = list(inspect.signature(m.forward).parameters)
    assert params == ["speakers", "texts", "src_lens", "max_src_len", "mels", "mel_lens", "max_mel_len", "p_targets",
                      "e_targets", "p_control", "e_control"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree only exists in the dev container")
def test_same_default_init_as_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "unidecode", "inflect"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["unidecode"].unidecode = lambda s: s
    sys.modules["inflect"].engine = lambda: None
    sys.path.insert(0, "/root/reference")
    try:
        from model import FastSpeech2Align as Ref  # type: ignore
        _, pc, mc = make()
        torch.manual_seed(123)
        with np.errstate(invalid="ignore"):
            r = Ref(pc, mc)
        torch.manual_seed(123)
        from smart_nar_fast_tts_b200 import FastSpeech2Align
        with np.errstate(invalid="ignore"):
            m = FastSpeech2Align(pc, mc)
        a, b = r.state_dict(), m.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert torch.equal(torch.nan_to_num(a[k].float()), torch.nan_to_num(b[k].float())), k
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "model" or k.startswith(("model.", "transformer", "utils.", "text"))]:
            sys.modules.pop(k, None)
