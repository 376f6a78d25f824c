"""Numeric parity with the CPU oracle at BASELINE.json's own sizes, on EVERY row of every output:

  C1  1 utterance, 60 phonemes                      (configs[0])
  C3  batch 256, lengths 40..120: the bench default (configs[2]; the oracle needs ~10 s on the GPU box's host cores)
  C5  batch 64, 300-phoneme inputs, T > max_seq_len (configs[4]: 2,300-key attention, on-the-fly positional table)

in the default arithmetic (f16x2 encoder / bf16 decoder: the benchmarked one) and in the fp32-faithful one
(f16x2 / f16x2).  Gates are those of tests/test_gpu_forward.py (helpers.GATES); the measured errors are printed
(`pytest -s`) and recorded by scripts/measure_parity.py under profiles/.
C2 (configs[1]) is tests/test_gpu_forward.py::test_forward_oracle_batch32; C4 (batch 1024 sharded) is the torchrun
parity script scripts/sharded_parity.py plus the 1-GPU shard-invariance properties in tests/test_gpu_properties.py.
"""
import pytest
import torch

import fs2_oracle as O
from helpers import build_model, max_abs
from test_gpu_forward import check_against, run_model

pytestmark = pytest.mark.gpu

CONFIGS = {"c1": (1, 60, 60, 1), "c3": (256, 40, 120, 1), "c5": (64, 300, 300, 1)}


@pytest.fixture(scope="module")
def sd():
    return O.make_state_dict(0)


@pytest.fixture(scope="module", params=list(CONFIGS))
def case(request, sd):
    b, lo, hi, seed = CONFIGS[request.param]
    inputs = O.make_inputs(b, lo, hi, seed=seed)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = list(O.forward(sd, O.Dims(), *inputs)[:10])
    return request.param, inputs, ref


@pytest.mark.parametrize("dec_prec", ["bf16", "f16x2"])
def test_config_against_oracle(lib, sd, case, dec_prec):
    name, inputs, ref = case
    m = build_model(sd, O.STATS_NAN_BINS).set_precision("f16x2", dec_prec)
    out = run_model(m, *inputs)
    st = check_against(ref, out[:10], sd, dec_prec)
    print(f"\n{name} dec={dec_prec}: frames {int(ref[9].sum())} T {ref[0].shape[1]} bucket flips {st['flips']} "
          f"(utterances compared: {st['kept_utterances']} of {ref[0].shape[0]}) max|dlog_d| {max_abs(out[4], ref[4]):.2e} "
          f"mel relRMS {st['mel'][0]:.2e} max {st['mel'][1]:.2e} postnet relRMS {st['postnet_mel'][0]:.2e} max {st['postnet_mel'][1]:.2e}")
    if name == "c5":
        assert ref[0].shape[1] > 1000      # the T > max_seq_len branch (Models.py:218-225) is really exercised
