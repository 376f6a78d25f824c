"""Stand-alone operators of the path under the reference's own names and argument meaning, backed by the CUDA entry
points of include/fs2_b200.h (the forward itself does not go through these: it runs the fused two-stage path).

    from smart_nar_fast_tts_b200.operators import LengthRegulator, GaussianUpsampling, get_mask_from_lengths

replaces `from model.modules import LengthRegulator, GaussianUpsampling` (model/modules.py:162-230) and
`from utils.tools import get_mask_from_lengths` (utils/tools.py:89-97).  CUDA tensors only: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch
from torch import nn

from .capi import load_library


def _cuda(t: torch.Tensor, what: str) -> torch.device:
    if t.device.type != "cuda":
        raise RuntimeError(f"{what}: the B200 operators run on CUDA tensors only; there is no CPU fallback")
    return t.device


def get_mask_from_lengths(lengths: torch.Tensor, max_len: Optional[int] = None) -> torch.Tensor:
    """utils/tools.py:89-97: mask[b, i] = i >= lengths[b] (True = padded).  `max_len=None` reads max(lengths) back, as
    the reference does with `.item()`."""
    dev = _cuda(lengths, "get_mask_from_lengths")
    lib = load_library()
    B = lengths.shape[0]
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    lens = lengths.to(torch.long).contiguous()
    mask = torch.empty(B, int(max_len), device=dev, dtype=torch.bool)
    if B > 0 and max_len > 0:
        with torch.cuda.device(dev):
            lib.check(lib.fs2_mask_from_lengths(lens.data_ptr(), B, int(max_len), mask.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream), None)
    return mask


class LengthRegulator(nn.Module):
    """model/modules.py:195-230.  forward(x[B,L,D], duration[B,L], max_len) -> (output[B,T,D], mel_len[B] int64):
    row i of utterance b repeated max(int(duration[b,i]), 0) times, utterances zero-padded to `max_len` (None: the
    batch maximum).  One scan + one gather kernel instead of B*L `.item()` round trips."""

    def forward(self, x: torch.Tensor, duration: torch.Tensor, max_len: Optional[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = _cuda(x, "LengthRegulator")
        lib = load_library()
        B, L, D = x.shape
        x = x.float().contiguous()
        d = duration.to(device=dev, dtype=torch.float32).contiguous()
        cum = torch.empty(B, L, dtype=torch.int32, device=dev)
        mel_len = torch.empty(B, dtype=torch.long, device=dev)
        t_max = C.c_int32(0)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            lib.check(lib.fs2_duration_scan(d.data_ptr(), B, L, cum.data_ptr(), mel_len.data_ptr(), C.byref(t_max), st), None)
            T = int(t_max.value)
            if max_len is not None:
                if int(max_len) < T:
                    raise ValueError(f"max_len ({int(max_len)}) is shorter than the longest expanded utterance ({T})")
                T = int(max_len)
            out = torch.empty(B, T, D, device=dev, dtype=torch.float32)
            lib.check(lib.fs2_length_regulate(x.data_ptr(), cum.data_ptr(), B, L, D, T, out.data_ptr(), st), None)
        return out, mel_len

    LR = forward   # the reference exposes the same computation under both names (modules.py:201-218, 228-230)


class GaussianUpsampling(nn.Module):
    """model/modules.py:162-192.  forward(x[B,L,D], durations[B,L], range_outputs, max_len) -> (output[B,T,D], s[B,1],
    w[B,L,T_w]) with T_w = ceil(max_b sum durations).  `range_outputs` is accepted and ignored exactly like the
    reference, which overwrites it with the constant 10.0 (:175).  Banded evaluation: exp(-0.01 (t-c)^2) is exactly 0
    in fp32 beyond |t-c| >= 104, so each frame only visits the phonemes near it."""

    def __init__(self, return_weights: bool = True):
        super().__init__()
        self.return_weights = return_weights   # False: skip materialising w (4*L bytes per frame); w is returned as None

    def forward(self, x: torch.Tensor, durations: torch.Tensor, range_outputs=None, max_len: Optional[int] = None):
        dev = _cuda(x, "GaussianUpsampling")
        lib = load_library()
        B, L, D = x.shape
        x = x.float().contiguous()
        d = durations.to(device=dev, dtype=torch.float32).contiguous()
        # torch.arange(0, max(s)) has ceil(max(s)) elements; one read-back (the reference synchronises here too, :182)
        T_w = max(0, int(math.ceil(float(torch.sum(d, dim=-1).max().item())))) if B > 0 else 0
        T = T_w
        if max_len is not None:
            if int(max_len) < T_w:
                raise ValueError(f"max_len ({int(max_len)}) is shorter than the upsampled length ({T_w})")
            T = int(max_len)
        out = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        s = torch.empty(B, device=dev, dtype=torch.float32)
        w = torch.empty(B, L, T_w, device=dev, dtype=torch.float32) if self.return_weights else None
        if B > 0 and T == 0:
            s = torch.sum(d, dim=-1)
        elif B > 0:
            with torch.cuda.device(dev):
                lib.check(lib.fs2_gaussian_upsample(x.data_ptr(), d.data_ptr(), B, L, D, T, T_w, out.data_ptr(), s.data_ptr(),
                                                    w.data_ptr() if w is not None and w.numel() else None,
                                                    torch.cuda.current_stream(dev).cuda_stream), None)
        return out, s.unsqueeze(-1), w
