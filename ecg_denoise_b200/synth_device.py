"""Device-side batch synthesiser (csrc/synth.cu + the SNR-mix kernel): a fresh (noisy, clean) training batch per step
with zero host traffic -- the data path of SURVEY.md section 8 f3.

The reference prepares its pairs on the host (wfdb records -> np_norm -> noise mixed at a target SNR with
single_snr_noise_add, local_utils/local_utils.py:176-192, 261-266 -> .npy -> DataLoader + collate, main.py:45-60).
`DeviceSynth.fill` does the same three things on the GPU: synthetic z-normalised ECG windows, bw / ma / em style noise,
and the reference's SNR formula (the pinned `ralenet_snr_mix` kernel).  `FusedTrainer.step_synth(synth)` puts it in
front of the training step inside the same CUDA graph; the batch index is the Adam step counter, so every replay
trains on new windows.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib

KINDS = {"bw": 0, "ma": 1, "em": 2, "emb": 3}


class DeviceSynth:
    def __init__(self, B: int, leads: int = 2, length: int = 256, seed: int = 2023, kind: str = "emb",
                 snr_db: float = -4.0, device: Optional[torch.device] = None):
        self.B, self.leads, self.L, self.seed, self.kind = int(B), int(leads), int(length), int(seed), KINDS[kind]
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.noise = torch.empty(self.B, self.leads, self.L, device=self.device, dtype=torch.float32)
        self.snr = torch.full((self.B,), float(snr_db), device=self.device, dtype=torch.float32)

    def fill(self, noisy: torch.Tensor, clean: torch.Tensor, counter: Optional[torch.Tensor] = None):
        """write a batch into `noisy`, `clean` (B, leads, L) CUDA fp32; `counter`: device int32 batch index or None."""
        assert tuple(noisy.shape) == tuple(clean.shape) == (self.B, self.leads, self.L)
        lib = _lib.load()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.ralenet_synth_windows(clean.data_ptr(), self.noise.data_ptr(), self.B, self.leads, self.L,
                                             self.seed, counter.data_ptr() if counter is not None else None,
                                             self.kind, st))
        _lib.check(lib.ralenet_snr_mix(clean.data_ptr(), self.noise.data_ptr(), self.snr.data_ptr(), noisy.data_ptr(),
                                       self.B, self.leads * self.L, st))

    def batch(self, counter: Optional[torch.Tensor] = None):
        noisy = torch.empty(self.B, self.leads, self.L, device=self.device, dtype=torch.float32)
        clean = torch.empty_like(noisy)
        self.fill(noisy, clean, counter)
        return noisy, clean
