"""CPU: the oracle's restatement of the SURVEY.md section 8f extras (efficient channel attention, SNR-targeted noise
mixing) against fixtures generated from the reference itself (oracle/make_golden_extras.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ralenet_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extras_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _close(got, ref, tol=1e-10):
    got, ref = got.double().numpy(), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= tol * (1 + np.max(np.abs(ref))), np.max(np.abs(got - ref))


@pytest.mark.parametrize("k", [3, 5])
def test_eca_oracle_matches_reference(gold, k):
    x, w, gy = _t(gold[f"eca{k}/x"]), _t(gold[f"eca{k}/w"]), _t(gold[f"eca{k}/gy"])
    y, saved = O.eca_fwd(x, w)
    dx, dw = O.eca_bwd(gy, saved, w)
    _close(y, gold[f"eca{k}/y"])
    _close(dx, gold[f"eca{k}/dx"])
    _close(dw, gold[f"eca{k}/dw"])


def block_params(gold):
    return {str(k): _t(gold[f"blk/p/{k}"]) for k in gold["blk/keys"]}


def test_eca_block_oracle_matches_reference(gold):
    p = block_params(gold)
    x, gy = _t(gold["blk/x"]), _t(gold["blk/gy"])
    x1, asaved = O.attn_block_fwd(x, p, 8, None, 0)
    y, fsaved = O.ffn_eca_block_fwd(x1, p)
    _close(y, gold["blk/y"])
    d1, gr = O.ffn_eca_block_bwd(gy, fsaved, p)
    dx, gra = O.attn_block_bwd(d1, asaved, p, 8, None, 0)
    gr.update(gra)
    _close(dx, gold["blk/dx"])
    for k in p:
        _close(gr[k].reshape(p[k].shape), gold[f"blk/g/{k}"], 1e-9)


def test_eca_block_mirror_has_reference_state_dict_layout(gold):
    from ecg_denoise_b200.model.transformer import TransformerBlock
    blk = TransformerBlock(32, 8, local_enhence=True, use_eca=True)
    assert list(blk.state_dict().keys()) == [str(k) for k in gold["blk/keys"]]
    for k, v in blk.state_dict().items():
        assert tuple(v.shape) == gold[f"blk/p/{k}"].shape, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_snr_mix_oracle_matches_reference(gold, dtype):
    data, noise, snr = _t(gold["snr/data"]).to(dtype), _t(gold["snr/noise"]).to(dtype), _t(gold["snr/snr"])
    out = O.snr_mix(data, noise, snr)
    _close(out, gold["snr/out"], 2e-6 if dtype == torch.float32 else 1e-6)   # the reference ran in float32
    # the definition: the mixed window has the requested SNR
    d, n = data.double().reshape(6, -1), (out.double() - data.double()).reshape(6, -1)
    got = 10 * torch.log10((d ** 2).mean(1) / (n ** 2).mean(1))
    assert torch.allclose(got, snr.double(), atol=1e-3)
