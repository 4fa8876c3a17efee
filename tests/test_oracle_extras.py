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


@pytest.mark.parametrize("tag", ["rpos", "rpos_edge"])
def test_rpos_block_oracle_matches_reference(gold, tag):
    """RelativePositionEmbedding.forward(R_pos) (model/transformer.py:542-543): the W x W bias block sits at offset
    R_pos - W//2 instead of the centre; a whole TransformerBlock driven by that mask, forward and backward."""
    p = {str(k): _t(gold[f"{tag}/p/{k}"]) for k in gold[f"{tag}/keys"]}
    x, gy, table = _t(gold[f"{tag}/x"]), _t(gold[f"{tag}/gy"]), _t(gold[f"{tag}/table"])
    W, L, H = 16, 128, 4
    c0 = int(gold[f"{tag}/R_pos"]) - W // 2
    _close(O.rw_bias_dense(table, W, L, c0).unsqueeze(0), gold[f"{tag}/dense_mask"])
    x1, asaved = O.attn_block_fwd(x, p, H, table, W, c0)
    y, fsaved = O.ffn_block_fwd(x1, p)
    _close(y, gold[f"{tag}/y"])
    d1, gr = O.ffn_block_bwd(gy, fsaved, p)
    dx, gra = O.attn_block_bwd(d1, asaved, p, H, table, W, c0)
    gr.update(gra)
    _close(dx, gold[f"{tag}/dx"])
    _close(gr["table"], gold[f"{tag}/d_table"], 1e-9)
    for k in p:
        _close(gr[k].reshape(p[k].shape), gold[f"{tag}/g/{k}"], 1e-9)


def test_rpos_mirror_builds_the_reference_mask(gold):
    """the module mirror's RelativePositionEmbedding.forward(R_pos) returns the reference's dense mask (CPU: the
    mask itself is plain torch; the kernels read the table and the offset instead)."""
    from ecg_denoise_b200.model.transformer import RelativePositionEmbedding
    rw = RelativePositionEmbedding(16, 128, 4).double()
    rw.relative_position_bias_table.data = _t(gold["rpos/table"])
    for tag in ("rpos", "rpos_edge"):
        m = rw(int(gold[f"{tag}/R_pos"]))
        _close(m.detach(), gold[f"{tag}/dense_mask"])
        assert m._rw_src[1] == int(gold[f"{tag}/R_pos"]) - 8


def test_records_to_windows_oracle_matches_reference_np_norm_and_cut(gold):
    """per-lead z-normalisation = the reference's np_norm (local_utils.py:261-266, population std, no epsilon) and the
    cut = its `signal[i:i+256, :]` over `range(0, T, 256)` (:53), both taken from the reference's own text."""
    x = _t(gold["records/x"])
    assert list(gold["records/cut"]) == [0, 650000, 256, 256]
    win, _ = O.records_to_windows(x.double(), 256, 256)
    _close(win, gold["records/windows"], 1e-12)
    win32, _ = O.records_to_windows(x, 256, 256)
    _close(win32, gold["records/windows"], 2e-5)


def test_weighted_loss_reduces_to_mse_and_matches_autograd():
    g = torch.Generator().manual_seed(3)
    pred = torch.randn(5, 2, 256, generator=g, dtype=torch.float64, requires_grad=True)
    tgt = torch.randn(5, 2, 256, generator=g, dtype=torch.float64)
    w = torch.rand(512, generator=g, dtype=torch.float64) * 3 + 0.1
    l1, d1 = O.weighted_mse_loss_fwd_bwd(pred.detach(), tgt, torch.ones(512, dtype=torch.float64))
    l0, d0 = O.mse_loss_fwd_bwd(pred.detach(), tgt)
    assert torch.equal(l1, l0) and torch.equal(d1, d0)
    lw, dw = O.weighted_mse_loss_fwd_bwd(pred.detach(), tgt, w)
    ref = (w.view(1, 2, 256) * (pred - tgt) ** 2).mean()
    ref.backward()
    _close(lw.reshape(1), ref.detach().reshape(1).numpy())
    _close(dw, pred.grad.numpy())
