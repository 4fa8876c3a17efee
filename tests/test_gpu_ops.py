"""GPU parity, per kernel: every C entry point (called through the ctypes ABI via the autograd Functions)
against the CPU oracle in float64 on the same seeded inputs.

Tolerance: |d| <= rtol * (|ref| + rms(ref)), rtol = 1e-3 for the fp32 path.
The measured error is written to gpurun_out/parity_ops.json for the record.
"""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import ralenet_oracle as O
from tests.common import rel_rms_err

pytestmark = pytest.mark.gpu
RTOL = 1e-3
REPORT = {}


def _record(name, err):
    REPORT[name] = float(err)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/parity_ops.json", "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _cmp(name, got: torch.Tensor, ref: torch.Tensor, rtol=RTOL):
    assert tuple(got.shape) == tuple(ref.shape), (name, got.shape, ref.shape)
    err = rel_rms_err(got.detach().double().cpu().numpy(), ref.detach().double().cpu().numpy())
    _record(name, err)
    assert math.isfinite(err) and err <= rtol, f"{name}: rel err {err:.3e} > {rtol}"


def _rand(rs, *shape, scale=1.0):
    return torch.from_numpy(rs.standard_normal(shape) * scale)


def _block_params(rs, C, le):
    p = {
        "norm1.weight": 1 + 0.1 * _rand(rs, C), "norm1.bias": 0.1 * _rand(rs, C),
        "attn.qkv_proj.to_q.weight": _rand(rs, C, C, scale=C ** -0.5), "attn.qkv_proj.to_q.bias": 0.1 * _rand(rs, C),
        "attn.qkv_proj.to_kv.weight": _rand(rs, 2 * C, C, scale=C ** -0.5),
        "attn.qkv_proj.to_kv.bias": 0.1 * _rand(rs, 2 * C),
        "attn.proj.weight": _rand(rs, C, C, scale=C ** -0.5), "attn.proj.bias": 0.1 * _rand(rs, C),
        "norm2.weight": 1 + 0.1 * _rand(rs, C), "norm2.bias": 0.1 * _rand(rs, C),
        "mlp.fc1.weight": _rand(rs, 4 * C, C, scale=C ** -0.5), "mlp.fc1.bias": 0.1 * _rand(rs, 4 * C),
        "mlp.fc2.weight": _rand(rs, C, 4 * C, scale=(4 * C) ** -0.5), "mlp.fc2.bias": 0.1 * _rand(rs, C),
    }
    if le == 1:
        p["mlp.leconv.partial_conv3.weight"] = _rand(rs, 1, 1, 3, scale=0.5)
    elif le == 2:
        p["mlp.leconv.weight"] = _rand(rs, 4 * C, 1, 3, scale=0.5)
    return p


def _dev(t, grad=True):
    return t.float().cuda().requires_grad_(grad)


@pytest.fixture(scope="module")
def ops():
    from ecg_denoise_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("stage", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("bias", [True, False])
def test_attn_block(ops, stage, bias):
    if bias and stage == 4:
        pytest.skip("no R-wave table at stage 4")
    rs = np.random.RandomState(100 + stage)
    C, H, L = O.CHANNELS[stage], O.HEADS[stage], O.LENGTHS[stage]
    B = 3
    p = _block_params(rs, C, 0)
    W = O.RW_WINDOW[stage] if bias else 0
    table = 0.5 * _rand(rs, 2 * W - 1, H) if bias else None
    x = _rand(rs, B, L, C)
    g = _rand(rs, B, L, C)
    y_ref, saved = O.attn_block_fwd(x, p, H, table, W)
    dx_ref, gr = O.attn_block_bwd(g, saved, p, H, table, W)

    d = {k: _dev(v) for k, v in p.items()}
    xt, tt = _dev(x), (_dev(table) if bias else None)
    y = ops.AttnBlockFn.apply(xt, d["norm1.weight"], d["norm1.bias"], d["attn.qkv_proj.to_q.weight"],
                              d["attn.qkv_proj.to_q.bias"], d["attn.qkv_proj.to_kv.weight"],
                              d["attn.qkv_proj.to_kv.bias"], d["attn.proj.weight"], d["attn.proj.bias"], tt, H, W,
                              (L - W) // 2 if bias else 0, ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
    tag = f"attn/s{stage}/{'rw' if bias else 'plain'}"
    _cmp(tag + "/y", y, y_ref)
    y.backward(g.float().cuda())
    _cmp(tag + "/dx", xt.grad, dx_ref)
    for k in ("norm1.weight", "norm1.bias", "attn.qkv_proj.to_q.weight", "attn.qkv_proj.to_q.bias",
              "attn.qkv_proj.to_kv.weight", "attn.qkv_proj.to_kv.bias", "attn.proj.weight", "attn.proj.bias"):
        _cmp(f"{tag}/d_{k}", d[k].grad, gr[k])
    if bias:
        _cmp(tag + "/d_table", tt.grad, gr["table"])


@pytest.mark.parametrize("stage", [3, 4])
@pytest.mark.parametrize("mode", [0, 2])
def test_attn_fwd_paired_windows_bit_identical(ops, stage, mode):
    """At the wide stages several windows share a CTA.  mode 0 (attn.cu): large batches put two windows on a CTA and
    the odd last windows go through the one-window kernel; mode 2 (attn_umma.cu): 4 / 8 windows form a 128-token
    tile, the last tile partial.  Outputs and every saved tensor must equal, bit for bit, what the same kernels give
    on the same windows in small batches, where they sit at other positions of their CTA / tile (no window may see
    its neighbours)."""
    from ecg_denoise_b200 import _lib
    prev = _lib.set_attn_umma(mode)
    try:
        _paired_windows_bit_identical(ops, stage)
    finally:
        _lib.set_attn_umma(prev)


def _paired_windows_bit_identical(ops, stage):
    rs = np.random.RandomState(300 + stage)
    C, H, L = O.CHANNELS[stage], O.HEADS[stage], O.LENGTHS[stage]
    B = 515                                # 257 two-window CTAs + one odd window
    p = _block_params(rs, C, 0)
    W = O.RW_WINDOW[stage] if stage < 4 else 0
    table = (0.5 * _rand(rs, 2 * W - 1, H)).float().cuda() if W else None
    x = _rand(rs, B, L, C).float().cuda()
    d = {k: v.float().cuda() for k, v in p.items()}

    def run(xs):
        with torch.no_grad():
            return ops.AttnBlockFn.apply(xs, d["norm1.weight"], d["norm1.bias"], d["attn.qkv_proj.to_q.weight"],
                                         d["attn.qkv_proj.to_q.bias"], d["attn.qkv_proj.to_kv.weight"],
                                         d["attn.qkv_proj.to_kv.bias"], d["attn.proj.weight"], d["attn.proj.bias"],
                                         table, H, W, (L - W) // 2 if W else 0, ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)

    y_big = run(x)
    y_small = torch.cat([run(x[i:i + 97].contiguous()) for i in range(0, B, 97)])
    assert torch.equal(y_big, y_small)
    # training mode (saved q, k, v, o, lse) through autograd: gradients agree bit for bit as well
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    g = _rand(rs, B, L, C).float().cuda()

    def run_grad(xs):
        return ops.AttnBlockFn.apply(xs, d["norm1.weight"], d["norm1.bias"], d["attn.qkv_proj.to_q.weight"],
                                     d["attn.qkv_proj.to_q.bias"], d["attn.qkv_proj.to_kv.weight"],
                                     d["attn.qkv_proj.to_kv.bias"], d["attn.proj.weight"], d["attn.proj.bias"],
                                     table, H, W, (L - W) // 2 if W else 0, ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)

    run_grad(xa).backward(g)
    for i in range(0, B, 97):
        run_grad(xb[i:i + 97]).backward(g[i:i + 97])
    assert torch.equal(xa.grad, xb.grad)


@pytest.mark.parametrize("stage", [3, 4])
@pytest.mark.parametrize("B", [3, 515])
def test_attn_fwd_tile_kernels_match_one_window_kernels(ops, stage, B):
    """Wide stages: the tcgen05 tile kernels (attn_umma.cu: 128 tokens = 4 / 8 windows per tile, heads split over a
    cluster, partial last tile at both batch sizes) and the one-window mma.sync kernels (attn.cu) compute the same
    function: outputs and the gradients driven by their saved q, k, v, o, lse agree far inside the tolerance."""
    from ecg_denoise_b200 import _lib
    rs = np.random.RandomState(500 + stage)
    C, H, L = O.CHANNELS[stage], O.HEADS[stage], O.LENGTHS[stage]
    p = _block_params(rs, C, 0)
    W = O.RW_WINDOW[stage] if stage < 4 else 0
    table = (0.5 * _rand(rs, 2 * W - 1, H)).float().cuda() if W else None
    x = _rand(rs, B, L, C).float().cuda()
    g = _rand(rs, B, L, C).float().cuda()
    d = {k: v.float().cuda() for k, v in p.items()}
    names = ("norm1.weight", "norm1.bias", "attn.qkv_proj.to_q.weight", "attn.qkv_proj.to_q.bias",
             "attn.qkv_proj.to_kv.weight", "attn.qkv_proj.to_kv.bias", "attn.proj.weight", "attn.proj.bias")

    def run(tile_kernels):
        prev = _lib.set_attn_umma(tile_kernels)
        try:
            xs = x.clone().requires_grad_(True)
            ps = [d[k].clone().requires_grad_(True) for k in names]
            y = ops.AttnBlockFn.apply(xs, *ps, table, H, W, (L - W) // 2 if W else 0,
                                      ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
            y.backward(g)
            torch.cuda.synchronize()
            return [y.detach(), xs.grad] + [q.grad for q in ps]
        finally:
            _lib.set_attn_umma(prev)

    a, b = run(2), run(0)
    for name, ta, tb in zip(("y", "dx") + tuple("d_" + k for k in names), a, b):
        _cmp(f"attn_tile_vs_window/s{stage}/B{B}/{name}", ta, tb, rtol=5e-5)


def test_attn_plain_msattention(ops):
    """MSAttention.forward alone: no pre-norm, no residual (flags = 0)."""
    rs = np.random.RandomState(7)
    C, H, L, B = 16, 4, 128, 2
    p = _block_params(rs, C, 0)
    x = _rand(rs, B, L, C)
    u = x
    q = u @ p["attn.qkv_proj.to_q.weight"].t() + p["attn.qkv_proj.to_q.bias"]
    kv = u @ p["attn.qkv_proj.to_kv.weight"].t() + p["attn.qkv_proj.to_kv.bias"]
    qh, kh, vh = (t.reshape(B, L, H, 4).permute(0, 2, 1, 3) for t in (q, kv[..., :C], kv[..., C:]))
    o = (torch.softmax(0.5 * qh @ kh.transpose(-1, -2), -1) @ vh).permute(0, 2, 1, 3).reshape(B, L, C)
    ref = o @ p["attn.proj.weight"].t() + p["attn.proj.bias"]
    d = {k: _dev(v, False) for k, v in p.items()}
    y = ops.AttnBlockFn.apply(_dev(x, False), None, None, d["attn.qkv_proj.to_q.weight"], d["attn.qkv_proj.to_q.bias"],
                              d["attn.qkv_proj.to_kv.weight"], d["attn.qkv_proj.to_kv.bias"], d["attn.proj.weight"],
                              d["attn.proj.bias"], None, H, 0, 0, 0)
    _cmp("attn/msattention_only/y", y, ref)


@pytest.mark.parametrize("stage", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("le", [0, 1, 2])
def test_ffn_block(ops, stage, le):
    rs = np.random.RandomState(200 + stage * 3 + le)
    C, L = O.CHANNELS[stage], O.LENGTHS[stage]
    B = 3
    p = _block_params(rs, C, le)
    x = _rand(rs, B, L, C)
    g = _rand(rs, B, L, C)
    y_ref, saved = O.ffn_block_fwd(x, p)
    dx_ref, gr = O.ffn_block_bwd(g, saved, p)
    d = {k: _dev(v) for k, v in p.items()}
    lew = d.get("mlp.leconv.partial_conv3.weight", d.get("mlp.leconv.weight"))
    xt = _dev(x)
    y = ops.FFNBlockFn.apply(xt, d["norm2.weight"], d["norm2.bias"], d["mlp.fc1.weight"], d["mlp.fc1.bias"],
                             d["mlp.fc2.weight"], d["mlp.fc2.bias"], lew, None, le,
                             ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
    tag = f"ffn/s{stage}/le{le}"
    _cmp(tag + "/y", y, y_ref)
    y.backward(g.float().cuda())
    _cmp(tag + "/dx", xt.grad, dx_ref)
    for k in gr:
        _cmp(f"{tag}/d_{k}", d[k].grad, gr[k])


@pytest.mark.parametrize("stage", [2, 3, 4])
@pytest.mark.parametrize("B", [5, 37])
def test_wgrad_tcgen05_kernels_match_mma_sync_kernels(ops, stage, B):
    """Weight-gradient GEMMs of a whole block (attention: Wp, Wq, Wkv + biases; feed-forward: W1, W2 + biases, one of
    them tiled as its transpose) at token counts that are not multiples of the chunk / tile sizes: the tcgen05 kernels
    (wgrad_umma.cu) and the mma.sync kernels (wgrad.cu) agree far inside the tolerance."""
    from ecg_denoise_b200 import _lib
    rs = np.random.RandomState(700 + stage)
    C, H, L = O.CHANNELS[stage], O.HEADS[stage], O.LENGTHS[stage]
    p = _block_params(rs, C, 1)
    x = _rand(rs, B, L, C).float().cuda()
    g = _rand(rs, B, L, C).float().cuda()
    d = {k: v.float().cuda() for k, v in p.items()}
    an = ("norm1.weight", "norm1.bias", "attn.qkv_proj.to_q.weight", "attn.qkv_proj.to_q.bias",
          "attn.qkv_proj.to_kv.weight", "attn.qkv_proj.to_kv.bias", "attn.proj.weight", "attn.proj.bias")
    fn = ("norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias",
          "mlp.leconv.partial_conv3.weight")

    def run(tc):
        prev = _lib.set_wgrad_umma(tc)
        prev_reg = _lib.load().ralenet_set_wgrad_reg(0)      # this test is about the two older kernel families
        try:
            xs = x.clone().requires_grad_(True)
            pa = [d[k].clone().requires_grad_(True) for k in an]
            pf = [d[k].clone().requires_grad_(True) for k in fn]
            y = ops.AttnBlockFn.apply(xs, *pa, None, H, 0, 0, ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
            y = ops.FFNBlockFn.apply(y, *pf, None, 1, ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
            y.backward(g)
            torch.cuda.synchronize()
            return [q.grad for q in pa + pf]
        finally:
            _lib.set_wgrad_umma(prev)
            _lib.load().ralenet_set_wgrad_reg(prev_reg)

    a, b = run(True), run(False)
    for name, ta, tb in zip(an + fn, a, b):
        _cmp(f"wgrad_tc_vs_mma/s{stage}/B{B}/d_{name}", ta, tb, rtol=5e-5)


@pytest.mark.parametrize("stage", [0, 1, 2, 3])
def test_patch_merge(ops, stage):
    rs = np.random.RandomState(300 + stage)
    C, L, B = O.CHANNELS[stage], O.LENGTHS[stage], 3
    p = {"norm.weight": 1 + 0.1 * _rand(rs, 2 * C), "norm.bias": 0.1 * _rand(rs, 2 * C),
         "reduction.weight": _rand(rs, 2 * C, 2 * C, scale=(2 * C) ** -0.5)}
    x, g = _rand(rs, B, L, C), _rand(rs, B, L // 2, 2 * C)
    y_ref, saved = O.patch_merge_fwd(x, p)
    dx_ref, gr = O.patch_merge_bwd(g, saved, p)
    d = {k: _dev(v) for k, v in p.items()}
    xt = _dev(x)
    y = ops.PatchFn.apply(xt, d["norm.weight"], d["norm.bias"], d["reduction.weight"], None, 0)
    _cmp(f"pm/s{stage}/y", y, y_ref)
    y.backward(g.float().cuda())
    _cmp(f"pm/s{stage}/dx", xt.grad, dx_ref)
    for k in gr:
        _cmp(f"pm/s{stage}/d_{k}", d[k].grad, gr[k])


@pytest.mark.parametrize("stage", [1, 2, 3, 4])
@pytest.mark.parametrize("with_skip", [True, False])
def test_patch_separate(ops, stage, with_skip):
    rs = np.random.RandomState(400 + stage)
    C, L, B = O.CHANNELS[stage], O.LENGTHS[stage], 3
    p = {"norm.weight": 1 + 0.1 * _rand(rs, C // 2), "norm.bias": 0.1 * _rand(rs, C // 2),
         "reduction.weight": _rand(rs, C // 2, C // 2, scale=(C // 2) ** -0.5)}
    x, g = _rand(rs, B, L, C), _rand(rs, B, 2 * L, C // 2)
    skip = _rand(rs, B, 2 * L, C // 2) if with_skip else None
    y_ref, saved = O.patch_separate_fwd(x, p, skip)
    dx_ref, gr = O.patch_separate_bwd(g, saved, p)
    d = {k: _dev(v) for k, v in p.items()}
    xt = _dev(x)
    st = _dev(skip) if with_skip else None
    y = ops.PatchFn.apply(xt, d["norm.weight"], d["norm.bias"], d["reduction.weight"], st, 1)
    tag = f"ps/s{stage}/{'skip' if with_skip else 'noskip'}"
    _cmp(tag + "/y", y, y_ref)
    y.backward(g.float().cuda())
    _cmp(tag + "/dx", xt.grad, dx_ref)
    if with_skip:
        _cmp(tag + "/dskip", st.grad, g)
    for k in gr:
        _cmp(f"{tag}/d_{k}", d[k].grad, gr[k])


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("L", [256, 512])
def test_stem(ops, training, L):
    rs = np.random.RandomState(500 + L)
    B = 5
    p = {"conv1.0.weight": _rand(rs, 8, 2, 3, scale=0.4), "conv1.0.bias": 0.1 * _rand(rs, 8),
         "conv1.2.weight": 1 + 0.1 * _rand(rs, 8), "conv1.2.bias": 0.1 * _rand(rs, 8),
         "conv1.2.running_mean": 0.1 * _rand(rs, 8), "conv1.2.running_var": 0.5 + torch.from_numpy(rs.uniform(size=8))}
    x, g = _rand(rs, B, 2, L), _rand(rs, B, L, 8)
    y_ref, saved, new_stats = O.stem_fwd(x, p, training)
    dx_ref, gr = O.stem_bwd(g, saved, p)
    d = {k: _dev(v, not k.startswith("conv1.2.running")) for k, v in p.items()}
    nbt = torch.tensor(3, dtype=torch.int64, device="cuda")
    xt = _dev(x)
    y = ops.StemFn.apply(xt, d["conv1.0.weight"], d["conv1.0.bias"], d["conv1.2.weight"], d["conv1.2.bias"],
                         d["conv1.2.running_mean"], d["conv1.2.running_var"], nbt, training, 0.1, 1e-5, None)
    tag = f"stem/{'train' if training else 'eval'}/L{L}"
    _cmp(tag + "/y", y, y_ref)
    if training:
        _cmp(tag + "/running_mean", d["conv1.2.running_mean"], new_stats[0])
        _cmp(tag + "/running_var", d["conv1.2.running_var"], new_stats[1])
        assert int(nbt.item()) == 4
    else:
        assert int(nbt.item()) == 3
    y.backward(g.float().cuda())
    _cmp(tag + "/dx", xt.grad, dx_ref)
    for k in gr:
        _cmp(f"{tag}/d_{k}", d[k].grad, gr[k])


def test_head(ops):
    rs = np.random.RandomState(600)
    B, L = 4, 256
    p = {"transconv.0.weight": _rand(rs, 2, 8, 3, scale=0.2), "transconv.0.bias": 0.1 * _rand(rs, 2)}
    x, skip, dout = _rand(rs, B, L, 8), _rand(rs, B, L, 8), _rand(rs, B, 2, L)
    out_ref, saved = O.head_fwd(x, skip, p)
    ds_ref, gr = O.head_bwd(dout, saved, p)
    d = {k: _dev(v) for k, v in p.items()}
    xt, st = _dev(x), _dev(skip)
    out = ops.HeadFn.apply(xt, st, d["transconv.0.weight"], d["transconv.0.bias"])
    _cmp("head/out", out, out_ref)
    out.backward(dout.float().cuda())
    _cmp("head/dx", xt.grad, ds_ref)
    _cmp("head/dskip", st.grad, ds_ref)
    for k in gr:
        _cmp(f"head/d_{k}", d[k].grad, gr[k])


@pytest.mark.parametrize("mma", [1, 0])
@pytest.mark.parametrize("ci,co,act", [(12, 6, True), (6, 2, True), (2, 6, True), (6, 12, False)])
def test_conv13(ops, ci, co, act, mma):
    """the four lead-mixing convolutions of newrale: implicit-GEMM tensor-core kernels (conv_mma.cu, mma = 1) and the
    scalar kernels (stem_head.cu, mma = 0), both against the fp64 oracle, at a ragged batch."""
    from ecg_denoise_b200 import _lib
    prev = _lib.load().ralenet_set_conv_mma(mma)
    try:
        _conv13_case(ops, ci, co, act, f"{'mma' if mma else 'fma'}")
    finally:
        _lib.load().ralenet_set_conv_mma(prev)


def _conv13_case(ops, ci, co, act, kind):
    rs = np.random.RandomState(700 + ci)
    B, L = 37, 256
    w, b = _rand(rs, co, ci, 13, scale=(13 * ci) ** -0.5), 0.1 * _rand(rs, co)
    x, dy = _rand(rs, B, ci, L), _rand(rs, B, co, L)
    c = O.conv1d_fwd(x, w, b)
    y_ref = O.leaky_relu(c, 0.01) if act else c
    dc = dy * torch.where(c > 0, torch.ones_like(c), torch.full_like(c, 0.01)) if act else dy
    dx_ref, dw_ref, db_ref = O.conv1d_bwd(dc, x, w)
    xt, wt, bt = _dev(x), _dev(w), _dev(b)
    y = ops.Conv1dFn.apply(xt, wt, bt, act, 0.01)
    tag = f"conv13/{kind}/{ci}to{co}"
    _cmp(tag + "/y", y, y_ref)
    y.backward(dy.float().cuda())
    _cmp(tag + "/dx", xt.grad, dx_ref)
    _cmp(tag + "/dw", wt.grad, dw_ref)
    _cmp(tag + "/db", bt.grad, db_ref)


def test_mse_and_metrics(ops):
    rs = np.random.RandomState(800)
    pred, tgt = _rand(rs, 9, 2, 256), _rand(rs, 9, 2, 256)
    loss_ref, dout_ref = O.mse_loss_fwd_bwd(pred, tgt)
    loss, dout, rmse, snr = ops.mse_loss_metrics(pred.float().cuda(), tgt.float().cuda())
    _cmp("mse/loss", loss, loss_ref.reshape(1))
    _cmp("mse/dout", dout, dout_ref)
    _cmp("mse/rmse", rmse, O.RMSE(tgt, pred))
    assert torch.allclose(snr.double().cpu(), O.SNR(tgt, pred), atol=1e-3)     # dB
    # weighted variant reduces exactly to MSE at w == 1
    loss_w, dout_w, _, _ = ops.mse_loss_metrics(pred.float().cuda(), tgt.float().cuda(),
                                                weight=torch.ones(512, device="cuda"))
    assert torch.equal(dout_w, dout)
    assert torch.allclose(loss_w, loss, rtol=1e-6)        # the loss is summed with fp32 atomics across CTAs


def test_adam_flat(ops):
    rs = np.random.RandomState(900)
    n = 10007
    p0, g1, g2 = _rand(rs, n), _rand(rs, n), _rand(rs, n)
    p, m, v = p0.clone(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    O.adam_step(p, g1, m, v, 1)
    O.adam_step(p, g2, m, v, 2)
    n_pad = (n + 3) // 4 * 4
    pd = torch.zeros(n_pad, device="cuda"); pd[:n] = p0.float().cuda()
    md, vd = torch.zeros_like(pd), torch.zeros_like(pd)
    gd = torch.zeros_like(pd)
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    for gg in (g1, g2):
        gd[:n] = gg.float().cuda()
        ops.adam_flat(pd, gd, md, vd, step)
    assert int(step.item()) == 2
    _cmp("adam/dev_step/p", pd[:n], p, rtol=1e-5)
    pd2 = torch.zeros(n_pad, device="cuda"); pd2[:n] = p0.float().cuda()
    md2, vd2 = torch.zeros_like(pd2), torch.zeros_like(pd2)
    for i, gg in enumerate((g1, g2)):
        gd[:n] = gg.float().cuda()
        ops.adam_flat(pd2, gd, md2, vd2, i + 1)
    _cmp("adam/host_step/p", pd2[:n], p, rtol=1e-5)


def test_errors_are_loud(ops):
    from ecg_denoise_b200 import _lib
    with pytest.raises(_lib.RalenetError):
        ops.mse_loss_metrics(torch.zeros(2, 2, 256), torch.zeros(2, 2, 256))            # CPU tensors
    with pytest.raises(_lib.RalenetError):
        ops.PatchFn.apply(torch.zeros(1, 100, 8, device="cuda"), torch.ones(16, device="cuda"),
                          torch.zeros(16, device="cuda"), torch.zeros(16, 16, device="cuda"), None, 0)  # bad shape


# ------------------------------------------------------------------------------------------------
# SURVEY.md section 8f extras: efficient channel attention, SNR-targeted noise mixing
def _extras():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extras_golden.npz"))


@pytest.mark.parametrize("k", [3, 5])
@pytest.mark.parametrize("with_res", [False, True])
def test_eca_vs_reference_golden(ops, k, with_res):
    gold = _extras()
    x, w, gy = (torch.from_numpy(gold[f"eca{k}/{n}"]) for n in ("x", "w", "gy"))
    xt, wt = _dev(x), _dev(w)
    res = _dev(_rand(np.random.RandomState(k), *x.shape)) if with_res else None
    y = ops.EcaFn.apply(xt, wt, res)
    y_ref = torch.from_numpy(gold[f"eca{k}/y"]) + (res.detach().double().cpu() if with_res else 0)
    tag = f"eca/k{k}/{'res' if with_res else 'plain'}"
    _cmp(tag + "/y", y, y_ref)
    y.backward(gy.float().cuda())
    _cmp(tag + "/dx", xt.grad, torch.from_numpy(gold[f"eca{k}/dx"]))
    _cmp(tag + "/dw", wt.grad, torch.from_numpy(gold[f"eca{k}/dw"]))
    if with_res:
        _cmp(tag + "/dres", res.grad, gy)


def test_transformer_block_with_eca_vs_reference_golden():
    """TransformerBlock(use_eca=True) through the module mirror: the gate sits between fc2 and the residual."""
    from ecg_denoise_b200.model.transformer import TransformerBlock
    gold = _extras()
    blk = TransformerBlock(32, 8, local_enhence=True, use_eca=True)
    sd = {str(k): torch.from_numpy(gold[f"blk/p/{k}"]).float() for k in gold["blk/keys"]}
    blk.load_state_dict(sd, strict=True)
    blk = blk.cuda()
    x = _dev(torch.from_numpy(gold["blk/x"]))
    y = blk(x)
    _cmp("eca_block/y", y, torch.from_numpy(gold["blk/y"]))
    y.backward(torch.from_numpy(gold["blk/gy"]).float().cuda())
    _cmp("eca_block/dx", x.grad, torch.from_numpy(gold["blk/dx"]))
    for k, p in blk.named_parameters():
        _cmp(f"eca_block/d_{k}", p.grad, torch.from_numpy(gold[f"blk/g/{k}"]))


def test_snr_mix_vs_reference_golden(ops):
    gold = _extras()
    data, noise, snr = (torch.from_numpy(gold[f"snr/{n}"]).cuda() for n in ("data", "noise", "snr"))
    out = ops.snr_mix(data, noise, snr)
    _cmp("snr_mix/out", out, torch.from_numpy(gold["snr/out"]), rtol=1e-5)
    # scalar SNR for the whole batch, and a benchmark-sized batch: the mixed windows have the requested SNR
    rs = np.random.RandomState(11)
    d = _rand(rs, 4096, 2, 256).float().cuda()
    n = _rand(rs, 4096, 2, 256, scale=0.3).float().cuda()
    o = ops.snr_mix(d, n, -2.0)
    got = 10 * torch.log10((d.double() ** 2).mean((1, 2)) / ((o.double() - d.double()) ** 2).mean((1, 2)))
    assert torch.allclose(got, torch.full_like(got, -2.0), atol=1e-3)


# ------------------------------------------------------------------------------------------------------
# round 2: R_pos, weighted loss with a non-uniform weight, tcgen05 kernels at a large batch against the oracle
EXTRAS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extras_golden.npz")


@pytest.mark.parametrize("tag", ["rpos", "rpos_edge"])
def test_rpos_block_vs_reference_golden(tag):
    """RelativePositionEmbedding.forward(R_pos) (model/transformer.py:542-543) through the module API: the bias
    block at offset R_pos - W//2 (c0 of the C ABI), TransformerBlock forward + backward incl. the table gradient,
    against fixtures generated by the reference itself (oracle/make_golden_extras.py)."""
    from ecg_denoise_b200.model.transformer import RelativePositionEmbedding, TransformerBlock
    gold = np.load(EXTRAS)
    blk = TransformerBlock(16, 4, local_enhence=True)
    blk.load_state_dict({str(k): torch.from_numpy(gold[f"{tag}/p/{k}"]).float() for k in gold[f"{tag}/keys"]},
                        strict=True)
    rw = RelativePositionEmbedding(16, 128, 4)
    rw.relative_position_bias_table.data = torch.from_numpy(gold[f"{tag}/table"]).float()
    blk, rw = blk.cuda(), rw.cuda()
    x = torch.from_numpy(gold[f"{tag}/x"]).float().cuda().requires_grad_(True)
    mask = rw(int(gold[f"{tag}/R_pos"]))
    _cmp(f"{tag}/dense_mask", mask, torch.from_numpy(gold[f"{tag}/dense_mask"]), 1e-6)
    y = blk(x, mask)
    _cmp(f"{tag}/y", y, torch.from_numpy(gold[f"{tag}/y"]))
    y.backward(torch.from_numpy(gold[f"{tag}/gy"]).float().cuda())
    _cmp(f"{tag}/dx", x.grad, torch.from_numpy(gold[f"{tag}/dx"]))
    _cmp(f"{tag}/d_table", rw.relative_position_bias_table.grad, torch.from_numpy(gold[f"{tag}/d_table"]))
    for k, p in blk.named_parameters():
        _cmp(f"{tag}/d_{k}", p.grad, torch.from_numpy(gold[f"{tag}/g/{k}"]))


def test_weighted_mse_nonuniform_weight(ops):
    """the R-wave-weighted loss extension with w != 1: value and gradient against the fp64 oracle
    (loss = mean(w (pred - target)^2), SURVEY F4: no reference counterpart; w == 1 is F.mse_loss)."""
    rs = np.random.RandomState(801)
    pred, tgt = _rand(rs, 9, 2, 256), _rand(rs, 9, 2, 256)
    centre = np.exp(-0.5 * ((np.arange(256) - 128) / 12.0) ** 2)           # emphasise the R wave at the window centre
    w = torch.from_numpy(np.concatenate([1 + 4 * centre, 0.5 + 2 * centre]))
    loss_ref, dout_ref = O.weighted_mse_loss_fwd_bwd(pred, tgt, w)
    loss, dout, rmse, snr = ops.mse_loss_metrics(pred.float().cuda(), tgt.float().cuda(), weight=w.float().cuda())
    _cmp("wmse/loss", loss, loss_ref.reshape(1), 1e-5)
    _cmp("wmse/dout", dout, dout_ref, 1e-5)
    _cmp("wmse/rmse", rmse, O.RMSE(tgt, pred))                              # the metrics stay unweighted


@pytest.mark.parametrize("wgrad_reg", [0, 1])
@pytest.mark.parametrize("stage", [3, 4])
def test_tcgen05_block_kernels_vs_oracle_large_batch(ops, stage, wgrad_reg):
    """(wgrad_reg = 1: the weight gradients come from the register-tile kernel of wgrad_reg.cu instead of wgrad_umma.cu)
    B = 515 windows (a partial last 128-token tile, > one wave of tile CTAs) with the tcgen05 tile kernels forced
    on (attn_umma.cu, ffn_umma.cu, wgrad_umma.cu): outputs, dx and EVERY parameter gradient directly against the fp64
    oracle -- not against the mma.sync kernels."""
    from ecg_denoise_b200 import _lib
    rs = np.random.RandomState(900 + stage)
    C, H, L = O.CHANNELS[stage], O.HEADS[stage], O.LENGTHS[stage]
    B = 515
    p = _block_params(rs, C, 1)
    W = O.RW_WINDOW[stage] if stage < 4 else 0
    table = 0.5 * _rand(rs, 2 * W - 1, H) if W else None
    x, g = _rand(rs, B, L, C), _rand(rs, B, L, C)
    x1_ref, asaved = O.attn_block_fwd(x, p, H, table, W)
    y_ref, fsaved = O.ffn_block_fwd(x1_ref, p)
    d1_ref, gr = O.ffn_block_bwd(g, fsaved, p)
    dx_ref, gra = O.attn_block_bwd(d1_ref, asaved, p, H, table, W)
    gr.update(gra)
    prev_a, prev_w = _lib.set_attn_umma(2), _lib.set_wgrad_umma(True)
    prev_r = _lib.load().ralenet_set_wgrad_reg(2 * wgrad_reg)
    try:
        d = {k: _dev(v) for k, v in p.items()}
        xt, tt = _dev(x), (_dev(table) if W else None)
        x1 = ops.AttnBlockFn.apply(xt, d["norm1.weight"], d["norm1.bias"], d["attn.qkv_proj.to_q.weight"],
                                   d["attn.qkv_proj.to_q.bias"], d["attn.qkv_proj.to_kv.weight"],
                                   d["attn.qkv_proj.to_kv.bias"], d["attn.proj.weight"], d["attn.proj.bias"], tt, H, W,
                                   (L - W) // 2 if W else 0, ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
        y = ops.FFNBlockFn.apply(x1, d["norm2.weight"], d["norm2.bias"], d["mlp.fc1.weight"], d["mlp.fc1.bias"],
                                 d["mlp.fc2.weight"], d["mlp.fc2.bias"], d["mlp.leconv.partial_conv3.weight"], None, 1,
                                 ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
        y.backward(g.float().cuda())
        torch.cuda.synchronize()
    finally:
        _lib.set_attn_umma(prev_a)
        _lib.set_wgrad_umma(prev_w)
        _lib.load().ralenet_set_wgrad_reg(prev_r)
    tag = f"tcgen05_vs_oracle/s{stage}/B{B}/reg{wgrad_reg}"
    _cmp(tag + "/x1", x1, x1_ref)
    _cmp(tag + "/y", y, y_ref)
    _cmp(tag + "/dx", xt.grad, dx_ref)
    for k in p:
        _cmp(f"{tag}/d_{k}", d[k].grad, gr[k].reshape(p[k].shape))
    if W:
        _cmp(tag + "/d_table", tt.grad, gr["table"])


def test_comm_kernels_two_ranks_on_one_gpu():
    """The exchange kernels of comm.cu (BatchNorm statistic sums; gradient reduce-scatter + broadcast fused with Adam)
    with TWO ranks emulated on one GPU: two "symmetric" buffers in the same memory, each rank's kernels on its own
    stream, spinning on each other's flags exactly as two GPUs would over NVLink (peer-pointer variant; the multimem
    variant needs a real multicast mapping and is covered by tests/dp_check.py on >= 2 GPUs).  Checks: sums equal and
    bit-identical on both ranks, Adam trajectory == torch.optim.Adam on the summed gradient, three consecutive calls
    (epoch counters / flag reuse)."""
    from ecg_denoise_b200.comm import SymmComm
    dev = torch.device("cuda", torch.cuda.current_device())
    n = 4 * 25_003
    floats = SymmComm.nbytes(n) // 4
    bufs = [torch.zeros(floats, device=dev) for _ in range(2)]
    ptrs = [b.data_ptr() for b in bufs]
    comms = [SymmComm(n, dev, ptrs, r, bufs[r]) for r in range(2)]
    for c in comms:
        c.grid = 24
    streams = [torch.cuda.Stream() for _ in range(2)]
    g = torch.Generator(device="cpu").manual_seed(5)
    p0 = torch.randn(n, generator=g)
    ps = [p0.clone().to(dev) for _ in range(2)]
    ms = [torch.zeros(n, device=dev) for _ in range(2)]
    vs = [torch.zeros(n, device=dev) for _ in range(2)]
    steps = [torch.zeros(1, device=dev, dtype=torch.int32) for _ in range(2)]
    ref_p = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    for it in range(3):
        vals = [torch.randn(17, generator=g).to(dev) for _ in range(2)]
        want = (vals[0].double() + vals[1].double()).cpu()
        grads = [torch.randn(n, generator=g) * (1 + it) for _ in range(2)]
        for r in range(2):
            comms[r].grad.copy_(grads[r].to(dev))
        torch.cuda.synchronize()
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                comms[r].exchange(it % 2, vals[r])
                comms[r].allreduce_adam(ps[r], ms[r], vs[r], steps[r], 1e-3, (0.9, 0.999), 1e-8, 0.5)
        torch.cuda.synchronize()
        assert torch.equal(vals[0], vals[1])
        _cmp(f"comm/exchange/{it}", vals[0], want, 1e-6)
        assert torch.equal(comms[0].grad, comms[1].grad)                 # both ranks hold the same sum
        gsum = grads[0].double() + grads[1].double()
        _cmp(f"comm/sum/{it}", comms[0].grad, gsum, 1e-6)
        opt.zero_grad()
        ref_p.grad = 0.5 * gsum
        opt.step()
        assert torch.equal(ps[0], ps[1])
        _cmp(f"comm/adam_p/{it}", ps[0], ref_p.detach(), 1e-5)
        assert int(steps[0].item()) == it + 1


def test_standalone_helper_module_forwards():
    """LinearProjection.forward, AbsPositionalEncoding.forward, PartialConv_1d.forward and drop_path() called on their
    own, as the reference's modules allow (model/transformer.py:226-247, 179-181, 54-59, 62-79): same values and
    gradients as the plain formulas in float64."""
    from ecg_denoise_b200.model import transformer as T
    rs = np.random.RandomState(77)
    B, L, C, H = 3, 64, 32, 8
    lp = T.LinearProjection(C, H, 4).cuda()
    x = _rand(rs, B, L, C)
    xt = _dev(x)
    q, k, v = lp(xt)
    wq, bq = lp.to_q.weight.detach().double().cpu(), lp.to_q.bias.detach().double().cpu()
    wkv, bkv = lp.to_kv.weight.detach().double().cpu(), lp.to_kv.bias.detach().double().cpu()
    q_ref = (x @ wq.t() + bq).reshape(B, L, H, 4).permute(0, 2, 1, 3)
    kv_ref = (x @ wkv.t() + bkv).reshape(B, L, 2, H, 4).permute(2, 0, 3, 1, 4)
    _cmp("standalone/linproj/q", q, q_ref)
    _cmp("standalone/linproj/k", k, kv_ref[0])
    _cmp("standalone/linproj/v", v, kv_ref[1])
    gq, gk, gv = _rand(rs, B, H, L, 4), _rand(rs, B, H, L, 4), _rand(rs, B, H, L, 4)
    (q * gq.float().cuda()).sum().add((k * gk.float().cuda()).sum()).add((v * gv.float().cuda()).sum()).backward()
    dq2 = gq.permute(0, 2, 1, 3).reshape(B * L, C)
    dkv2 = torch.stack([gk, gv]).permute(1, 3, 0, 2, 4).reshape(B * L, 2 * C)
    _cmp("standalone/linproj/dx", xt.grad, (dq2 @ wq + dkv2 @ wkv).reshape(B, L, C))
    _cmp("standalone/linproj/d_wq", lp.to_q.weight.grad, dq2.t() @ x.reshape(-1, C))
    _cmp("standalone/linproj/d_bkv", lp.to_kv.bias.grad, dkv2.sum(0))
    # AbsPositionalEncoding
    pe = T.AbsPositionalEncoding(C)
    y = pe(_dev(x, False))
    _cmp("standalone/abs_pe", y, x + pe.P[0, :L].double(), 1e-6)
    # PartialConv_1d on a channels-first tensor
    pc = T.PartialConv_1d(4 * C, 4 * C, "split_cat").cuda()
    xc = _rand(rs, B, 4 * C, L)
    xct = _dev(xc)
    yc = pc(xct)
    w = pc.partial_conv3.weight.detach().double().cpu()
    ref = xc.clone()
    ref[:, :1] = torch.nn.functional.conv1d(xc[:, :1], w, padding=1)
    _cmp("standalone/pconv/y", yc, ref, 1e-5)
    g = _rand(rs, B, 4 * C, L)
    yc.backward(g.float().cuda())
    xr = xc.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr = torch.cat([torch.nn.functional.conv1d(xr[:, :1], wr, padding=1), xr[:, 1:]], 1)
    yr.backward(g)
    _cmp("standalone/pconv/dx", xct.grad, xr.grad, 1e-5)
    _cmp("standalone/pconv/dw", pc.partial_conv3.weight.grad, wr.grad, 1e-5)
    # drop_path: identity whenever the reference's would be
    t = torch.ones(2, 3, device="cuda")
    assert T.drop_path(t, 0.0, True) is t and T.drop_path(t, 0.3, False) is t
    with pytest.raises(NotImplementedError):
        T.drop_path(t, 0.3, True)


@pytest.mark.parametrize("kind", ["bw", "ma", "em", "emb"])
def test_device_synth_batch_properties(kind):
    """device-side batch synthesiser (synth.cu + snr_mix): z-normalised clean leads (np_norm), R peak of the middle
    beat at L/2 +- 8, zero-mean noise, the mixed window has exactly the requested SNR (single_snr_noise_add's
    definition, local_utils/local_utils.py:176-192), deterministic in (seed, counter), new windows per counter."""
    from ecg_denoise_b200.synth_device import DeviceSynth
    B, L = 64, 256
    sy = DeviceSynth(B, 2, L, seed=7, kind=kind, snr_db=-4.0)
    ctr = torch.zeros(1, device="cuda", dtype=torch.int32)
    noisy, clean = sy.batch(ctr)
    noisy2, clean2 = sy.batch(ctr)
    assert torch.equal(noisy, noisy2) and torch.equal(clean, clean2)
    ctr += 1
    noisy3, clean3 = sy.batch(ctr)
    assert not torch.equal(clean, clean3)
    c = clean.double().cpu()
    assert torch.isfinite(noisy).all() and torch.isfinite(clean).all()
    assert float(c.mean(-1).abs().max()) < 1e-4
    assert float((c.std(-1, unbiased=False) - 1).abs().max()) < 1e-3
    lead0 = c[:, 0]
    centre = lead0[:, L // 2 - 9:L // 2 + 10].max(-1).values
    assert bool((centre >= 0.7 * lead0.max(-1).values).all())          # an R wave sits at the window centre
    n = (noisy.double().cpu() - c).reshape(B, -1)
    assert float(n.reshape(B, 2, L).mean(-1).abs().max()) < 1e-3 * float(n.abs().max())
    snr = 10 * torch.log10((c.reshape(B, -1) ** 2).mean(1) / (n ** 2).mean(1))
    assert float((snr + 4.0).abs().max()) < 1e-3


@pytest.mark.parametrize("graph", [False, True])
def test_fused_trainer_step_synth(graph):
    """FusedTrainer.step_synth: the generator runs inside the step (inside its CUDA graph when use_graph), each step
    sees the batch of its own Adam step index, and the result equals step() on that same batch."""
    from ecg_denoise_b200.engine import FusedTrainer
    from ecg_denoise_b200.model import transformer
    from ecg_denoise_b200.synth_device import DeviceSynth
    from oracle import synth_weights as SW
    sd = SW.make_state_dict("rw", 1, 9)
    B = 16
    sy = DeviceSynth(B, 2, 256, seed=3, kind="emb", snr_db=0.0)

    def fresh():
        m = transformer.ralenet(high_level_enhence=True)
        m.load_state_dict(sd)
        return m.cuda()

    ma, mb = fresh(), fresh()
    ta, tb = FusedTrainer(ma, use_graph=graph), FusedTrainer(mb, use_graph=False)
    ctr = torch.zeros(1, device="cuda", dtype=torch.int32)
    for it in range(3):
        la = ta.step_synth(sy)[0].item()
        x, t = sy.batch(ctr)                    # the batch step `it` must have drawn (counter = steps done so far)
        lb = tb.step(x, t)[0].item()
        ctr += 1
        assert abs(la - lb) <= 1e-5 * abs(lb), (it, la, lb)
    # the two runs differ only by the summation order of the atomically accumulated weight gradients; Adam turns that
    # noise into +-lr moves of elements whose gradient is ~0, so parameters are compared with that bound (3 steps x
    # lr) and the gradients of the last step with the usual tolerance
    _cmp(f"step_synth/graph{int(graph)}/flat_grad", ma._plan.flat_grad, mb._plan.flat_grad, RTOL)
    for (n, p), (_, q) in zip(ma.named_parameters(), mb.named_parameters()):
        assert float((p.detach() - q.detach()).abs().max()) <= 3.2e-3, n


@pytest.mark.parametrize("graph", [False, True])
def test_fused_trainer_step_host_with_prefetch(graph):
    """FusedTrainer.prefetch_host / step_host: the batch staged on the copy stream one step ahead is the batch the step
    trains on (and a step_host without a matching prefetch copies from the host itself); equals step() on the same
    batches."""
    from ecg_denoise_b200.engine import FusedTrainer
    from ecg_denoise_b200.model import transformer
    from oracle import synth_weights as SW
    sd = SW.make_state_dict("rw", 1, 9)
    B, NB = 16, 3
    g = torch.Generator().manual_seed(5)
    hx = [torch.randn(B, 2, 256, generator=g).pin_memory() for _ in range(NB)]
    ht = [torch.randn(B, 2, 256, generator=g).pin_memory() for _ in range(NB)]

    def fresh():
        m = transformer.ralenet(high_level_enhence=True)
        m.load_state_dict(sd)
        return m.cuda()

    ma, mb = fresh(), fresh()
    ta, tb = FusedTrainer(ma, use_graph=graph), FusedTrainer(mb, use_graph=False)
    ta.prefetch_host(hx[0], ht[0])
    for it in range(5):
        i = it % NB
        la_dev = ta.step_host(hx[i], ht[i])
        assert ta._staged_key is None                      # the staged copy (or, at it == 3, the host copy) was consumed
        if it != 2:                                        # step 3 runs without a prefetch: direct host copy
            ta.prefetch_host(hx[(it + 1) % NB], ht[(it + 1) % NB])
        la = la_dev.item()
        lb = tb.step(hx[i].cuda(), ht[i].cuda())[0].item()
        assert abs(la - lb) <= 1e-5 * abs(lb), (it, la, lb)
    _cmp(f"step_host_prefetch/graph{int(graph)}/flat_grad", ma._plan.flat_grad, mb._plan.flat_grad, RTOL)


@pytest.mark.parametrize("stage", [2, 3, 4])
@pytest.mark.parametrize("B", [1, 5, 37])
def test_wgrad_register_tile_kernel_vs_oracle(ops, stage, B):
    """wgrad_reg.cu (operands straight from global memory into mma.sync fragments, register accumulators, token
    sub-ranges folded in shared memory, vector reductions into dW): every weight / bias gradient of a block against the
    fp64 oracle at token counts that are not multiples of the 8-token step or of the CTA slices, incl. one window."""
    from ecg_denoise_b200 import _lib
    rs = np.random.RandomState(1200 + 10 * stage + B)
    C, H, L = O.CHANNELS[stage], O.HEADS[stage], O.LENGTHS[stage]
    p = _block_params(rs, C, 1)
    x, g = _rand(rs, B, L, C), _rand(rs, B, L, C)
    x1_ref, asaved = O.attn_block_fwd(x, p, H, None, 0)
    y_ref, fsaved = O.ffn_block_fwd(x1_ref, p)
    d1_ref, gr = O.ffn_block_bwd(g, fsaved, p)
    dx_ref, gra = O.attn_block_bwd(d1_ref, asaved, p, H, None, 0)
    gr.update(gra)
    prev = _lib.load().ralenet_set_wgrad_reg(2)          # 2: force it for every group (default 1: small dW only)
    try:
        d = {k: _dev(v) for k, v in p.items()}
        xt = _dev(x)
        x1 = ops.AttnBlockFn.apply(xt, d["norm1.weight"], d["norm1.bias"], d["attn.qkv_proj.to_q.weight"],
                                   d["attn.qkv_proj.to_q.bias"], d["attn.qkv_proj.to_kv.weight"],
                                   d["attn.qkv_proj.to_kv.bias"], d["attn.proj.weight"], d["attn.proj.bias"], None, H, 0,
                                   0, ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
        y = ops.FFNBlockFn.apply(x1, d["norm2.weight"], d["norm2.bias"], d["mlp.fc1.weight"], d["mlp.fc1.bias"],
                                 d["mlp.fc2.weight"], d["mlp.fc2.bias"], d["mlp.leconv.partial_conv3.weight"], None, 1,
                                 ops.RL_F_PRENORM | ops.RL_F_RESIDUAL)
        y.backward(g.float().cuda())
        torch.cuda.synchronize()
    finally:
        _lib.load().ralenet_set_wgrad_reg(prev)
    for k in p:
        _cmp(f"wgrad_reg/s{stage}/B{B}/d_{k}", d[k].grad, gr[k].reshape(p[k].shape))
