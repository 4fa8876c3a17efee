"""CPU oracle for the RA-LENet forward/backward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement, in plain torch-CPU tensor arithmetic (used as an
array library: no nn.Module, and the *_bwd functions use no autograd), of the algorithm in
the reference `caprilovel/ECG_Denoise`:

  * model/transformer.py        (RA-LENet with R-wave bias)           -> variant "rw"
  * model/raletransformer.py    (no R-wave ablation, Sequential keys) -> variant "nra"
  * model/ralenet_12leads.py    (copy of "rw" + `newrale` wrapper)    -> `newrale_*`
  * local_utils/evaluate.py     (SNR / RMSE)
  * denoise_train.py:24,53      (Adam lr 1e-3 defaults, F.mse_loss)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import it, and only as the checker / the timed CPU baseline -- never from the
product package `ecg_denoise_b200` (which fails loudly when its CUDA library is missing).

Parity pin: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the *imported reference itself*, generated in the build
container by `oracle/make_golden.py` and committed under `tests/golden/` (forward outputs,
loss, every parameter gradient, a 3-step Adam trajectory).  `tests/test_oracle_golden.py`
checks the oracle against those fixtures on CPU.

Every function cites the reference lines it restates (paths relative to the reference root).
All tensors are token-major `(B, L, C)` inside the network, like the reference after its
`rearrange(x, 'b c l -> b l c')` (model/transformer.py:630).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------
# network geometry  (model/transformer.py:566-568)
# ----------------------------------------------------------------------------------------
CHANNELS = [8, 16, 32, 64, 128]
HEADS = [2, 4, 8, 16, 32]
LENGTHS = [256, 128, 64, 32, 16]
RW_WINDOW = [32, 16, 8, 4]          # model/transformer.py:576-579
HEAD_DIM = 4
LN_EPS = 1e-5
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# (layer attribute name, stage index, rwattn index or None)  in forward order
# model/transformer.py:632-661 -- note the reference's typo'd names are part of the contract.
LAYERS = [
    ("dtransformer1", 0, 1), ("dtransformer2", 1, 2), ("dtransformer3", 2, 3),
    ("dtransformer34", 3, 4), ("transformer", 4, None), ("utransformer4", 4, None),
    ("utranformer3", 3, 4), ("utransformer2", 2, 3), ("utransformer1", 1, 2),
]


def block_prefix(variant: str, layer: str, i: int) -> str:
    """state_dict prefix of block i of a layer: BasicLayer.blocks (transformer.py:469) vs
    nn.Sequential (raletransformer.py:574-577)."""
    return f"{layer}.blocks.{i}." if variant == "rw" else f"{layer}.{i}."


# ----------------------------------------------------------------------------------------
# elementary pieces
# ----------------------------------------------------------------------------------------
def pos_encoding(L: int, C: int, dtype=torch.float32) -> Tensor:
    """AbsPositionalEncoding.P[:, :L]  (model/transformer.py:172-177).  The table is built in
    fp32 by the reference, so build it in fp32 and cast."""
    X = torch.arange(L, dtype=torch.float32).reshape(-1, 1) / torch.pow(
        10000, torch.arange(0, C, 2, dtype=torch.float32) / C)
    P = torch.zeros(L, C, dtype=torch.float32)
    P[:, 0::2] = torch.sin(X)
    P[:, 1::2] = torch.cos(X)
    return P.to(dtype)


def gelu(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (model/transformer.py:139)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def gelu_grad(x: Tensor) -> Tensor:
    return 0.5 * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0)))) + \
        x * torch.exp(-0.5 * x * x) * (1.0 / math.sqrt(2.0 * math.pi))


def layer_norm_fwd(z: Tensor, w: Tensor, b: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """nn.LayerNorm over the last dim, eps 1e-5, biased variance."""
    mu = z.mean(-1, keepdim=True)
    var = ((z - mu) ** 2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + LN_EPS)
    zh = (z - mu) * rstd
    return zh * w + b, zh, rstd


def layer_norm_bwd(du: Tensor, zh: Tensor, rstd: Tensor, w: Tensor):
    dzh = du * w
    dz = rstd * (dzh - dzh.mean(-1, keepdim=True) - zh * (dzh * zh).mean(-1, keepdim=True))
    red = tuple(range(du.dim() - 1))
    return dz, (du * zh).sum(red), du.sum(red)


def rw_bias_dense(table: Tensor, W: int, L: int, c0: Optional[int] = None) -> Tensor:
    """RelativePositionEmbedding.forward + mask_fill (model/transformer.py:534-558):
    bias[h, c0+i, c0+j] = table[i - j + W - 1, h] on the W x W block at offset c0, 0 elsewhere.
    Returns (H, L, L)."""
    H = table.shape[1]
    if c0 is None:
        c0 = (L - W) // 2
    idx = torch.arange(W).view(-1, 1) - torch.arange(W).view(1, -1) + (W - 1)
    blk = table[idx.reshape(-1)].view(W, W, H).permute(2, 0, 1)
    out = table.new_zeros(H, L, L)
    out[:, c0:c0 + W, c0:c0 + W] = blk
    return out


# ----------------------------------------------------------------------------------------
# attention half of TransformerBlock   (model/transformer.py:383-390, 289-323, 226-247)
# ----------------------------------------------------------------------------------------
def attn_block_fwd(x: Tensor, p: Dict[str, Tensor], H: int,
                   table: Optional[Tensor] = None, W: int = 0, c0: Optional[int] = None):
    """x1 = x + proj(softmax(0.5 q k^T + bias) v),  q,k,v from LN1(x*sqrt(C) + P).
    p keys: norm1.weight/bias, attn.qkv_proj.to_q.weight/bias, attn.qkv_proj.to_kv.weight/bias,
    attn.proj.weight/bias."""
    B, L, C = x.shape
    z = x * math.sqrt(C) + pos_encoding(L, C, x.dtype)               # :386, :180
    u, zh, rstd = layer_norm_fwd(z, p["norm1.weight"], p["norm1.bias"])  # :387
    q = u @ p["attn.qkv_proj.to_q.weight"].t() + p["attn.qkv_proj.to_q.bias"]      # :243
    kv = u @ p["attn.qkv_proj.to_kv.weight"].t() + p["attn.qkv_proj.to_kv.bias"]   # :244
    k, v = kv[..., :C], kv[..., C:]
    qh = q.view(B, L, H, HEAD_DIM).permute(0, 2, 1, 3)
    kh = k.reshape(B, L, H, HEAD_DIM).permute(0, 2, 1, 3)
    vh = v.reshape(B, L, H, HEAD_DIM).permute(0, 2, 1, 3)
    s = (qh * HEAD_DIM ** -0.5) @ kh.transpose(-2, -1)                # :299-300
    if table is not None:
        s = s + rw_bias_dense(table, W, L, c0).unsqueeze(0)          # :306
    pr = torch.softmax(s, -1)                                         # :307/:310
    oh = pr @ vh                                                      # :316
    o = oh.permute(0, 2, 1, 3).reshape(B, L, C)                       # :318
    a = o @ p["attn.proj.weight"].t() + p["attn.proj.bias"]           # :320
    x1 = x + a                                                        # :405
    saved = dict(zh=zh, rstd=rstd, u=u, qh=qh, kh=kh, vh=vh, pr=pr, o=o)
    return x1, saved


def attn_block_bwd(g: Tensor, saved, p: Dict[str, Tensor], H: int,
                   table: Optional[Tensor] = None, W: int = 0, c0: Optional[int] = None):
    """Manual backward of attn_block_fwd.  Returns dx and a dict of parameter grads
    (plus 'table' when a bias table is given)."""
    B, L, C = g.shape
    zh, rstd, u = saved["zh"], saved["rstd"], saved["u"]
    qh, kh, vh, pr, o = saved["qh"], saved["kh"], saved["vh"], saved["pr"], saved["o"]
    grads = {}
    g2 = g.reshape(-1, C)
    grads["attn.proj.weight"] = g2.t() @ o.reshape(-1, C)
    grads["attn.proj.bias"] = g2.sum(0)
    do = g @ p["attn.proj.weight"]
    doh = do.view(B, L, H, HEAD_DIM).permute(0, 2, 1, 3)
    oh = o.view(B, L, H, HEAD_DIM).permute(0, 2, 1, 3)
    D = (doh * oh).sum(-1, keepdim=True)
    dp = doh @ vh.transpose(-2, -1)
    ds = pr * (dp - D)
    scale = HEAD_DIM ** -0.5
    dqh = scale * (ds @ kh)
    dkh = scale * (ds.transpose(-2, -1) @ qh)
    dvh = pr.transpose(-2, -1) @ doh
    if table is not None:
        if c0 is None:
            c0 = (L - W) // 2
        dsc = ds[:, :, c0:c0 + W, c0:c0 + W].sum(0)                    # (H, W, W)
        idx = (torch.arange(W).view(-1, 1) - torch.arange(W).view(1, -1) + (W - 1)).reshape(-1)
        dt = torch.zeros_like(table)
        dt.index_add_(0, idx, dsc.permute(1, 2, 0).reshape(W * W, H))
        grads["table"] = dt
    dq = dqh.permute(0, 2, 1, 3).reshape(B, L, C)
    dk = dkh.permute(0, 2, 1, 3).reshape(B, L, C)
    dv = dvh.permute(0, 2, 1, 3).reshape(B, L, C)
    dkv = torch.cat([dk, dv], -1)
    u2 = u.reshape(-1, C)
    grads["attn.qkv_proj.to_q.weight"] = dq.reshape(-1, C).t() @ u2
    grads["attn.qkv_proj.to_q.bias"] = dq.reshape(-1, C).sum(0)
    grads["attn.qkv_proj.to_kv.weight"] = dkv.reshape(-1, 2 * C).t() @ u2
    grads["attn.qkv_proj.to_kv.bias"] = dkv.reshape(-1, 2 * C).sum(0)
    du = dq @ p["attn.qkv_proj.to_q.weight"] + dkv @ p["attn.qkv_proj.to_kv.weight"]
    dz, dgam, dbet = layer_norm_bwd(du, zh, rstd, p["norm1.weight"])
    grads["norm1.weight"], grads["norm1.bias"] = dgam, dbet
    dx = g + math.sqrt(C) * dz
    return dx, grads


# ----------------------------------------------------------------------------------------
# feed-forward half   (model/transformer.py:392-395, 149-161, 54-59)
# ----------------------------------------------------------------------------------------
LE_NONE, LE_PARTIAL, LE_DEPTHWISE = 0, 1, 2


def _fir3(x: Tensor, w: Tensor) -> Tensor:
    """zero-padded 3-tap cross-correlation along dim 1 (tokens): y[t] = sum_k w[k] x[t+k-1].
    x: (B, L, n), w: (n, 3)  (Conv1d(k=3, pad=1, bias=False), transformer.py:36 / :146)."""
    xp = torch.nn.functional.pad(x, (0, 0, 1, 1))
    return w[:, 0] * xp[:, :-2] + w[:, 1] * xp[:, 1:-1] + w[:, 2] * xp[:, 2:]


def _fir3_T(dy: Tensor, w: Tensor) -> Tensor:
    """adjoint of _fir3 w.r.t. x."""
    dp = torch.nn.functional.pad(dy, (0, 0, 1, 1))
    return w[:, 0] * dp[:, 2:] + w[:, 1] * dp[:, 1:-1] + w[:, 2] * dp[:, :-2]


def le_mode_of(p: Dict[str, Tensor]) -> int:
    if "mlp.leconv.partial_conv3.weight" in p:
        return LE_PARTIAL
    if "mlp.leconv.weight" in p:
        return LE_DEPTHWISE
    return LE_NONE


def ffn_block_fwd(x1: Tensor, p: Dict[str, Tensor]):
    """y = x1 + fc2(GELU(leconv(GELU(fc1(LN2(x1))))))   (LE)   or   x1 + fc2(GELU(fc1(LN2(x1)))).
    The shipped 'local enhancement' is a 1-channel conv on hidden channel 0."""
    mode = le_mode_of(p)
    u, zh, rstd = layer_norm_fwd(x1, p["norm2.weight"], p["norm2.bias"])       # :393
    h = u @ p["mlp.fc1.weight"].t() + p["mlp.fc1.bias"]                        # :150
    g1 = gelu(h)                                                               # :151
    if mode == LE_PARTIAL:
        w = p["mlp.leconv.partial_conv3.weight"].reshape(1, 3)
        f = torch.cat([_fir3(g1[..., :1], w), g1[..., 1:]], -1)                # :56-58
        g2 = gelu(f)                                                           # :156
    elif mode == LE_DEPTHWISE:
        w = p["mlp.leconv.weight"].reshape(-1, 3)
        f = _fir3(g1, w)                                                       # :146
        g2 = gelu(f)
    else:
        f, g2 = None, g1
    y = x1 + g2 @ p["mlp.fc2.weight"].t() + p["mlp.fc2.bias"]                  # :158, :410
    return y, dict(zh=zh, rstd=rstd, u=u, h=h, g1=g1, f=f, g2=g2, mode=mode)


def ffn_block_bwd(g: Tensor, saved, p: Dict[str, Tensor]):
    C = g.shape[-1]
    mode = saved["mode"]
    zh, rstd, u, h, g1, f, g2 = (saved[k] for k in ("zh", "rstd", "u", "h", "g1", "f", "g2"))
    grads = {}
    gf = g.reshape(-1, C)
    grads["mlp.fc2.weight"] = gf.t() @ g2.reshape(-1, 4 * C)
    grads["mlp.fc2.bias"] = gf.sum(0)
    dg2 = g @ p["mlp.fc2.weight"]
    if mode == LE_PARTIAL:
        w = p["mlp.leconv.partial_conv3.weight"].reshape(1, 3)
        df = dg2 * gelu_grad(f)
        df0, g10 = df[..., :1], g1[..., :1]
        g1p = torch.nn.functional.pad(g10, (0, 0, 1, 1))
        dw = torch.stack([(df0 * g1p[:, :-2]).sum(), (df0 * g1p[:, 1:-1]).sum(),
                          (df0 * g1p[:, 2:]).sum()])
        grads["mlp.leconv.partial_conv3.weight"] = dw.view(1, 1, 3)
        dg1 = torch.cat([_fir3_T(df0, w), df[..., 1:]], -1)
    elif mode == LE_DEPTHWISE:
        w = p["mlp.leconv.weight"].reshape(-1, 3)
        df = dg2 * gelu_grad(f)
        g1p = torch.nn.functional.pad(g1, (0, 0, 1, 1))
        dw = torch.stack([(df * g1p[:, :-2]).sum((0, 1)), (df * g1p[:, 1:-1]).sum((0, 1)),
                          (df * g1p[:, 2:]).sum((0, 1))], -1)
        grads["mlp.leconv.weight"] = dw.view(-1, 1, 3)
        dg1 = _fir3_T(df, w)
    else:
        dg1 = dg2
    dh = dg1 * gelu_grad(h)
    grads["mlp.fc1.weight"] = dh.reshape(-1, 4 * C).t() @ u.reshape(-1, C)
    grads["mlp.fc1.bias"] = dh.reshape(-1, 4 * C).sum(0)
    du = dh @ p["mlp.fc1.weight"]
    dz, dgam, dbet = layer_norm_bwd(du, zh, rstd, p["norm2.weight"])
    grads["norm2.weight"], grads["norm2.bias"] = dgam, dbet
    return g + dz, grads


# ----------------------------------------------------------------------------------------
# PatchMerging / PatchSeparate   (model/transformer.py:440-460, 418-424)
# ----------------------------------------------------------------------------------------
def patch_merge_fwd(x: Tensor, p: Dict[str, Tensor]):
    """cat(x[:,0::2], x[:,1::2], -1) is exactly x viewed as (B, L/2, 2C) -> LN(2C) -> Linear."""
    B, L, C = x.shape
    xm = x.reshape(B, L // 2, 2 * C)
    u, zh, rstd = layer_norm_fwd(xm, p["norm.weight"], p["norm.bias"])
    return u @ p["reduction.weight"].t(), dict(zh=zh, rstd=rstd, u=u)


def patch_merge_bwd(g: Tensor, saved, p: Dict[str, Tensor]):
    B, Lh, C2 = g.shape
    grads = {"reduction.weight": g.reshape(-1, C2).t() @ saved["u"].reshape(-1, C2)}
    du = g @ p["reduction.weight"]
    dz, dgam, dbet = layer_norm_bwd(du, saved["zh"], saved["rstd"], p["norm.weight"])
    grads["norm.weight"], grads["norm.bias"] = dgam, dbet
    return dz.reshape(B, Lh * 2, C2 // 2), grads


def patch_separate_fwd(x: Tensor, p: Dict[str, Tensor], skip: Optional[Tensor] = None):
    """'b l (c1 c2) -> b (c1 l) c2', c1=2: rows 0..L-1 take channels [0,C/2), rows L..2L-1 take
    [C/2,C)  -> LN(C/2) -> Linear(C/2,C/2,no bias)  (+ U-skip, transformer.py:650/654/658)."""
    B, L, C = x.shape
    xs = torch.cat([x[..., :C // 2], x[..., C // 2:]], 1)
    u, zh, rstd = layer_norm_fwd(xs, p["norm.weight"], p["norm.bias"])
    y = u @ p["reduction.weight"].t()
    if skip is not None:
        y = y + skip
    return y, dict(zh=zh, rstd=rstd, u=u)


def patch_separate_bwd(g: Tensor, saved, p: Dict[str, Tensor]):
    """returns dx (B, L, C); the skip gradient is g itself."""
    B, L2, Ch = g.shape
    grads = {"reduction.weight": g.reshape(-1, Ch).t() @ saved["u"].reshape(-1, Ch)}
    du = g @ p["reduction.weight"]
    dz, dgam, dbet = layer_norm_bwd(du, saved["zh"], saved["rstd"], p["norm.weight"])
    grads["norm.weight"], grads["norm.bias"] = dgam, dbet
    L = L2 // 2
    return torch.cat([dz[:, :L], dz[:, L:]], -1), grads


# ----------------------------------------------------------------------------------------
# Conv1d helpers (channels-first tensors), stem, head, 12-lead wrapper convs
# ----------------------------------------------------------------------------------------
def conv1d_fwd(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """'same' cross-correlation, odd kernel: y[b,o,t] = b[o] + sum_{i,k} w[o,i,k] x[b,i,t+k-pad]."""
    Bn, Ci, L = x.shape
    Co, _, K = w.shape
    pad = K // 2
    xp = torch.nn.functional.pad(x, (pad, pad))
    y = x.new_zeros(Bn, Co, L)
    for k in range(K):
        y = y + torch.einsum("oi,bil->bol", w[:, :, k], xp[:, :, k:k + L])
    if b is not None:
        y = y + b.view(1, -1, 1)
    return y


def conv1d_bwd(dy: Tensor, x: Tensor, w: Tensor):
    """returns dx, dw, db."""
    Bn, Ci, L = x.shape
    Co, _, K = w.shape
    pad = K // 2
    xp = torch.nn.functional.pad(x, (pad, pad))
    dyp = torch.nn.functional.pad(dy, (pad, pad))
    dw = torch.stack([torch.einsum("bol,bil->oi", dy, xp[:, :, k:k + L]) for k in range(K)], -1)
    dx = x.new_zeros(Bn, Ci, L)
    for k in range(K):
        # dx[t] += w[:, :, k]^T dy[t - k + pad]
        dx = dx + torch.einsum("oi,bol->bil", w[:, :, k], dyp[:, :, 2 * pad - k:2 * pad - k + L])
    return dx, dw, dy.sum((0, 2))


def leaky_relu(x: Tensor, slope: float) -> Tensor:
    return torch.where(x > 0, x, x * slope)


def stem_fwd(x: Tensor, p: Dict[str, Tensor], training: bool,
             stats_allreduce=None):
    """conv1 = Conv1d(2,8,3,p1) -> LeakyReLU(0.2) -> BatchNorm1d(8)  (model/transformer.py:570-574).
    Returns token-major (B, L, 8), saved, and (new_running_mean, new_running_var) in training.
    `stats_allreduce(t)` (optional) sums the [sum, sumsq, n] vector across data-parallel ranks."""
    c = conv1d_fwd(x, p["conv1.0.weight"], p["conv1.0.bias"])
    a = leaky_relu(c, 0.2)
    new_stats = None
    if training:
        st = torch.cat([a.sum((0, 2)), (a * a).sum((0, 2)),
                        a.new_tensor([float(a.shape[0] * a.shape[2])])])
        if stats_allreduce is not None:
            st = stats_allreduce(st)
        n = st[16]
        mu = st[:8] / n
        var = st[8:16] / n - mu * mu
        new_stats = ((1 - BN_MOMENTUM) * p["conv1.2.running_mean"] + BN_MOMENTUM * mu,
                     (1 - BN_MOMENTUM) * p["conv1.2.running_var"] + BN_MOMENTUM * var * n / (n - 1),
                     n)
    else:
        mu, var = p["conv1.2.running_mean"], p["conv1.2.running_var"]
    rstd = torch.rsqrt(var + BN_EPS)
    ah = (a - mu.view(1, -1, 1)) * rstd.view(1, -1, 1)
    y = ah * p["conv1.2.weight"].view(1, -1, 1) + p["conv1.2.bias"].view(1, -1, 1)
    return y.permute(0, 2, 1).contiguous(), dict(x=x, c=c, ah=ah, rstd=rstd, training=training), new_stats


def stem_bwd(g_tok: Tensor, saved, p: Dict[str, Tensor], n_total=None, sums_allreduce=None):
    """g_tok: (B, L, 8) gradient w.r.t. the token-major stem output."""
    dy = g_tok.permute(0, 2, 1)
    ah, rstd, c, x = saved["ah"], saved["rstd"], saved["c"], saved["x"]
    grads = {"conv1.2.weight": (dy * ah).sum((0, 2)), "conv1.2.bias": dy.sum((0, 2))}
    dah = dy * p["conv1.2.weight"].view(1, -1, 1)
    if saved["training"]:
        s = torch.cat([dah.sum((0, 2)), (dah * ah).sum((0, 2))])
        if sums_allreduce is not None:
            s = sums_allreduce(s)
        n = float(dy.shape[0] * dy.shape[2]) if n_total is None else float(n_total)
        m1, m2 = (s[:8] / n).view(1, -1, 1), (s[8:] / n).view(1, -1, 1)
        da = rstd.view(1, -1, 1) * (dah - m1 - ah * m2)
    else:
        da = dah * rstd.view(1, -1, 1)
    dc = da * torch.where(c > 0, torch.ones_like(c), torch.full_like(c, 0.2))
    dx, dw, db = conv1d_bwd(dc, x, p["conv1.0.weight"])
    grads["conv1.0.weight"], grads["conv1.0.bias"] = dw, db
    return dx, grads


def head_fwd(x_tok: Tensor, skip_tok: Tensor, p: Dict[str, Tensor]):
    """transconv(x_1^T + stem_out)   (model/transformer.py:664-667)."""
    s = (x_tok + skip_tok).permute(0, 2, 1)
    return conv1d_fwd(s, p["transconv.0.weight"], p["transconv.0.bias"]), dict(s=s)


def head_bwd(dout: Tensor, saved, p: Dict[str, Tensor]):
    ds, dw, db = conv1d_bwd(dout, saved["s"], p["transconv.0.weight"])
    return ds.permute(0, 2, 1).contiguous(), {"transconv.0.weight": dw, "transconv.0.bias": db}


def mse_loss_fwd_bwd(pred: Tensor, target: Tensor):
    """F.mse_loss, mean over every element (denoise_train.py:53) and its gradient."""
    d = pred - target
    return (d * d).mean(), d * (2.0 / d.numel())


def weighted_mse_loss_fwd_bwd(pred: Tensor, target: Tensor, w: Tensor):
    """the north star's R-wave-weighted reconstruction loss, an opt-in EXTENSION with no counterpart in the
    reference (the reference's loss is plain F.mse_loss, denoise_train.py:53):
        loss = mean_{b,i} w[i] (pred[b,i] - target[b,i])^2,   dL/dpred = 2 w[i] (pred - target) / numel,
    w a per-position weight over the flattened (lead, sample) axis of a window; w == 1 is F.mse_loss exactly."""
    B = pred.shape[0]
    d = (pred - target).reshape(B, -1)
    wv = w.reshape(1, -1).to(d.dtype)
    return (wv * d * d).mean(), (wv * d * (2.0 / d.numel())).reshape(pred.shape)


def RMSE(y: Tensor, y_pred: Tensor) -> Tensor:
    """local_utils/evaluate.py:27-29."""
    d = (y.flatten(1) - y_pred.flatten(1))
    return torch.sqrt((d * d).mean(-1))


def SNR(y: Tensor, y_pred: Tensor) -> Tensor:
    """local_utils/evaluate.py:49-51."""
    yf, d = y.flatten(1), (y.flatten(1) - y_pred.flatten(1))
    return 10 * torch.log10((yf * yf).mean(-1) / (d * d).mean(-1))


# ----------------------------------------------------------------------------------------
# whole network
# ----------------------------------------------------------------------------------------
def _sub(sd: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def detect_variant(sd: Dict[str, Tensor]) -> str:
    return "rw" if "rwattn1.relative_position_bias_table" in sd else "nra"


def ralenet_fwd(x: Tensor, sd: Dict[str, Tensor], training: bool = False,
                stats_allreduce=None):
    """ralenet.forward (model/transformer.py:621-667 / raletransformer.py:640-680).
    Returns out (B,2,L), tape (for ralenet_bwd), new BN running stats (training only)."""
    variant = detect_variant(sd)
    tape: List = []
    x0, st_saved, new_stats = stem_fwd(x, sd, training, stats_allreduce)
    t = x0
    skips: Dict[int, Tensor] = {}
    for li, (layer, s, rw) in enumerate(LAYERS):
        H = HEADS[s]
        table, W = None, 0
        if variant == "rw" and rw is not None:
            table, W = sd[f"rwattn{rw}.relative_position_bias_table"], RW_WINDOW[rw - 1]
        for i in range(2):
            p = _sub(sd, block_prefix(variant, layer, i))
            t, sa = attn_block_fwd(t, p, H, table, W)
            t, sf = ffn_block_fwd(t, p)
            tape.append(("blk", layer, i, s, rw, sa, sf))
        if li < 4:                                   # down path: pm1..pm4 (:633-642)
            t, sp = patch_merge_fwd(t, _sub(sd, f"pm{li + 1}."))
            skips[li + 1] = t
            tape.append(("pm", li + 1, sp))
        elif li == 4:                                # x_mid += x4  (:646)
            t = t + skips[4]
            tape.append(("mid",))
        else:                                        # up path: ps4..ps1 (:649-661)
            j = 9 - li                               # li=5->ps4 ... li=8->ps1
            skip = skips[j - 1] if j > 1 else None
            t, sp = patch_separate_fwd(t, _sub(sd, f"ps{j}."), skip)
            tape.append(("ps", j, sp))
    out, hs = head_fwd(t, x0, sd)
    return out, dict(variant=variant, stem=st_saved, tape=tape, head=hs), new_stats


def ralenet_bwd(dout: Tensor, ctx, sd: Dict[str, Tensor], n_total=None, sums_allreduce=None):
    """Manual backward of ralenet_fwd: returns dx (B,2,L) and {state_dict key: grad}."""
    variant = ctx["variant"]
    grads: Dict[str, Tensor] = {}

    def acc(prefix, g):
        for k, v in g.items():
            key = prefix + k
            grads[key] = grads[key] + v if key in grads else v

    ds, gh = head_bwd(dout, ctx["head"], sd)
    acc("", gh)
    g = ds                      # gradient w.r.t. x_1 (token-major)
    g_x0 = ds.clone()           # and w.r.t. the stem output through the final skip
    g_skip: Dict[int, Tensor] = {}
    for item in reversed(ctx["tape"]):
        kind = item[0]
        if kind == "ps":
            _, j, sp = item
            if j > 1:
                g_skip[j - 1] = g
            g, gp = patch_separate_bwd(g, sp, _sub(sd, f"ps{j}."))
            acc(f"ps{j}.", gp)
        elif kind == "mid":
            g_skip[4] = g
        elif kind == "pm":
            _, j, sp = item
            g = g + g_skip[j]
            g, gp = patch_merge_bwd(g, sp, _sub(sd, f"pm{j}."))
            acc(f"pm{j}.", gp)
        else:
            _, layer, i, s, rw, sa, sf = item
            pre = block_prefix(variant, layer, i)
            p = _sub(sd, pre)
            table, W = None, 0
            if variant == "rw" and rw is not None:
                table, W = sd[f"rwattn{rw}.relative_position_bias_table"], RW_WINDOW[rw - 1]
            g, gf = ffn_block_bwd(g, sf, p)
            g, ga = attn_block_bwd(g, sa, p, HEADS[s], table, W)
            if "table" in ga:
                acc(f"rwattn{rw}.relative_position_bias_", {"table": ga.pop("table")})
            acc(pre, gf)
            acc(pre, ga)
    g = g + g_x0
    dx, gs = stem_bwd(g, ctx["stem"], sd, n_total, sums_allreduce)
    acc("", gs)
    return dx, grads


def newrale_fwd(x: Tensor, sd: Dict[str, Tensor], training: bool = False):
    """newrale.forward (model/ralenet_12leads.py:698-709): LReLU(conv1) -> LReLU(conv2) -> rale ->
    LReLU(conv3) -> conv4; all Conv1d k=13 p=6; nn.LeakyReLU() default slope 0.01."""
    c1 = conv1d_fwd(x, sd["conv1.weight"], sd["conv1.bias"])
    a1 = leaky_relu(c1, 0.01)
    c2 = conv1d_fwd(a1, sd["conv2.weight"], sd["conv2.bias"])
    a2 = leaky_relu(c2, 0.01)
    core_sd = _sub(sd, "rale.")
    r, ctx, new_stats = ralenet_fwd(a2, core_sd, training)
    c3 = conv1d_fwd(r, sd["conv3.weight"], sd["conv3.bias"])
    a3 = leaky_relu(c3, 0.01)
    out = conv1d_fwd(a3, sd["conv4.weight"], sd["conv4.bias"])
    return out, dict(x=x, c1=c1, a1=a1, c2=c2, a2=a2, r=r, c3=c3, a3=a3, core=ctx, core_sd=core_sd), new_stats


def newrale_bwd(dout: Tensor, ctx, sd: Dict[str, Tensor]):
    """grads of the four trainable convs only (the core is frozen, ralenet_12leads.py:695-696),
    but the gradient still flows *through* the core."""
    def dl(c):
        return torch.where(c > 0, torch.ones_like(c), torch.full_like(c, 0.01))
    grads = {}
    da3, grads["conv4.weight"], grads["conv4.bias"] = conv1d_bwd(dout, ctx["a3"], sd["conv4.weight"])
    dr, grads["conv3.weight"], grads["conv3.bias"] = conv1d_bwd(da3 * dl(ctx["c3"]), ctx["r"], sd["conv3.weight"])
    da2, _ = ralenet_bwd(dr, ctx["core"], ctx["core_sd"])
    da1, grads["conv2.weight"], grads["conv2.bias"] = conv1d_bwd(da2 * dl(ctx["c2"]), ctx["a1"], sd["conv2.weight"])
    dx, grads["conv1.weight"], grads["conv1.bias"] = conv1d_bwd(da1 * dl(ctx["c1"]), ctx["x"], sd["conv1.weight"])
    return dx, grads


def adam_step(param: Tensor, grad: Tensor, m: Tensor, v: Tensor, step: int,
              lr: float = 1e-3, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam defaults as used at denoise_train.py:24 (no weight decay, no amsgrad).
    In-place on param, m, v; `step` is the 1-based step count."""
    m.mul_(b1).add_(grad, alpha=1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(m, denom, value=-lr / bc1)


def trainable_keys(sd: Dict[str, Tensor]) -> List[str]:
    return [k for k, v in sd.items() if v.dtype.is_floating_point
            and not k.endswith("running_mean") and not k.endswith("running_var")]


# ----------------------------------------------------------------------------------------
# record <-> window pipeline (local_utils/local_utils.py:47-65, 116-130, 261-266)
# ----------------------------------------------------------------------------------------
def records_to_windows(x: Tensor, window: int = 256, stride: int = 256, znorm: bool = True):
    """x: (R, C, T).  Per-lead z-normalisation like np_norm ((x - mean) / std, population std), then the
    reference's cut `signal[i:i+window]` for i in range(0, T - window + 1, stride).  Returns windows
    (R*nper, C, window) and (mean, std) of shape (R, C, 1)."""
    R, C, T = x.shape
    mean = x.mean(-1, keepdim=True) if znorm else torch.zeros(R, C, 1, dtype=x.dtype)
    std = x.std(-1, unbiased=False, keepdim=True) if znorm else torch.ones(R, C, 1, dtype=x.dtype)
    xn = (x - mean) / std
    nper = (T - window) // stride + 1
    win = torch.stack([xn[:, :, w * stride:w * stride + window] for w in range(nper)], 1)   # (R, nper, C, W)
    return win.reshape(R * nper, C, window), (mean, std)


def windows_to_records(win: Tensor, x: Tensor, stats, stride: int = 256) -> Tensor:
    """overlap-add average of the windows, de-normalised; samples not covered by a full window come from x."""
    R, C, T = x.shape
    window = win.shape[-1]
    nper = (T - window) // stride + 1
    w4 = win.reshape(R, nper, C, window)
    acc = torch.zeros_like(x)
    cnt = torch.zeros(T, dtype=x.dtype)
    for w in range(nper):
        acc[:, :, w * stride:w * stride + window] += w4[:, w]
        cnt[w * stride:w * stride + window] += 1
    mean, std = stats
    y = torch.where(cnt > 0, acc / cnt.clamp(min=1) * std + mean, x)
    return y


# ------------------------------------------------------------------------------------------------
# Section 8f extras: efficient channel attention and SNR-targeted noise mixing
def eca_fwd(x: Tensor, w: Tensor, res: Optional[Tensor] = None):
    """eca_layer_1d.forward on (B, L, C) tokens (model/transformer.py:109-113): avg-pool over the tokens, Conv1d(1,1,k)
    over the channel axis (zero padded, no bias), sigmoid, channel-wise scaling; `res` = the block residual that
    TransformerBlock.forward adds afterwards (:410).  Returns (y, saved)."""
    B, L, C = x.shape
    K = w.numel()
    P = (K - 1) // 2
    m = x.mean(dim=1)                                               # (B, C)
    mp = torch.zeros(B, C + 2 * P, dtype=x.dtype)
    mp[:, P:P + C] = m
    z = sum(w.reshape(-1)[k] * mp[:, k:k + C] for k in range(K))    # z[c] = sum_k w[k] m[c + k - P]
    s = torch.sigmoid(z)
    y = x * s[:, None, :]
    if res is not None:
        y = y + res
    return y, (x, m, s)


def eca_bwd(g: Tensor, saved, w: Tensor):
    """gradients of eca_fwd w.r.t. x and w (the gradient w.r.t. res is g)."""
    x, m, s = saved
    B, L, C = x.shape
    K = w.numel()
    P = (K - 1) // 2
    ds = (g * x).sum(dim=1)                                         # (B, C)
    dz = ds * s * (1 - s)
    dzp = torch.zeros(B, C + 2 * P, dtype=x.dtype)
    dzp[:, P:P + C] = dz
    mp = torch.zeros(B, C + 2 * P, dtype=x.dtype)
    mp[:, P:P + C] = m
    # dm[c'] = sum_k w[k] dz[c' - k + P];  dw[k] = sum_{b,c} dz[c] m[c + k - P]
    dm = sum(w.reshape(-1)[k] * dzp[:, 2 * P - k:2 * P - k + C] for k in range(K))
    dw = torch.stack([(dz * mp[:, k:k + C]).sum() for k in range(K)]).reshape(w.shape)
    dx = g * s[:, None, :] + dm[:, None, :] / L
    return dx, dw


def snr_mix(data: Tensor, noise: Tensor, snr_db: Tensor) -> Tensor:
    """single_snr_noise_add (local_utils/local_utils.py:176-192) per window of a batch: energies are means over ALL
    elements of a window (leads and samples), scale = sqrt(signal_energy / 10**(snr/10) / noise_energy)."""
    B = data.shape[0]
    d, n = data.reshape(B, -1), noise.reshape(B, -1)
    ps = (d.abs() ** 2).mean(dim=1)
    pn = (n.abs() ** 2).mean(dim=1)
    target = ps / (10 ** (snr_db.to(data.dtype) / 10))
    scale = torch.sqrt(target / pn)
    return data + noise * scale.reshape(B, *([1] * (data.dim() - 1)))


def ffn_eca_block_fwd(x1: Tensor, p: Dict[str, Tensor]):
    """feed-forward half of a TransformerBlock built with use_eca=True: x1 + eca(Mlp(LN2(x1)))
    (model/transformer.py:158, :392-395, :410); p carries "mlp.eca.conv.weight"."""
    y_res, saved = ffn_block_fwd(x1, p)
    y, esaved = eca_fwd(y_res - x1, p["mlp.eca.conv.weight"], res=x1)
    return y, (saved, esaved)


def ffn_eca_block_bwd(g: Tensor, saved, p: Dict[str, Tensor]):
    fsaved, esaved = saved
    d, dw = eca_bwd(g, esaved, p["mlp.eca.conv.weight"])
    dx_with_res, grads = ffn_block_bwd(d, fsaved, p)          # = d + dz
    grads["mlp.eca.conv.weight"] = dw
    return g + (dx_with_res - d), grads
