"""CPU: the oracle restatement vs golden vectors produced by the imported reference
(oracle/make_golden.py).  fp64 so the pin is sharp (1e-9), independent of summation order."""
import numpy as np
import pytest
import torch

from oracle import ralenet_oracle as O
from oracle import synth_weights as SW
from tests.common import CASES, check_packed, rel_rms_err, unpack_golden


def _sd64(variant, le, seed):
    return {k: (v.double() if v.dtype.is_floating_point else v) for k, v in SW.make_state_dict(variant, le, seed).items()}


@pytest.mark.parametrize("tag", list(CASES))
def test_forward_eval_matches_reference(golden, tag):
    variant, le, seed, xk, tk = CASES[tag]
    sd = _sd64(variant, le, seed)
    x = torch.from_numpy(golden[xk]).double()
    out, _, _ = O.ralenet_fwd(x, sd, training=False)
    assert rel_rms_err(out.numpy(), golden[f"{tag}/eval_out64"]) < 1e-9


@pytest.mark.parametrize("tag", list(CASES))
def test_train_forward_backward_matches_reference(golden, tag):
    variant, le, seed, xk, tk = CASES[tag]
    sd = _sd64(variant, le, seed)
    x = torch.from_numpy(golden[xk]).double()
    tgt = torch.from_numpy(golden[tk]).double()
    out, ctx, new_stats = O.ralenet_fwd(x, sd, training=True)
    assert rel_rms_err(out.numpy(), golden[f"{tag}/train_out64"]) < 1e-9
    loss, dout = O.mse_loss_fwd_bwd(out, tgt)
    assert abs(loss.item() - float(golden[f"{tag}/loss64"])) < 1e-12 * max(1.0, abs(loss.item()))
    dx, grads = O.ralenet_bwd(dout, ctx, sd)
    assert rel_rms_err(dx.numpy(), golden[f"{tag}/dx64"]) < 1e-8
    names = [str(n) for n in golden[f"{tag}/grad_names"]]
    assert sorted(names) == sorted(grads.keys())
    gold = unpack_golden(names, golden[f"{tag}/grads"], {n: tuple(sd[n].shape) for n in names})
    for n in names:
        assert tuple(grads[n].shape) == tuple(sd[n].shape), n
        check_packed(grads[n], gold[n], 1e-8, n)
    rm, rv, n = new_stats
    assert np.allclose(rm.numpy(), golden[f"{tag}/bn_running_mean"], rtol=1e-10, atol=1e-12)
    assert np.allclose(rv.numpy(), golden[f"{tag}/bn_running_var"], rtol=1e-10, atol=1e-12)
    assert int(golden[f"{tag}/bn_nbt"]) == 4


@pytest.mark.parametrize("tag", ["rw_le", "nra"])
def test_adam_trajectory_matches_reference(golden, tag):
    variant, le, seed, xk, tk = CASES[tag]
    sd = _sd64(variant, le, seed)
    x = torch.from_numpy(golden[xk]).double()
    tgt = torch.from_numpy(golden[tk]).double()
    names = [str(n) for n in golden[f"{tag}/grad_names"]]
    m = {n: torch.zeros_like(sd[n]) for n in names}
    v = {n: torch.zeros_like(sd[n]) for n in names}
    losses = []
    for step in range(1, 4):
        out, ctx, new_stats = O.ralenet_fwd(x, sd, training=True)
        loss, dout = O.mse_loss_fwd_bwd(out, tgt)
        losses.append(loss.item())
        _, grads = O.ralenet_bwd(dout, ctx, sd)
        for n in names:
            O.adam_step(sd[n], grads[n], m[n], v[n], step)
        sd["conv1.2.running_mean"], sd["conv1.2.running_var"] = new_stats[0], new_stats[1]
    assert np.allclose(losses, golden[f"{tag}/adam_losses64"], rtol=1e-9)
    gold = unpack_golden(names, golden[f"{tag}/adam_params"], {n: tuple(sd[n].shape) for n in names})
    for n in names:
        check_packed(sd[n], gold[n], 1e-7, n)


def test_newrale_matches_reference(golden):
    sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in SW.make_newrale_state_dict(2023).items()}
    x = torch.from_numpy(golden["x12"]).double()
    tgt = torch.from_numpy(golden["target12"]).double()
    out, _, _ = O.newrale_fwd(x, sd, training=False)
    assert rel_rms_err(out.numpy(), golden["newrale/eval_out64"]) < 1e-9
    out, ctx, _ = O.newrale_fwd(x, sd, training=True)
    assert rel_rms_err(out.numpy(), golden["newrale/train_out64"]) < 1e-9
    loss, dout = O.mse_loss_fwd_bwd(out, tgt)
    dx, grads = O.newrale_bwd(dout, ctx, sd)
    assert rel_rms_err(dx.numpy(), golden["newrale/dx64"]) < 1e-8
    names = [str(n) for n in golden["newrale/grad_names"]]
    assert sorted(names) == sorted(grads.keys())
    gold = unpack_golden(names, golden["newrale/grads"], {n: tuple(sd[n].shape) for n in names})
    for n in names:
        check_packed(grads[n], gold[n], 1e-8, n)


def test_metrics_match_reference(golden):
    sd = SW.make_state_dict("rw", 1, 2023)
    x = torch.from_numpy(golden["x"])
    tgt = torch.from_numpy(golden["target"])
    out, _, _ = O.ralenet_fwd(x, sd, training=False)
    assert rel_rms_err(out.numpy(), golden["rw_le/eval_out32"]) < 1e-4
    assert np.allclose(O.SNR(tgt, out).numpy(), golden["rw_le/eval_snr32"], atol=1e-3)
    assert np.allclose(O.RMSE(tgt, out).numpy(), golden["rw_le/eval_rmse32"], rtol=1e-4)


def test_depthwise_le_manual_backward_matches_autograd():
    """use_partial=False branch (model/transformer.py:145-146) is unreachable from ralenet, so it has
    no golden; pin the manual backward against autograd of the forward restatement instead."""
    torch.manual_seed(0)
    C = 16
    p = {k: v.double() for k, v in SW.make_state_dict("rw", 2, 5).items()
         if k.startswith("dtransformer2.blocks.0.")}
    p = {k[len("dtransformer2.blocks.0."):]: v.requires_grad_(True) for k, v in p.items()}
    x = torch.randn(3, 128, C, dtype=torch.float64, requires_grad=True)
    y, saved = O.ffn_block_fwd(x, p)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    dx, grads = O.ffn_block_bwd(g, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in saved.items()},
                                {k: v.detach() for k, v in p.items()})
    assert torch.allclose(dx, x.grad, rtol=1e-9, atol=1e-11)
    for k, gv in grads.items():
        assert torch.allclose(gv, p[k].grad, rtol=1e-9, atol=1e-11), k
