"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol declared in
include/ralenet_b200.h (no compute calls), the ctypes structs generated from the header match the C
compiler's layout, the nn.Module mirrors reproduce the reference's state_dict layout, missing CUDA is loud,
and the data-parallel algorithm (sharded batch + all-reduced BatchNorm statistics + summed gradients) equals
the single-process result -- run with 2 gloo ranks on the oracle."""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from oracle import ralenet_oracle as O
from oracle import synth_weights as SW

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge
    from ecg_denoise_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    assert len(built_lib.FUNCTIONS) >= 26
    for name in built_lib.FUNCTIONS:
        assert hasattr(lib, name), f"{name} declared in include/ralenet_b200.h but not exported"
    assert built_lib.load().ralenet_abi_version() == built_lib.CONSTS["RL_ABI_VERSION"]


def test_ctypes_structs_match_c_layout(built_lib):
    """compile a tiny C program printing sizeof/offsetof of every struct and compare with ctypes."""
    lines = ['#include "ralenet_b200.h"', "#include <stdio.h>", "#include <stddef.h>", "int main(){"]
    for sname, st in built_lib.STRUCTS.items():
        lines.append(f'printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "l.c"), os.path.join(d, "l")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split("\n")
    got = {l.split()[0]: int(l.split()[1]) for l in out if l.strip()}
    for sname, st in built_lib.STRUCTS.items():
        assert got[sname] == ctypes.sizeof(st), sname
        for fname, _ in st._fields_:
            assert got[f"{sname}.{fname}"] == getattr(st, fname).offset, (sname, fname)


def test_kernel_selection_switches_round_trip(built_lib):
    """ralenet_set_attn_umma / ralenet_set_wgrad_umma select between two implementations of the same function; the
    setters return the previous value, clamp bad input to the default and need no GPU."""
    prev = built_lib.set_attn_umma(0)
    try:
        assert prev in (0, 1, 2)
        assert built_lib.set_attn_umma(2) == 0
        assert built_lib.set_attn_umma(7) == 2            # out of range -> build default (1)
        assert built_lib.set_attn_umma(1) == 1
    finally:
        built_lib.set_attn_umma(prev)
    prev = built_lib.set_wgrad_umma(False)
    try:
        assert built_lib.set_wgrad_umma(True) is False
        assert built_lib.set_wgrad_umma(True) is True
    finally:
        built_lib.set_wgrad_umma(prev)


def test_windows_per_record_matches_the_reference_cut(built_lib):
    """record -> window cut of the reference (local_utils/local_utils.py:53: range(0, len - window + 1, stride)): the
    host-side count the inference pipeline sizes its buffers with, checked against the Python range for the
    BASELINE.json record (650000 samples -> 2539 windows of 256, 5077 at stride 128) and edge cases."""
    from ecg_denoise_b200 import inference
    assert inference.windows_per_record(650000, 256, 256) == 2539
    assert inference.windows_per_record(650000, 256, 128) == 5077
    rs = np.random.RandomState(0)
    for _ in range(200):
        T, W, st = int(rs.randint(1, 5000)), int(rs.choice([16, 256, 512])), int(rs.randint(1, 600))
        assert inference.windows_per_record(T, W, st) == len(range(0, T - W + 1, st)), (T, W, st)
    assert inference.windows_per_record(255, 256, 256) == 0      # shorter than one window
    assert inference.windows_per_record(256, 256, 256) == 1


def test_workspace_size_is_monotonic(built_lib):
    lib = built_lib.load()
    a = lib.ralenet_net_workspace_bytes(32, 256, 1)
    b = lib.ralenet_net_workspace_bytes(64, 256, 1)
    c = lib.ralenet_net_workspace_bytes(64, 256, 0)
    assert 0 < a < b and 0 < c < b
    assert lib.ralenet_net_workspace_bytes(0, 256, 1) == 0


@pytest.mark.parametrize("variant,le", [("rw", 1), ("rw", 0), ("nra", 1), ("rw12", 1)])
def test_module_mirror_has_reference_state_dict_layout(variant, le):
    from ecg_denoise_b200.model import ralenet_12leads, raletransformer, transformer
    if variant == "nra":
        m = raletransformer.ralenet()
    elif variant == "rw12":
        m = ralenet_12leads.ralenet(high_level_enhence=True)
    else:
        m = transformer.ralenet(high_level_enhence=bool(le))
    ref = SW.make_state_dict(variant, le)        # layout proven against the reference by oracle/make_golden.py
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape) and sd[k].dtype == ref[k].dtype, k
    m.load_state_dict(ref, strict=True)
    n = sum(p.numel() for p in m.parameters())
    assert n == {("rw", 1): 1087282, ("rw", 0): 1087228, ("nra", 1): 1086800, ("rw12", 1): 1087282}[(variant, le)]


def test_newrale_layout_and_frozen_core():
    from ecg_denoise_b200.model import ralenet_12leads as TW
    m = TW.newrale(TW.ralenet(high_level_enhence=True))
    ref = SW.make_newrale_state_dict()
    assert list(m.state_dict().keys()) == list(ref.keys())
    m.load_state_dict(ref, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 1089492
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 2210


def test_default_init_is_zero_rw_tables_and_reference_ctor_kwargs():
    from ecg_denoise_b200.model import transformer
    torch.manual_seed(2023)
    m = transformer.ralenet(qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0., mlp_ratio=4.,
                            low_level_enhence=False, high_level_enhence=False)
    assert float(m.rwattn1.relative_position_bias_table.abs().sum()) == 0.0
    assert not hasattr(m.dtransformer1.blocks[0].mlp, "leconv")
    dense = m.rwattn2()
    assert tuple(dense.shape) == (1, 4, 128, 128)


def test_cpu_forward_is_loud_not_a_fallback():
    from ecg_denoise_b200 import _lib
    from ecg_denoise_b200.model import transformer
    m = transformer.ralenet()
    with pytest.raises(_lib.RalenetError):
        m(torch.zeros(2, 2, 256))
    with pytest.raises(_lib.RalenetError):
        m.dtransformer1(torch.zeros(2, 256, 8))


def test_install_registers_reference_import_names():
    import ecg_denoise_b200
    saved = {k: sys.modules.get(k) for k in ("model", "model.transformer", "model.raletransformer",
                                             "model.ralenet_12leads")}
    try:
        ecg_denoise_b200.install()
        from model.transformer import ralenet as r1            # noqa: the reference's import line (main.py:75)
        from model.raletransformer import ralenet as r2        # main.py:72
        from model.ralenet_12leads import newrale, ralenet as r3   # Transfer_learning.py:72
        from ecg_denoise_b200.model import transformer
        assert r1 is transformer.ralenet and r2 is not r1 and r3 is not r1 and newrale is not None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_synthetic_generator_is_deterministic_and_snr_targeted():
    from ecg_denoise_b200 import synth
    n1, c1 = synth.make_batch(6, 2, 256, seed=3, kind="emb", snr_db=-4.0)
    n2, c2 = synth.make_batch(6, 2, 256, seed=3, kind="emb", snr_db=-4.0)
    assert np.array_equal(n1, n2) and np.array_equal(c1, c2)
    assert np.allclose(c1.mean(-1), 0, atol=1e-5) and np.allclose(c1.std(-1), 1, atol=1e-3)
    snr = 10 * np.log10((c1.astype(np.float64) ** 2).mean(-1) / ((n1 - c1).astype(np.float64) ** 2).mean(-1))
    assert np.allclose(snr, -4.0, atol=1e-3)
    # R peak of the middle beat within +-8 samples of the window centre on lead 0
    assert np.all(np.abs(np.argmax(np.abs(c1[:, 0, 96:160]), -1) + 96 - 128) <= 10)


# ---------------------------------------------------------------------------------------------------------
def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ecg_denoise_b200 import synth
        torch.set_num_threads(2)
        B = 6
        noisy, clean = synth.make_batch(B, 2, 256, seed=42)
        sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in SW.make_state_dict("rw", 1, 8).items()}
        lo, hi = rank * B // world, (rank + 1) * B // world
        x, t = torch.from_numpy(noisy[lo:hi]).double(), torch.from_numpy(clean[lo:hi]).double()

        def allreduce(v):
            v = v.clone()
            dist.all_reduce(v)
            return v

        out, ctx, new_stats = O.ralenet_fwd(x, sd, training=True, stats_allreduce=allreduce)
        d = out - t
        dout = d * (2.0 / (B * 2 * 256))                      # 1 / GLOBAL numel
        _, grads = O.ralenet_bwd(dout, ctx, sd, n_total=B * 256, sums_allreduce=allreduce)
        flat = torch.cat([grads[k].reshape(-1) for k in sorted(grads)])
        dist.all_reduce(flat)                                  # the one gradient all-reduce (sum)
        loss = (d * d).sum() / (B * 2 * 256)
        dist.all_reduce(loss)
        if rank == 0:
            q.put((flat.numpy(), loss.item(), new_stats[0].numpy(), new_stats[1].numpy()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_equals_single_process_gloo_world2():
    """batch sharded over 2 ranks + all-reduced BN stats + summed gradients == single-process global batch."""
    import torch.multiprocessing as mp
    from ecg_denoise_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, loss, rm, rv = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B = 6
    noisy, clean = synth.make_batch(B, 2, 256, seed=42)
    sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in SW.make_state_dict("rw", 1, 8).items()}
    out, c, new_stats = O.ralenet_fwd(torch.from_numpy(noisy).double(), sd, training=True)
    loss_ref, dout = O.mse_loss_fwd_bwd(out, torch.from_numpy(clean).double())
    _, grads = O.ralenet_bwd(dout, c, sd)
    flat_ref = torch.cat([grads[k].reshape(-1) for k in sorted(grads)]).numpy()
    assert abs(loss - loss_ref.item()) < 1e-12
    assert np.allclose(flat, flat_ref, rtol=1e-9, atol=1e-12)
    assert np.allclose(rm, new_stats[0].numpy(), rtol=1e-12) and np.allclose(rv, new_stats[1].numpy(), rtol=1e-12)


def test_kernel_gelu_formula_matches_exact_erf_gelu():
    """common.cuh::gelu_parts replaces libdevice erff by the Abramowitz-Stegun 7.1.26 tail form.  Restated here in
    numpy float32, operation by operation, and held against the fp64 nn.GELU() (model/transformer.py:138, exact erf)
    and its derivative: absolute error below 1e-6 on [-12, 12] -- the level of the erff-based fp32 evaluation."""
    from scipy.special import erf
    f = np.float32
    x = np.linspace(-12, 12, 400001).astype(f)
    z = np.abs(x) * f(0.70710678118654752)
    t = (f(1) / (f(0.3275911) * z + f(1))).astype(f)
    e = np.exp2((x * x * f(-0.72134752044448170)).astype(f)).astype(f)
    q = (f(0.5 * 1.061405429) * t + f(0.5 * -1.453152027)).astype(f)
    for coef in (0.5 * 1.421413741, 0.5 * -0.284496736, 0.5 * 0.254829592):
        q = (q * t + f(coef)).astype(f)
    q = (q * t * e).astype(f)
    cdf = np.where(x < 0, q, f(1) - q).astype(f)
    g = (x * cdf).astype(f)
    dg = (x * e * f(0.3989422804014327) + cdf).astype(f)
    xd = x.astype(np.float64)
    cdf64 = 0.5 * (1 + erf(xd / np.sqrt(2)))
    g64 = xd * cdf64
    dg64 = cdf64 + xd * np.exp(-0.5 * xd * xd) / np.sqrt(2 * np.pi)
    assert np.abs(g - g64).max() < 1e-6
    assert np.abs(dg - dg64).max() < 1e-6


# ------------------------------------------------------------------------------------------------------
# round 2: host-side behaviour of the module mirror that needs no GPU
def test_ralenet_deepcopy_and_pickle():
    """copy.deepcopy (EMA / best-model snapshots) and torch.save(model) work like on the reference's modules; the
    copy gets its own plan (ctypes pointer tables are never copied)."""
    import copy
    import io
    from ecg_denoise_b200.model import ralenet_12leads, transformer
    m = transformer.ralenet(high_level_enhence=True)
    m2 = copy.deepcopy(m)
    assert m2._plan is not m._plan and m2._plan.net is m2
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    assert list(m3.state_dict().keys()) == list(m.state_dict().keys()) and m3._plan.net is m3
    w = ralenet_12leads.newrale(ralenet_12leads.ralenet(high_level_enhence=True))
    w2 = copy.deepcopy(w)
    assert w2.rale._plan.net is w2.rale


def test_flat_parameter_plan_tracks_every_parameter_and_partial_grads():
    """NetPlan (host logic only, on CPU tensors): parameters become views of one flat buffer; re-assigning ANY
    parameter's storage is detected; attach_grads() zeroes only the slices whose .grad is missing and folds a
    foreign .grad in, keeping what the other slices have accumulated."""
    from ecg_denoise_b200.model import transformer
    m = transformer.ralenet(high_level_enhence=True)
    plan = m._plan
    dev = torch.device("cpu")
    plan.ensure(dev)
    assert plan._aliased()
    ps = list(m.parameters())
    assert all(p.data_ptr() == plan.flat.data_ptr() + 4 * o for p, o in zip(ps, plan.offsets))
    victim = ps[17]                                     # not one of the three the old check sampled
    victim.data = victim.data.clone()
    assert not plan._aliased()
    plan.ensure(dev)
    assert plan._aliased()
    assert plan.attach_grads() is True                  # all None -> one memset, every grad aliased
    assert plan.attach_grads() is False
    plan.flat_grad.fill_(1.0)
    a, b, c = ps[3], ps[40], ps[77]
    a.grad = None
    b.grad = torch.full_like(b, 5.0)                    # foreign tensor
    assert plan.attach_grads() is True
    assert float(a.grad.abs().max()) == 0.0
    assert float((b.grad - 5.0).abs().max()) == 0.0
    assert float((c.grad - 1.0).abs().max()) == 0.0     # untouched slice kept its accumulated value
    assert a.grad.data_ptr() == plan.flat_grad.data_ptr() + 4 * plan.offsets[3]


def test_step_graph_capture_priority_rule(monkeypatch):
    """engine._capture_stream: per-GPU batches up to PRIO_MAX_BATCH capture the step on the highest-priority stream
    (the data-gradient chain ranks above the forked weight-gradient branch), larger ones on torch's default capture
    stream; RALENET_MAIN_PRIO overrides both ways (profiles/r2_v50_exp_prio*.txt is the measurement behind the rule)."""
    from ecg_denoise_b200 import engine

    made = []

    class FakeStream:
        def __init__(self, priority=0):
            made.append(priority)

        @staticmethod
        def priority_range():
            return (0, -3)

    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.delenv("RALENET_MAIN_PRIO", raising=False)
    assert isinstance(engine._capture_stream(256), FakeStream) and made == [-3]
    assert isinstance(engine._capture_stream(engine.PRIO_MAX_BATCH), FakeStream) and made == [-3, -3]
    assert engine._capture_stream(engine.PRIO_MAX_BATCH + 1) is None and engine._capture_stream(4096) is None
    monkeypatch.setenv("RALENET_MAIN_PRIO", "0")
    assert isinstance(engine._capture_stream(256), FakeStream) and made[-1] == 0
    monkeypatch.setenv("RALENET_MAIN_PRIO", "-2")
    assert isinstance(engine._capture_stream(4096), FakeStream) and made[-1] == -2
    monkeypatch.setenv("RALENET_MAIN_PRIO", "")
    assert engine._capture_stream(4096) is None
