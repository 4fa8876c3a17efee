"""Debug helper: phase timeline of the attention kernels (library built with -DRL_TRACE, e.g.
tools/build_variant.sh trace -DRL_TRACE; RALENET_B200_LIB=build/variants/trace.so python tools/trace_attn.py [C ...]).
Runs forward and backward at B = 256 and prints, per phase, the median / max SM-clock delta over the CTAs."""
import ctypes, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecg_denoise_b200 import _lib

B = 256
dev = torch.device("cuda:0")
lib = _lib.load()
lib.ralenet_debug_trace_read_attn.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
FWD = ["prefetch+pdl", "PE+LN", "qkv GEMM", "store qkv", "core", "store o", "proj GEMM", "epilogue"]
BWD = ["prefetch+pdl", "stage", "do GEMM", "amax + D + wgrad(wp)", "repack fp16 pairs", "core (single pass)", "dqkv scratch",
       "du GEMM", "LN bwd", "wgrad(q,kv)+table"]


def read(n):
    buf = (ctypes.c_longlong * (B * 16))()
    assert lib.ralenet_debug_trace_read_attn(buf, B * 16) == 0
    return np.array(buf[:], dtype=np.int64).reshape(B, 16)[:, :n]


def report(tag, names, t, us):
    d = np.diff(t, axis=1)
    print(f"--- {tag}: {us:.1f} us per launch (back to back)")
    for i, n in enumerate(names):
        print(f"  {n:20s} median {np.median(d[:, i]) / 1.965e3:7.2f} us   max {d[:, i].max() / 1.965e3:7.2f} us")
    tot = t[:, -1] - t[:, 0]
    print(f"  {'CTA total':20s} median {np.median(tot) / 1.965e3:7.2f} us   max {tot.max() / 1.965e3:7.2f} us")


for C in [int(c) for c in sys.argv[1:]] or [8, 16, 32, 64, 128]:
    L, H = 2048 // C, C // 4
    W = {8: 32, 16: 16, 32: 8, 64: 4, 128: 0}[C]
    g = torch.Generator(device="cpu").manual_seed(0)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.2).to(dev).contiguous()
    x, pe = r(B, L, C), r(L, C)
    ln_w, ln_b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    wq, bq, wkv, bkv, wp, bp = r(C, C), r(C), r(2 * C, C), r(2 * C), r(C, C), r(C)
    table = r(max(2 * W - 1, 1), H)
    y, q, k, v, o = (torch.empty(B, L, C, device=dev) for _ in range(5))
    lse = torch.empty(B, H, L, device=dev)
    A = _lib.STRUCTS["rl_attn_fwd_args"]()
    A.B, A.L, A.C, A.H, A.W, A.c0, A.flags = B, L, C, H, W, (L - W) // 2 if W else 0, 3
    for kk, t in dict(x=x, pe=pe, ln_w=ln_w, ln_b=ln_b, wq=wq, bq=bq, wkv=wkv, bkv=bkv, wp=wp, bp=bp, y=y, q=q, k=k, v=v,
                      o=o, lse=lse).items():
        setattr(A, kk, t.data_ptr())
    if W:
        A.table = table.data_ptr()

    def run(fn, arg, n):
        for _ in range(3):
            assert fn(ctypes.byref(arg), st) == 0, lib.ralenet_last_error()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn(ctypes.byref(arg), st)
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / n

    us = run(lib.ralenet_attn_fwd, A, 20)
    report(f"attn_fwd<{C}>", FWD, read(9), us)

    Bw = _lib.STRUCTS["rl_attn_bwd_args"]()
    Bw.B, Bw.L, Bw.C, Bw.H, Bw.W, Bw.c0, Bw.flags = B, L, C, H, W, A.c0, 3
    gy, dx, dqkv, u = r(B, L, C), torch.empty(B, L, C, device=dev), torch.empty(B, L, 3 * C, device=dev), torch.empty(B, L, C, device=dev)
    grads = {n: torch.zeros_like(t) for n, t in dict(d_ln_w=ln_w, d_ln_b=ln_b, d_wq=wq, d_bq=bq, d_wkv=wkv, d_bkv=bkv, d_wp=wp,
                                                      d_bp=bp).items()}
    d_table = torch.zeros_like(table)
    for kk, t in dict(g=gy, x=x, pe=pe, ln_w=ln_w, ln_b=ln_b, wq=wq, wkv=wkv, wp=wp, q=q, k=k, v=v, o=o, lse=lse, dx=dx,
                      dqkv=dqkv, u=u, **grads).items():
        setattr(Bw, kk, t.data_ptr())
    if W:
        Bw.table, Bw.d_table = table.data_ptr(), d_table.data_ptr()
    lib.ralenet_attn_bwd_main = getattr(lib, "ralenet_attn_bwd")
    us = run(lib.ralenet_attn_bwd, Bw, 20)
    report(f"attn_bwd<{C}> (+ its wgrad launch when C > 16)", BWD, read(11), us)
