"""Debug helper: phase timeline of the tcgen05 feed-forward kernel (needs ecg_denoise_b200/csrc/ffn_umma.cu compiled
with -DRL_TRACE and relinked into libralenet_b200.so).  Runs the forward op at the s4 shape (B = 256, L = 16, C = 128) and prints, per phase, the median /
max SM-clock delta over the CTAs."""
import ctypes, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecg_denoise_b200 import _lib

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L, B = 2048 // C, 256
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(B, L, C, generator=g).to(dev)
p = dict(ln_w=torch.ones(C), ln_b=torch.zeros(C), w1=torch.randn(4 * C, C, generator=g) * 0.1, b1=torch.zeros(4 * C),
         w2=torch.randn(C, 4 * C, generator=g) * 0.1, b2=torch.zeros(C), lew=torch.randn(3, generator=g))
p = {k: v.to(dev).contiguous() for k, v in p.items()}
lib = _lib.load()
A = _lib.STRUCTS["rl_ffn_fwd_args"]()
y = torch.empty_like(x); h = torch.empty(B, L, 4 * C, device=dev)
A.B, A.L, A.C, A.le_mode, A.flags = B, L, C, 1, 3
for k, t in dict(x=x, ln_w=p["ln_w"], ln_b=p["ln_b"], w1=p["w1"], b1=p["b1"], w2=p["w2"], b2=p["b2"], lew=p["lew"], y=y, h=h).items():
    setattr(A, k, t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    rc = lib.ralenet_ffn_fwd(ctypes.byref(A), st)
    assert rc == 0, lib.ralenet_last_error()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    lib.ralenet_ffn_fwd(ctypes.byref(A), st)
e1.record(); torch.cuda.synchronize()
print("avg kernel+launch us (back to back):", 1e3 * e0.elapsed_time(e1) / 20)
ncta = (B * L // 128) * (4 * C // 128)
buf = (ctypes.c_longlong * (ncta * 16))()
lib.ralenet_debug_trace_read_umma.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
print("read rc", lib.ralenet_debug_trace_read_umma(buf, ncta * 16))
t = np.array(buf[:], dtype=np.int64).reshape(ncta, 16)[:, :12]
d = np.diff(t, axis=1)
names = ["prefetch->pdl", "pdl->alloc/init", "LN tile", "fc1 stage+issue", "wait fc1", "epilogue1", "fc2 stage+issue+wait",
         "epilogue2 ld", "cluster.sync", "reduce+store", "cluster.sync2+dealloc"]
for i, n in enumerate(names):
    print(f"{n:24s} median {np.median(d[:, i]):8.0f}  max {d[:, i].max():8.0f} cycles")
print("total median", np.median(t[:, 11] - t[:, 0]), "max", (t[:, 11] - t[:, 0]).max(), "cycles; /1.965 GHz =",
      np.median(t[:, 11] - t[:, 0]) / 1965, "us")
