import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dit4sr_b200 import _lib
from dit4sr_b200._lib import AttnParams
def bench(libpath, variant, S=2381, B=8, reps=40):
    lib = C.CDLL(libpath)
    lib.dit4sr_attn_fwd.restype = C.c_int
    lib.dit4sr_attn_fwd.argtypes = [C.POINTER(AttnParams), C.c_void_p]
    H, D = 24, 1536
    torch.manual_seed(0)
    qkv = torch.randn(B, S, 3 * D, device="cuda"); qkv[..., :D] *= 1.4426950408889634 / 8.0; qkv = qkv.bfloat16()
    o = torch.empty(B, S, D, device="cuda", dtype=torch.bfloat16)
    p = AttnParams()
    p.q, p.k, p.v = qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D
    p.qkv_row_stride, p.qkv_batch_stride = 3 * D, S * 3 * D
    p.o, p.o_row_stride, p.o_batch_stride = o.data_ptr(), D, S * D
    p.batches, p.heads, p.seq_q, p.seq_kv, p.scale = B, H, S, S, 0.125
    p.mode, p.score_bound, p.variant = 1, 30.0, variant
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        assert lib.dit4sr_attn_fwd(C.byref(p), st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.dit4sr_attn_fwd(C.byref(p), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
here = os.path.dirname(os.path.abspath(__file__))
libs = [("b (sp_out)", here + "/b/libattn.so"), ("c (current)", ROOT + "/dit4sr_b200/libdit4sr_b200.so"), ("c1 (no decode)", here + "/c1/libattn.so")]
for rnd in range(3):
    for name, path in libs:
        for v in (3,):
            print(f"{name} variant {v}: {bench(path, v):.4f} ms", flush=True)
