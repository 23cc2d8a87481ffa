"""A/B timing of lina_gla_chunk_fwd from two builds of the library in the same process / on the same GPU.
usage: ab_gla_chunk.py libA.so libB.so"""
import ctypes as C, sys, torch, torch.nn.functional as F
B, H, T, K, V = 32, 4, 2048, 256, 512
torch.manual_seed(0)
q, k = (torch.randn(B, H, T, K, device="cuda", dtype=torch.bfloat16) for _ in range(2))
v = torch.randn(B, H, T, V, device="cuda", dtype=torch.bfloat16)
gk = (F.logsigmoid(torch.randn(B, H, T, K, device="cuda")) / 16).bfloat16()
o = torch.empty_like(v)
libs = []
for path in sys.argv[1:]:
    l = C.CDLL(path)
    l.lina_gla_chunk_fwd.restype = C.c_int
    l.lina_gla_chunk_fwd.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_float, C.c_void_p]
    libs.append((path, l))
def run(l, n):
    for _ in range(n):
        rc = l.lina_gla_chunk_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), gk.data_ptr(), None, 0, o.data_ptr(), None, None,
                                  B, H, T, K, V, 1, K ** -0.5, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
for rnd in range(3):
    for path, l in libs:
        run(l, 3); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(l, 20); e1.record(); torch.cuda.synchronize()
        print(f"round {rnd} {path.split('/')[-1]:28s} {e0.elapsed_time(e1) / 20:.4f} ms")
