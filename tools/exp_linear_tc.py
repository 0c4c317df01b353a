import sys, torch
sys.path.insert(0, '.')
from keisei_b200 import _lib, model_ops
dev = torch.device('cuda:0')
lib = _lib.load()
def bench(M, K, N, reps=50, interleave=False):
    Kp, Np, Nb = (K + 63)//64*64, (N + 127)//128*128, (N + 63)//64*64
    xb = torch.randn(M, Kp, device=dev).bfloat16()
    wp = torch.randn(Np, Kp, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    yf = torch.empty(M, N, device=dev); yb = torch.empty(M, Nb, device=dev, dtype=torch.bfloat16)
    small = torch.zeros(1024, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        def call():
            lib.kb_linear_tc(xb.data_ptr(), M, Kp, wp.data_ptr(), N, Np, None, bias.data_ptr(), 1, yf.data_ptr(), N, yb.data_ptr(), Nb, Nb, 0, 0, 148, s.cuda_stream)
            if interleave: small.add_(1.0)
        for _ in range(3): call()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): call()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); g.replay(); e1.record(s); s.synchronize()
    print(f"M={M} K={K} N={N} interleave={interleave}: {e0.elapsed_time(e1)/(2*reps)*1000:.1f} us/iter (graph replay)")
for shp in [(4096,768,128),(4096,128,256),(4096,16,512),(512,768,128)]:
    bench(*shp)
    bench(*shp, interleave=True)
