"""Times the tensor-core GEMM vs the FFMA GEMM on the step's dominant shapes (CUDA events, warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
from sheeprl_b200.lib import CudaOps
cu = CudaOps("cuda")
def bench(M, N, K, reps=20):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    res = {}
    for name, fn in (("tc", cu.lib.b200rl_gemm_tc), ("ffma", None)):
        def call():
            if fn is None:
                os.environ  # noqa
                cu.lib.b200rl_gemm_f32(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(C.data_ptr()), None, M, N, K, K + 0, K, N, 0, 1, 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            else:
                fn(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(C.data_ptr()), None, M, N, K, K, K, N, 0, 1, 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        for _ in range(3): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = (ms, 2.0 * M * N * K / ms / 1e9)
        if fn is not None:
            ref = A.double() @ B.double().t()
            res["err"] = float((C.double() - ref).abs().max() / ref.abs().max())
    print(f"M{M} N{N} K{K}: tc {res['tc'][0]*1e3:.1f} us {res['tc'][1]:.1f} TF/s (rel err {res['err']:.2e}) | dispatch {res['ffma'][0]*1e3:.1f} us {res['ffma'][1]:.1f} TF/s")
for shp in ((16384, 512, 1536), (16384, 512, 512), (15360, 255, 512), (1024, 4096, 1536), (1024, 512, 4096), (1024, 1536, 1024)):
    bench(*shp)
