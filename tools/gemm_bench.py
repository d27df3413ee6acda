"""DMMA GEMM micro-benchmark: TFLOP/s by shape and operand layout (engine C-ABI gpb_gemm)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from portfoliooptgp_b200 import ops

eng = ops.shared_engine(0)
ops.sync_stream(eng)
shapes = [(4096, 4096, 4096), (8192, 8192, 1024), (2048, 2048, 2048), (8192, 8192, 8192)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    for ta, tb in [(0, 1), (0, 0), (1, 0), (1, 1)]:
        for tri in ([0, 1] if (M == N and os.environ.get("BENCH_TRI", "1") == "1") else [0]):
            A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device="cuda")
            B = torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device="cuda")
            Cm = torch.zeros((M, N), dtype=torch.float64, device="cuda")
            def run():
                eng.gemm(ta, tb, M, N, K, 1.0, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], 0.0, Cm.data_ptr(), N, tri)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            tiles = (M // 128) * (N // 128)
            fl = 2.0 * M * N * K * ((tiles + M // 128) / 2 / tiles if tri else 1.0)
            print(f"M={M} N={N} K={K} ta={ta} tb={tb} tri={tri}: {best:.3f} ms  {fl / best / 1e9:.2f} TFLOP/s", flush=True)
            del A, B, Cm
