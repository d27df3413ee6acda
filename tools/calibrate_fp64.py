"""Box calibration: cuBLAS fp64 DGEMM rate via torch.matmul, as the
measured FP64 'tensor' peak that roofline fractions of the DMMA kernels are quoted against.
Comparison point only -- cuBLAS is never on the product path."""
import json, sys, time
import torch

def main():
    dev = torch.device("cuda:0")
    out = {"gpu": torch.cuda.get_device_name(0)}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        for _ in range(3):
            c = a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"dgemm_{n}_tflops"] = 2.0 * n ** 3 / best / 1e9
        # sustained: back-to-back for ~2 s
        t0 = time.time(); k = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 2.0:
            for _ in range(10):
                c = a @ b
            k += 10
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        out[f"dgemm_{n}_tflops_sustained"] = 2.0 * n ** 3 * k / e0.elapsed_time(e1) / 1e9
    # potrf via cuSOLVER for context
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    spd = a @ a.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    torch.linalg.cholesky(spd); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.linalg.cholesky(spd); e1.record(); torch.cuda.synchronize()
    out["cusolver_potrf_8192_ms"] = e0.elapsed_time(e1)
    out["cusolver_potrf_8192_tflops"] = n ** 3 / 3 / e0.elapsed_time(e1) / 1e9
    print(json.dumps(out))
    json.dump(out, open(sys.argv[1], "w"), indent=1)

if __name__ == "__main__":
    main()
