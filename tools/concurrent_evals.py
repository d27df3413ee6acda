"""Two (or more) independent LML+gradient evaluations in flight on ONE GPU (restarts / kernel candidates,
models/model_trainer.py:26-48, GPR/main.py:105-114): one engine handle, one CUDA stream and one host thread each.
Does the latency-bound bottom of one factorisation hide behind the bulk products of the other?"""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import portfoliooptgp_b200 as gpflow
from portfoliooptgp_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 10
out = {}
counts = tuple(int(c) for c in sys.argv[3].split(",")) if len(sys.argv) > 3 else (1, 2, 3, 4)
for k_threads in counts:
    models, streams = [], []
    for t in range(k_threads):
        X, Y = bench.make_c2(seed=2 + t, n=n)
        k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Matern52() + gpflow.kernels.Linear()
        m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=1e-2)
        m._engine = _capi.Engine(0)          # its own handle: own workspaces, own side streams
        models.append(m)
        streams.append(torch.cuda.Stream())

    def work(i, reps):
        with torch.cuda.stream(streams[i]):
            for _ in range(reps):
                models[i].lml_and_constrained_grads()

    for i in range(k_threads):
        work(i, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i, evals)) for i in range(k_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[k_threads] = {"evals_per_s": k_threads * evals / dt, "ms_per_eval_aggregate": 1e3 * dt / (k_threads * evals)}
    del models
    torch.cuda.empty_cache()
print(json.dumps({"N": n, "concurrent": out}))
