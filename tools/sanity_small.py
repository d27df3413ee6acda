"""Small end-to-end pass over every kernel family (for compute-sanitizer runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import portfoliooptgp_b200 as gpflow
from tests.helpers import make_multi_input
X, Y = make_multi_input(1, 300, 4)
K = gpflow.kernels
k = K.Exponential(active_dims=slice(0, 3)) * K.Exponential(active_dims=[3])
m = gpflow.models.GPR((X, Y), kernel=k, noise_variance=0.05)
print("gpr", m.lml_and_constrained_grads()[0], float(m.predict_f(X[:20])[1].sum()))
k2 = K.SquaredExponential() + K.Matern52() + K.Linear()
m2 = gpflow.models.GPR((X[:130], Y[:130]), kernel=k2, noise_variance=0.05)
print("gpr2", m2.lml_and_constrained_grads()[0])
Xb = np.stack([X[i:i + 100] for i in range(3)]); Yb = np.stack([Y[i:i + 100, 0] for i in range(3)])
b = gpflow.BatchedGPR(Xb, Yb, k, noise_variance=0.1)
print("batched", b.lml_and_grads()[0], float(b.predict_f(Xb[:, :5])[0].sum()))
sv = gpflow.models.SVGP(kernel=K.SquaredExponential(), likelihood=gpflow.likelihoods.Gaussian(0.05), inducing_variable=X[:20].copy(), num_data=300)
print("svgp", sv.training_loss_closure((X, Y)).value_and_grads(sv.trainable_variables)[0])
sg = gpflow.models.SGPR((X, Y), kernel=K.Matern32(), inducing_variable=X[:20].copy(), noise_variance=0.05)
print("sgpr", sg.training_loss_closure().value_and_grads(sg.trainable_variables)[0], float(sg.predict_f(X[:9])[0].sum()))
from portfoliooptgp_b200 import data_prep
r = data_prep.returns(np.abs(X[:, :2]) + 1.0)
Xd, _, _ = data_prep.design_matrix([r, np.arange(300.0)])
print("prep", float(data_prep.rolling_windows(Xd, r[:, 0].contiguous(), window=64, stride=16)[0].sum()))
torch.cuda.synchronize()
