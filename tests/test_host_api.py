"""CPU: host-side mirror of the GPflow interface (no CUDA calls) and the C-ABI export check."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest

import portfoliooptgp_b200 as gpflow
from portfoliooptgp_b200 import _capi
from portfoliooptgp_b200.kernels import compile_kernel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/gpb200.h <-> libgpb200.so <-> the ctypes table agree (no compute call)."""
    hdr = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gpb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    from portfoliooptgp_b200 import build
    build.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in gpb200.h but not exported"
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    assert _capi.load_library().gpb_version() >= 100


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.EngineError):
        _capi.Engine(0)
    with pytest.raises(_capi.EngineError):
        gpflow.models.GPR((np.zeros((3, 1)), np.zeros((3, 1))), kernel=gpflow.kernels.SquaredExponential())


def test_spec_struct_layout_matches_header():
    assert ctypes.sizeof(_capi.GpbGroup) == 16 and ctypes.sizeof(_capi.GpbLeaf) == 20
    assert ctypes.sizeof(_capi.GpbTerm) == 4 * (1 + _capi.GPB_MAX_FACTORS)
    assert ctypes.sizeof(_capi.GpbKernelSpec) == 20 + 16 * 8 + 20 * 8 + 20 * 8


def test_parameter_softplus_semantics():
    p = gpflow.Parameter(1.0, transform=gpflow.utilities.positive())
    assert p.unconstrained_variable.numpy() == pytest.approx(np.log(np.e - 1.0), rel=1e-15)
    p.assign(1e-5)
    assert float(p.numpy()) == pytest.approx(1e-5, rel=1e-12)
    lik = gpflow.likelihoods.Gaussian(1e-3)
    u = lik.variance.unconstrained_variable.numpy()
    assert float(np.log1p(np.exp(u)) + 1e-6) == pytest.approx(1e-3, rel=1e-12)   # softplus + shift 1e-6
    with pytest.raises(ValueError):
        gpflow.likelihoods.Gaussian(1e-7)
    with pytest.raises(ValueError):
        p.assign(-1.0)
    # arithmetic / array protocol on the constrained value
    p.assign(2.0)
    assert float(p * 3) == pytest.approx(6.0) and float(np.sqrt(p)) == pytest.approx(np.sqrt(2.0))


def test_trainable_variable_order_matches_tf_module_traversal():
    k = gpflow.kernels.Exponential() + gpflow.kernels.Periodic(gpflow.kernels.SquaredExponential()) + gpflow.kernels.Linear()
    names = [n for n, _ in k.named_parameters()]
    assert names == ["kernels[0].lengthscales", "kernels[0].variance", "kernels[1].base_kernel.lengthscales",
                     "kernels[1].base_kernel.variance", "kernels[1].period", "kernels[2].variance"]
    rq = gpflow.kernels.RationalQuadratic()
    assert [n for n, _ in rq.named_parameters()] == ["alpha", "lengthscales", "variance"]
    gpflow.set_trainable(k.kernels[0].variance, False)
    assert len(k.trainable_variables) == 5 and len(k.variables) == 6
    gpflow.set_trainable(k, False)
    assert len(k.trainable_variables) == 0


def test_sum_product_flattening_and_sharing():
    K = gpflow.kernels
    a, b, c = K.SquaredExponential(), K.Matern12(), K.Linear()
    s = a + b + c
    assert isinstance(s, K.Sum) and len(s.kernels) == 3
    p = (a * b) * c
    assert isinstance(p, K.Product) and len(p.kernels) == 3
    mixed = (a + b) * c
    ck = compile_kernel(mixed, 2)
    assert ck.spec.n_terms == 2 and ck.spec.n_leaves == 3 and ck.spec.terms[0].n_factors == 2
    # the same instance used twice is one leaf, one set of parameters
    twice = a + a
    ck2 = compile_kernel(twice, 1)
    assert ck2.spec.n_leaves == 1 and ck2.spec.n_terms == 2 and ck2.n_params == 2
    assert len(twice.parameters) == 2


def test_compile_reference_kernels():
    K = gpflow.kernels
    D = 8
    comp = K.Exponential(active_dims=slice(0, D - 1)) * K.Exponential(active_dims=slice(D - 1, D))  # Multi-Input_GPR/main.py:126-135
    ck = compile_kernel(comp, D)
    assert ck.spec.n_groups == 2 and ck.spec.groups[0].dim_mask == 0x7F and ck.spec.groups[1].dim_mask == 0x80
    assert ck.spec.n_terms == 1 and ck.spec.terms[0].n_factors == 2 and ck.n_params == 4
    per = K.Periodic(K.SquaredExponential())
    ck = compile_kernel(per, 1)
    assert ck.spec.groups[0].kind == _capi.GROUP_PERIODIC_SQ and ck.spec.groups[0].period_index == 2
    ck = compile_kernel(K.Periodic(K.Matern32()), 1)
    assert ck.spec.groups[0].kind == _capi.GROUP_PERIODIC_ABS
    # leaves on the same columns share one distance group
    ck = compile_kernel(K.SquaredExponential() + K.Matern52() + K.Linear(), D)
    assert ck.spec.n_groups == 2 and ck.spec.n_leaves == 3
    ard = K.SquaredExponential(lengthscales=np.ones(3))
    ck = compile_kernel(ard, 3)
    assert ck.spec.groups[0].ard_index == 0 and ck.n_params == 4 and ck.spec.leaves[0].ls_index == -1
    with pytest.raises(ValueError):
        compile_kernel(K.SquaredExponential(active_dims=[9]), 3)
    with pytest.raises(ValueError):
        compile_kernel(K.SquaredExponential(), 17)


def test_theta_and_gradient_scatter():
    K = gpflow.kernels
    k = K.SquaredExponential(variance=2.0, lengthscales=0.5) + K.Linear(variance=0.3)
    ck = compile_kernel(k, 2)
    assert np.allclose(ck.theta(), [0.5, 2.0, 0.3])
    g = ck.scatter_grad(np.array([1.0, 1.0, 1.0]))
    u = k.kernels[0].lengthscales.unconstrained_variable.numpy()
    assert g[id(k.kernels[0].lengthscales)] == pytest.approx(1.0 / (1.0 + np.exp(-u)))


def test_deepcopy_kernel_is_independent():
    k = gpflow.kernels.Exponential(active_dims=slice(0, 2)) * gpflow.kernels.Exponential(active_dims=slice(2, 3))
    k2 = copy.deepcopy(k)
    k3 = gpflow.utilities.deepcopy(k)
    k2.kernels[0].variance.assign(5.0)
    assert float(k.kernels[0].variance.numpy()) == 1.0 and float(k3.kernels[0].variance.numpy()) == 1.0
    assert k2.kernels[0].active_dims == slice(0, 2)


def test_scipy_pack_unpack():
    K = gpflow.kernels
    k = K.SquaredExponential(lengthscales=np.array([1.0, 2.0])) + K.Linear()
    vs = k.trainable_variables
    S = gpflow.optimizers.Scipy
    x = S.initial_parameters(vs)
    assert x.shape == (4,)
    S.assign_tensors(vs, x + 1.0)
    assert np.allclose(S.initial_parameters(vs), x + 1.0)
    with pytest.raises(TypeError):
        gpflow.optimizers.Scipy().minimize(lambda: 0.0, vs)


def test_print_summary_runs(capsys):
    k = gpflow.kernels.SquaredExponential() + gpflow.kernels.Linear()
    gpflow.utilities.print_summary(k)
    gpflow.utilities.print_summary(k, "notebook")
    out = capsys.readouterr().out
    assert "lengthscales" in out and "Softplus" in out
