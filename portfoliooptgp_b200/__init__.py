"""portfoliooptgp_b200 -- B200-native (sm_100a CUDA, fp64) exact-GP / SVGP engine behind the
GPflow model API that LUOJIUzxy/PortfolioOptGP drives.  Use it as::

    import portfoliooptgp_b200 as gpflow

and the reference call sites (GPR/model_trainer.py, GPR/predictor.py,
Multi-Input_GPR/models/model_trainer.py, test_scripts/SVGP.py) run unchanged apart from the
import and from passing numpy / torch arrays instead of tf tensors.  There is no CPU fallback."""
from . import config, kernels, likelihoods, mean_functions, models, optimizers, utilities  # noqa: F401
from . import inducing_variables  # noqa: F401
from .base import Module, Parameter, set_trainable  # noqa: F401
from .config import default_float, default_jitter  # noqa: F401
from ._capi import CholeskyError, EngineError  # noqa: F401

__version__ = "0.1.0"
from . import batched  # noqa: E402,F401
from .batched import BatchedGPR, lockstep_lbfgsb  # noqa: E402,F401
from . import data_prep, postprocess  # noqa: E402,F401
from . import mean_functions as functions  # noqa: E402,F401  (gpflow.functions alias)
from . import trainers  # noqa: E402,F401
from .trainers import GPRModelTrainer, MultiInputModelTrainer, fit_concurrently  # noqa: E402,F401
