"""gpflow.inducing_variables.InducingPoints (test_scripts/SVGP.py:581-583 reads model.inducing_variable.Z)."""
from .models import InducingPoints  # noqa: F401
