"""B200-native quadrature-point engine behind the dolfinx-external-operator API.

Public names mirror the reference package (`src/dolfinx_external_operator/__init__.py:1-19`)
for the hot path - `evaluate_operands`, `evaluate_external_operators` - and re-export its
symbolic layer unchanged; the constitutive callables of the reference demos are provided as
GPU-backed `external_function` factories (`VonMises`, `HeatConductivity`, `HeatFlux`, ...).
"""

from ._lib import EOError, LIB_PATH  # noqa: F401
from .context import Context, DeviceArray, default_context  # noqa: F401
from .constitutive import HeatConductivity, HeatFlux, MohrCoulomb, VonMises  # noqa: F401
from .isihara import Isihara, register_torch_op  # noqa: F401
from .tabulation import GeneralTabulator, Tabulator  # noqa: F401
from .forms import QuadratureForms  # noqa: F401
from .jit import JitModel  # noqa: F401
from . import jit_models  # noqa: F401
from .external_operator import (  # noqa: F401
    FEMExternalOperator,
    evaluate_external_operators,
    evaluate_operands,
    replace_external_operators,
)

__version__ = "0.1.0"
