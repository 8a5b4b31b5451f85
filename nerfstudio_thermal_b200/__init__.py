"""nerfstudio_thermal_b200 -- the thermal-nerfacto per-ray hot path as sm_100a CUDA kernels behind the
reference's module API (yvette256/nerfstudio-thermal).  See DESIGN.md / INTEGRATION.md.

(The directory is spelled with an underscore because Python cannot import a hyphenated package name.)
"""
from . import ops  # noqa: F401
from ._lib import EXPORTED_SYMBOLS, LIB_PATH, TnKernelError, TnLibraryError  # noqa: F401
from .field_components import (  # noqa: F401
    MLP,
    Embedding,
    HashEncoding,
    MLPWithHashEncoding,
    SceneContraction,
    SHEncoding,
    trunc_exp,
)
from .fields import (  # noqa: F401
    FieldHeadNames,
    HashMLPDensityField,
    NerfactoField,
    ThermalNerfactoField,
)
from .model import (  # noqa: F401
    CameraOptimizer,
    CameraOptimizerConfig,
    NearFarCollider,
    ThermalNerfactoModel,
    ThermalNerfactoModelConfig,
)
from .rays import Frustums, RayBundle, RayLayout, RaySamples  # noqa: F401
from .renderers import AccumulationRenderer, DepthRenderer, RGBRenderer, RGBTRenderer  # noqa: F401
from .samplers import (  # noqa: F401
    PDFSampler,
    ProposalNetworkSampler,
    UniformLinDispPiecewiseSampler,
    UniformSampler,
)

TCNNNerfactoField = NerfactoField  # upstream-legacy name used by the north star
