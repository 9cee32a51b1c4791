"""biomedkg_b200 - B200-native (sm_100a) implementation of BioMedKG's GCL training step.

Mirrors the reference's module surface for that path only:
    biomedkg.model            -> biomedkg_b200.model            (GCNEncoder, GRACE, DGI, GGD)
    biomedkg.utils.fusion     -> biomedkg_b200.utils.fusion     (AttentionFusion, ReDAF)
    biomedkg.factory          -> biomedkg_b200.factory          (FusionFactory)
    biomedkg.gcl_module       -> biomedkg_b200.gcl_module       (BaseGCL, GRACEModule, DGIModule, GGDModule)
Importing the package loads libbmkg_b200.so and fails loudly if it is missing.
"""
from . import _cabi  # noqa: F401  (raises ImportError if the kernel library was not built)
from . import ops  # noqa: F401
from . import export  # noqa: F401
from .factory import FusionFactory
from .gcl_module import BaseGCL, DGIModule, GGDModule, GRACEModule
from .model import DGI, GGD, GRACE, GCNConv, GCNEncoder

__all__ = ["FusionFactory", "BaseGCL", "DGIModule", "GGDModule", "GRACEModule", "DGI", "GGD", "GRACE", "GCNConv", "GCNEncoder", "ops", "export"]
