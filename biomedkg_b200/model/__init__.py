"""Mirror of biomedkg/model/__init__.py for the GCL path (KGE encoders/decoders are out of scope)."""
from .encoder import GATConv, GATEncoder, GCNConv, GCNEncoder
from .gcl import DGI, GGD, GRACE

__all__ = ["GATConv", "GATEncoder", "GCNConv", "GCNEncoder", "DGI", "GRACE", "GGD"]
