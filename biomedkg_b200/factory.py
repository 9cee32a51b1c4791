"""biomedkg/factory.py:8-15 (FusionFactory only; KGEModelFactory is the out-of-scope KGE path)."""
from .utils.fusion import AttentionFusion, ReDAF


class FusionFactory:
    @staticmethod
    def create_fuser(method: str, embed_dim):
        if method == "attention":
            return AttentionFusion(embed_dim=embed_dim)
        if method == "redaf":
            return ReDAF(embed_dim=embed_dim)
        return None
