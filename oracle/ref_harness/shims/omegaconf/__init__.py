"""Import shim so the UNMODIFIED reference modules import without omegaconf (absent from this image).

Test infrastructure only.  Re-exports the product's config nodes, which implement the slice of
the omegaconf API the reference touches (engine.py:6-7,20,36-37; optimizer.py:51).
"""
from mcluminescence_b200.config import DictConfig, ListConfig, OmegaConf  # noqa: F401
