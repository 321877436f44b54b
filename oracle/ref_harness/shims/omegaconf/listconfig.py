from mcluminescence_b200.config import ListConfig  # noqa: F401
