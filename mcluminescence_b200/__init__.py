"""B200-native trapped-charge kinetics path of MCLuminescence (hot path only).

Module layout mirrors the reference's ``src/class`` so call sites translate one to one:
  ``mcluminescence_b200.simulate.simulate(cfg)``     -> ``(x_ax, Lum, electron_ratio, configs)`` + CSV
  ``mcluminescence_b200.tl_trap_lab.TLTrapSim(cfg)`` -> ``.TL_lab(csv) / .ISO_lab(csv)`` -> mse
  ``mcluminescence_b200.optimizer.objective``        (+ ``objective_batched`` for SciPy DE)
  ``mcluminescence_b200.config.compose / initialize_runs``   Hydra-surface config helpers
"""
from .config import DictConfig, ListConfig, OmegaConf, compose, initialize_runs  # noqa: F401
from .engine import seed_replay  # noqa: F401

__all__ = ["compose", "initialize_runs", "DictConfig", "ListConfig", "OmegaConf", "seed_replay"]
