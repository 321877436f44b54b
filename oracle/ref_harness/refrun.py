"""Run the UNMODIFIED reference (``/root/reference/src/class``) under seeded, recorded conditions.

TEST INFRASTRUCTURE ONLY -- used to generate the golden vectors under ``tests/golden/`` and to
validate the C restatement in ``oracle/``.  It works only in the build container (the reference
tree does not exist on the GPU box); nothing in the product imports it.

What it does:
  * puts the import shims for the absent hydra / omegaconf / matplotlib packages on ``sys.path``;
  * imports the reference's own ``engine``, ``tl_trap_lab``, ``simulate`` and ``optimizer`` modules;
  * redirects ``simulate.PROJECT_ROOT`` (CSV output) because the reference tree is read-only;
  * wraps ``Box.seed/add_electron/remove_pair`` to log the structural events
    ``(kind, electron index, hole index)`` and wraps ``np.random.rand/exponential`` to count draws.
"""
from __future__ import annotations

import os
import sys
import tempfile
from contextlib import contextmanager
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

REFERENCE_ROOT = os.environ.get("MCL_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.abspath(os.path.join(_HERE, "..", ".."))

_mods: Dict[str, Any] = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "class", "engine.py"))


def load():
    """Import the reference modules (once)."""
    if _mods:
        return _mods
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (os.path.join(REFERENCE_ROOT, "src", "class"), os.path.join(_HERE, "shims"), _REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    import engine  # type: ignore
    import tl_trap_lab  # type: ignore
    import simulate  # type: ignore
    import optimizer  # type: ignore

    out_root = tempfile.mkdtemp(prefix="mcl_ref_out_")
    os.makedirs(os.path.join(out_root, "results", "simulations"), exist_ok=True)
    simulate.PROJECT_ROOT = out_root
    _mods.update(engine=engine, tl_trap_lab=tl_trap_lab, simulate=simulate, optimizer=optimizer,
                 out_root=out_root)
    return _mods


def compose(overrides: Optional[Sequence[str]] = None):
    """Compose from the REFERENCE's own conf/ directory."""
    from mcluminescence_b200.config import compose as _compose
    return _compose("config_fp", overrides, config_dir=os.path.join(REFERENCE_ROOT, "conf"))


class Recorder:
    """Event + draw-count log for one seeded reference call."""

    def __init__(self):
        self.events: List[tuple] = []      # (kind, e_idx, h_idx); kind 0=seed,1=fill,2=recomb
        self.n_rand_calls = 0
        self.n_rand_values = 0
        self.n_exp_calls = 0
        self.n_exp_values = 0

    @property
    def n_uniforms(self) -> int:
        return self.n_rand_values + self.n_exp_values


@contextmanager
def recording():
    m = load()
    Box = m["engine"].Box
    rec = Recorder()
    o_seed, o_add, o_rm = Box.seed, Box.add_electron, Box.remove_pair
    o_rand, o_exp = np.random.rand, np.random.exponential

    def seed(self, n_e, n_h):
        rec.events.append((0, int(n_e), int(n_h * self.boundary_factor ** 3)))
        return o_seed(self, n_e, n_h)

    def add(self):
        rec.events.append((1, self.electrons.shape[0], self.holes.shape[0]))
        return o_add(self)

    def rm(self, e_idx, h_idx):
        rec.events.append((2, int(e_idx), int(h_idx)))
        return o_rm(self, e_idx, h_idx)

    def rand(*shape):
        rec.n_rand_calls += 1
        rec.n_rand_values += int(np.prod(shape)) if shape else 1
        return o_rand(*shape)

    def exponential(scale=1.0, size=None):
        rec.n_exp_calls += 1
        rec.n_exp_values += int(np.size(scale)) if size is None else int(np.prod(size))
        return o_exp(scale, size)

    Box.seed, Box.add_electron, Box.remove_pair = seed, add, rm
    np.random.rand, np.random.exponential = rand, exponential
    try:
        yield rec
    finally:
        Box.seed, Box.add_electron, Box.remove_pair = o_seed, o_add, o_rm
        np.random.rand, np.random.exponential = o_rand, o_exp


def run_simulate(overrides: Sequence[str], seed: int):
    """``np.random.seed(seed); simulate(cfg)`` on the genuine reference."""
    m = load()
    cfg = compose(overrides)
    with recording() as rec:
        np.random.seed(seed)
        x_ax, lum, e_ratio, configs = m["simulate"].simulate(cfg)
    return dict(x_ax=x_ax, lum=lum, e_ratio=e_ratio, configs=configs, rec=rec, cfg=cfg)


def run_objective(p: Sequence[float], exp: str, seed: int,
                  overrides: Sequence[str] = ("exp_type_fp=TLlab", "physics_fp=lab_TL")):
    m = load()
    cfg = compose(list(overrides))
    with recording() as rec:
        np.random.seed(seed)
        import io
        from contextlib import redirect_stdout
        buf = io.StringIO()
        with redirect_stdout(buf):
            val = m["optimizer"].objective(np.asarray(p, dtype=float), cfg, exp)
    return dict(value=float(val), rec=rec, printed=buf.getvalue())


def run_lab(exp: str, seed: int,
            overrides: Sequence[str] = ("exp_type_fp=TLlab", "physics_fp=lab_TL")):
    """``run_one_sim(cfg, exp)`` at the YAML defaults (no parameter vector)."""
    m = load()
    cfg = compose(list(overrides))
    with recording() as rec:
        np.random.seed(seed)
        import io
        from contextlib import redirect_stdout
        buf = io.StringIO()
        with redirect_stdout(buf):
            val = m["optimizer"].run_one_sim(cfg, exp)
    return dict(value=float(val), rec=rec, printed=buf.getvalue())


# ---------------------------------------------------------------------------------------
# Legacy (pre-refactor) code: src/est_params/functions.py, which produced results/lab_sims/*.csv
# ---------------------------------------------------------------------------------------
_legacy = {}


def load_legacy():
    """Import the reference's UNMODIFIED legacy module ``src/est_params/functions.py`` (once)."""
    if _legacy:
        return _legacy["F"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import importlib.util
    for p in (os.path.join(_HERE, "shims"), _REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the legacy package has its own `paths` module with the same name as src/class/paths.py: load both explicitly
    d = os.path.join(REFERENCE_ROOT, "src", "est_params")
    saved = sys.modules.pop("paths", None)
    sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location("mcl_legacy_functions", os.path.join(d, "functions.py"))
        F = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(F)
    finally:
        sys.path.remove(d)
        sys.modules.pop("paths", None)
        if saved is not None:
            sys.modules["paths"] = saved
    _legacy["F"] = F
    return F


def run_legacy_tl(p, seed: int, lab_data: str = "CLBR_IRSL50_0.25KperGy",
                  overrides: Sequence[str] = ("exp_type_fp=TLlab", "physics_fp=lab_TL")):
    """``np.random.seed(seed); sim_lab_TL_residuals(run_cfg, lab_data)`` on the genuine legacy code, with the number
    of electron additions / recombinations / uniforms per lab row captured by wrapping its helpers."""
    import io
    from contextlib import redirect_stdout
    from mcluminescence_b200.config import initialize_runs
    from mcluminescence_b200.optimizer import cfg_with_params
    F = load_legacy()
    cfg = compose(list(overrides))
    if p is not None:
        cfg = cfg_with_params(cfg, np.asarray(p, dtype=float))
    run = initialize_runs(cfg)[0]
    rows = []                       # per lab row: [e0, holes_n, adds, recombinations, uniforms]
    o_init, o_add, o_rec = F.initialize_box_bg, F.add_electron, F.recomber
    o_rand, o_exp = np.random.rand, np.random.exponential

    def init(cfg_, e_ratio_start=0):
        e, h, dim = o_init(cfg_, e_ratio_start)
        rows.append([e.shape[0], h.shape[0], 0, 0, 0])
        rows[-1][4] += 3 * (e.shape[0] + h.shape[0])         # drawn inside o_init through the wrapped rand
        return e, h, dim

    def add(*a, **k):
        rows[-1][2] += 1
        return o_add(*a, **k)

    def rec(*a, **k):
        rows[-1][3] += 1
        return o_rec(*a, **k)

    def rand(*shape):
        if rows and not _in_init[0]:
            rows[-1][4] += int(np.prod(shape)) if shape else 1
        return o_rand(*shape)

    def exponential(scale=1.0, size=None):
        if rows:
            rows[-1][4] += int(np.size(scale)) if size is None else int(np.prod(size))
        return o_exp(scale, size)

    _in_init = [False]

    def init_guarded(cfg_, e_ratio_start=0):
        _in_init[0] = True
        try:
            return init(cfg_, e_ratio_start)
        finally:
            _in_init[0] = False

    F.initialize_box_bg, F.add_electron, F.recomber = init_guarded, add, rec
    np.random.rand, np.random.exponential = rand, exponential
    try:
        np.random.seed(seed)
        with redirect_stdout(io.StringIO()):
            val = F.sim_lab_TL_residuals(run, lab_data)
    finally:
        F.initialize_box_bg, F.add_electron, F.recomber = o_init, o_add, o_rec
        np.random.rand, np.random.exponential = o_rand, o_exp
    return dict(value=float(val), rows=np.asarray(rows, dtype=np.int64))
