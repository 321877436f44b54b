"""Host-side launcher: torch tensors as buffers, the C ABI for the work.

``run_replicas`` is the one place Python touches the device path: it allocates the output
tensors and the scratch slab with torch (device memory + the current CUDA stream), fills a
``mcl_run_args`` and calls ``mcl_run``.  Everything that the reference does per replica inside
``simulate()`` / ``TL_lab`` / ``ISO_lab`` (``src/class/simulate.py:46-92``,
``src/class/tl_trap_lab.py:75-111,135-172``) happens inside that call, for all replicas at once.

Two random-number modes (see ``include/mcl_b200.h``):

* ``MODE_PHILOX`` -- native counter-based streams, FP32/SFU kernel (the product path);
* ``MODE_REPLAY`` -- consumes NumPy's legacy MT19937 uniforms in the reference's draw order,
  FP64; per-step integer traces are bit-exact against the reference.  Because the reference
  never re-seeds between replicas (``simulate.py:46-48``), replicas are chained: replica r+1
  starts where replica r stopped consuming.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np

from . import _native
from .replicas import MODE_PHILOX, MODE_REPLAY, REPLICA_DTYPE, SEGMENT_DTYPE, STATUS_MESSAGES

AXIS_TIME_LIN, AXIS_TIME_LOG, AXIS_TEMP = 0, 1, 2


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _native.NativeError(
            "no CUDA device: the kinetics path is sm_100a CUDA only and has no CPU fallback")
    return torch


@dataclass
class HistSpec:
    axis: int
    n_bins: int
    lo: float
    hi: float
    n_groups: int = 1


@dataclass
class RunResult:
    """Outputs of one launch; tensors live on the device until ``.host()`` is called."""
    n_replicas: int
    max_steps: int
    tensors: Dict[str, "object"] = field(default_factory=dict)
    _host: Dict[str, np.ndarray] = field(default_factory=dict)

    def host(self, name: str) -> np.ndarray:
        if name not in self._host:
            self._host[name] = self.tensors[name].cpu().numpy()
        return self._host[name]

    def __getattr__(self, name):
        if name.startswith("_") or name in ("tensors", "n_replicas", "max_steps"):
            raise AttributeError(name)
        if name in self.tensors:
            return self.host(name)
        raise AttributeError(name)

    def raise_on_error(self) -> None:
        st = self.host("status")
        bad = np.nonzero(st)[0]
        if bad.size:
            r, code = int(bad[0]), int(st[bad[0]])
            msg = STATUS_MESSAGES.get(code, f"status {code}")
            if code == -1:
                # what the reference raises when a replica needs more than `steps` records
                raise IndexError(f"index {self.max_steps} is out of bounds for axis 0 with size "
                                 f"{self.max_steps} (replica {r}: {msg})")
            if code == -5:
                raise IndexError(f"list index out of range (replica {r}: {msg})")
            raise _native.NativeError(f"replica {r}: {msg}")


_ws_cache: Dict[int, "object"] = {}


def _workspace(torch, device, nbytes: int):
    key = device.index if device.index is not None else torch.cuda.current_device()
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        _ws_cache.pop(key, None)
        buf = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def release_workspace() -> None:
    _ws_cache.clear()


def run_replicas(replicas: np.ndarray, segments: np.ndarray, max_steps: int, *,
                 mode: int = MODE_PHILOX, seed: int = 0, replica_id0: int = 0,
                 obs_time: Optional[np.ndarray] = None,
                 replay_u=None, replay_off=None,
                 trace: bool = True, structure: bool = False,
                 hist: Optional[HistSpec] = None, hist_group: Optional[np.ndarray] = None,
                 hist_out: Optional[Dict[str, "object"]] = None,
                 device=None, sync: bool = False) -> RunResult:
    """Launch one batch of independent replicas on the current CUDA device."""
    torch = _torch()
    L = _native.load()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    reps = np.ascontiguousarray(replicas)
    segs = np.ascontiguousarray(segments)
    assert reps.dtype.itemsize == REPLICA_DTYPE.itemsize and segs.dtype.itemsize == SEGMENT_DTYPE.itemsize
    R = int(reps.shape[0])
    obs = np.ascontiguousarray(obs_time if obs_time is not None else np.zeros(0), dtype=np.float64)

    with torch.cuda.device(dev):
        a = _native.RunArgs()
        a.replicas, a.n_replicas = reps.ctypes.data, R
        a.segments, a.n_segments = segs.ctypes.data, int(segs.shape[0])
        a.obs_time, a.n_obs = (obs.ctypes.data if obs.size else None), int(obs.size)
        a.max_steps, a.mode = int(max_steps), int(mode)
        a.seed, a.replica_id0 = int(seed) & (2 ** 64 - 1), int(replica_id0)
        out = RunResult(R, int(max_steps))
        T = out.tensors

        def dev_i32(*shape):
            return torch.zeros(shape, dtype=torch.int32, device=dev)

        if trace:
            T["event"], T["n_e"] = dev_i32(R, max_steps), dev_i32(R, max_steps)
            T["t"] = torch.zeros((R, max_steps), dtype=torch.float64, device=dev)
            a.event, a.n_e, a.t = T["event"].data_ptr(), T["n_e"].data_ptr(), T["t"].data_ptr()
        if structure:
            T["kind"], T["e_idx"], T["h_idx"] = dev_i32(R, max_steps), dev_i32(R, max_steps), dev_i32(R, max_steps)
            a.kind, a.e_idx, a.h_idx = T["kind"].data_ptr(), T["e_idx"].data_ptr(), T["h_idx"].data_ptr()
        T["steps_used"], T["final_n_e"], T["status"] = dev_i32(R), dev_i32(R), dev_i32(R)
        T["esteps"] = torch.zeros(R, dtype=torch.int64, device=dev)
        T["consumed"] = torch.zeros(R, dtype=torch.int64, device=dev)
        T["obs_n_e"] = torch.full((max(int(obs.size), 1),), -1, dtype=torch.int32, device=dev)
        a.steps_used, a.final_n_e, a.status = T["steps_used"].data_ptr(), T["final_n_e"].data_ptr(), T["status"].data_ptr()
        a.esteps, a.consumed, a.obs_n_e = T["esteps"].data_ptr(), T["consumed"].data_ptr(), T["obs_n_e"].data_ptr()

        keep = [reps, segs, obs]
        if mode == MODE_REPLAY:
            if replay_u is None or replay_off is None:
                raise ValueError("replay mode needs replay_u and replay_off")
            u = replay_u if hasattr(replay_u, "data_ptr") else torch.as_tensor(np.ascontiguousarray(replay_u, dtype=np.float64)).to(dev)
            off = replay_off if hasattr(replay_off, "data_ptr") else torch.as_tensor(np.ascontiguousarray(replay_off, dtype=np.int64)).to(dev)
            keep += [u, off]
            a.replay_u, a.replay_off = u.data_ptr(), off.data_ptr()

        hs = None
        if hist is not None:
            hs = _native.HistSpec(int(hist.axis), int(hist.n_bins), int(hist.n_groups), 0, float(hist.lo), float(hist.hi))
            a.hist = C.pointer(hs)
            if hist_group is not None:
                grp = np.ascontiguousarray(hist_group, dtype=np.int32)
                keep.append(grp)
                a.hist_group = grp.ctypes.data
            for name in ("hist_events", "hist_occ", "hist_occ_sq"):
                if hist_out is not None and name in hist_out:
                    T[name] = hist_out[name]
                else:
                    T[name] = torch.zeros((hist.n_groups, hist.n_bins), dtype=torch.int64, device=dev)
            a.hist_events, a.hist_occ, a.hist_occ_sq = (T["hist_events"].data_ptr(), T["hist_occ"].data_ptr(),
                                                       T["hist_occ_sq"].data_ptr())

        need = L.mcl_workspace_bytes(C.byref(a))
        if need == 0:
            raise _native.NativeError(f"mcl_workspace_bytes: {_native.last_error()}")
        ws = _workspace(torch, dev, need)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        _native.check(L.mcl_run(C.byref(a)), "mcl_run")
        if sync:
            torch.cuda.current_stream(dev).synchronize()
        out._keep = keep  # noqa: SLF001  (host tables must outlive the async copies)
    return out


# ---------------------------------------------------------------------------------------
# Replay stream: NumPy's legacy global generator, consumed in the reference's order
# ---------------------------------------------------------------------------------------
class ReplayStream:
    """The uniforms ``np.random.seed(seed)`` would hand to the reference, with a cursor."""

    def __init__(self, seed: int):
        self._rs = np.random.RandomState(int(seed))
        self._buf = np.zeros(0)
        self._base = 0          # absolute index of _buf[0]
        self.pos = 0            # absolute cursor

    def window(self, n: int) -> np.ndarray:
        """Uniforms [pos, pos+n) without consuming them."""
        end = self.pos + n
        have = self._base + self._buf.size
        if end > have:
            if self.pos > have:          # the cursor was advanced past what has been generated
                self._rs.random_sample(self.pos - have)
                self._buf, self._base, have = np.zeros(0), self.pos, self.pos
            grow = max(end - have, 1 << 20)
            self._buf = np.concatenate([self._buf[self.pos - self._base:], self._rs.random_sample(grow)])
            self._base = self.pos
        s = self.pos - self._base
        return self._buf[s:s + n]

    def advance(self, n: int) -> None:
        self.pos += int(n)


_global_replay: Optional[ReplayStream] = None


def seed_replay(seed: int) -> ReplayStream:
    """Equivalent of ``np.random.seed(seed)`` for replay mode (process-global, like NumPy's)."""
    global _global_replay
    _global_replay = ReplayStream(seed)
    return _global_replay


def global_replay() -> ReplayStream:
    if _global_replay is None:
        raise RuntimeError("replay mode: call mcluminescence_b200.seed_replay(seed) first")
    return _global_replay


def run_replay_chained(replicas: np.ndarray, segments: np.ndarray, max_steps: int, stream: ReplayStream, *,
                       obs_time: Optional[np.ndarray] = None, structure: bool = False,
                       device=None) -> Dict[str, np.ndarray]:
    """Replay replicas one after another on ONE continuing stream (reference semantics).

    Each replica is one launch; its ``consumed`` count positions the next one.  If the guessed
    window of uniforms is too short the replica is re-run with a window twice as long.
    """
    R = int(replicas.shape[0])
    n_obs = int(obs_time.size) if obs_time is not None else 0
    res = {
        "event": np.zeros((R, max_steps), np.int32), "n_e": np.zeros((R, max_steps), np.int32),
        "t": np.zeros((R, max_steps), np.float64), "steps_used": np.zeros(R, np.int32),
        "final_n_e": np.zeros(R, np.int32), "esteps": np.zeros(R, np.int64),
        "consumed": np.zeros(R, np.int64), "status": np.zeros(R, np.int32),
        "obs_n_e": np.full(max(n_obs, 1), -1, np.int32),
    }
    if structure:
        for k in ("kind", "e_idx", "h_idx"):
            res[k] = np.zeros((R, max_steps), np.int32)
    for r in range(R):
        rp = replicas[r:r + 1]
        n_e0, n_h0, N_e = int(rp["n_e0"][0]), int(rp["n_h0"][0]), int(rp["N_e"][0])
        guess = 3 * (n_e0 + n_h0) + 512 * (2 * max(N_e, n_e0) + 7) + 64
        while True:
            u = stream.window(guess)
            off = np.array([0, guess], dtype=np.int64)
            out = run_replicas(rp, segments, max_steps, mode=MODE_REPLAY, obs_time=obs_time,
                               replay_u=u, replay_off=off, trace=True, structure=structure,
                               device=device, sync=True)
            st = int(out.status[0])
            if st == -3:
                guess *= 2
                continue
            break
        for k in ("event", "n_e", "t") + (("kind", "e_idx", "h_idx") if structure else ()):
            res[k][r] = out.host(k)[0]
        for k in ("steps_used", "final_n_e", "esteps", "consumed", "status"):
            res[k][r] = out.host(k)[0]
        if n_obs:
            ob, oc = int(rp["obs_begin"][0]), int(rp["obs_count"][0])
            res["obs_n_e"][ob:ob + oc] = out.host("obs_n_e")[ob:ob + oc]
        if st != 0:
            break
        stream.advance(int(out.consumed[0]))
    return res


def launch_count() -> int:
    """Kernels libmcl_b200.so has launched in this process so far (bench.py's `gpu_launches`)."""
    return int(_native.load().mcl_launch_count())


def device_peaks() -> Dict[str, float]:
    """Measured issue-rate peaks of the current device (for the roofline denominators)."""
    _torch()
    L = _native.load()
    pk = _native.Peaks()
    _native.check(L.mcl_device_peaks(C.byref(pk)), "mcl_device_peaks")
    return {k: getattr(pk, k) for k, _ in _native.Peaks._fields_ if k != "reserved"}
