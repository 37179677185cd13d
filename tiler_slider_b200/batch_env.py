"""BatchedTilerSliderEnv: N independent Tiler-Slider boards advanced in lock-step on one B200.

Host-side mirror of the reference's TilerSliderEnv (explainrl/environment/environment.py:
14-194) for a batch: same vocabulary (reset / step / done / is_won / invalid_move / timeout /
max_steps), PyTorch tensors as containers, every computation done by the sm_100a kernels of
libtiler_slider.so through the C-ABI (include/tiler_slider.h).  No CPU fallback.

Packed state (see include/tiler_slider.h): `pos` is uint8[N, pos_bytes(T)], byte i of a row
= row*pos_stride + col of tile i.  `positions()` decodes to uint8[N, T, 2].
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (CAP_ALIGN, F_WON, GOAL_ORDERED, GOAL_SET, MAX_SIZE, MAX_TILES, EncodeArgs, GoalArgs, ObserveArgs,
                   StepArgs, SynthArgs, ValidArgs, check, lib)
from .puzzle import Puzzle, as_puzzle, uniform_shape

DEFAULT_REWARDS = (1.0, -0.01, -0.05)   # r_win, r_step, r_invalid -- repo-defined (the reference has no reward)


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


class BatchedTilerSliderEnv:
    """A batch of boards of one shape (size S, T tiles, one colour mode) on one CUDA device."""

    def __init__(self, size: int, n_tiles: int, n_envs: int, multi_color: bool = False, *,
                 max_steps: int = 100, auto_reset: bool = False, device: str | torch.device = "cuda",
                 rewards: Sequence[float] = DEFAULT_REWARDS, track_terminal: bool = False,
                 n_targets: int | None = None, track_flags: bool = True, host_io: bool = False):
        if not torch.cuda.is_available():
            raise _lib.TilerSliderError("BatchedTilerSliderEnv needs a CUDA device (no CPU fallback)")
        self._lib = lib()
        if not 1 <= size <= MAX_SIZE:
            raise ValueError(f"board size {size} outside 1..{MAX_SIZE}")
        if not 0 <= n_tiles <= MAX_TILES:
            raise ValueError(f"tile count {n_tiles} outside 0..{MAX_TILES}")
        if n_envs < 1:
            raise ValueError("n_envs must be positive")
        if not self._lib.ts_supported(size, n_tiles):
            raise ValueError(f"no step kernel for size {size} with {n_tiles} tiles")
        self.size, self.n_tiles, self.n_envs = int(size), int(n_tiles), int(n_envs)
        self.multi_color = bool(multi_color)
        self.n_targets = self.n_tiles if n_targets is None else int(n_targets)
        # T == 1 with one target: ordered and set equality coincide; use the cheaper compare
        self.goal_mode = GOAL_ORDERED if (self.multi_color or (self.n_tiles == 1 and self.n_targets == 1)) else GOAL_SET
        # ordered list equality with a length mismatch is never true (state.py:183-184): such a
        # batch plays normally and never reports a win (never_win); its targets do not fit the
        # packed word the step kernel compares, so the observation kernel gets them separately
        self.never_win = self.multi_color and self.n_targets != self.n_tiles
        if self.never_win and self.n_targets > MAX_TILES:
            raise ValueError(f"multi-colour boards with {self.n_targets} targets exceed the supported maximum of {MAX_TILES}")
        self.max_steps = int(max_steps)
        self.auto_reset = bool(auto_reset)
        # the status byte (is_won / invalid_move / timeout / stale) is optional with auto-reset:
        # reward and done are always written, and without it the step moves 26 instead of 27 bytes
        # per env (6x6 / 4 tiles); without auto-reset the done state of an env lives in that byte
        self.track_flags = bool(track_flags) or not self.auto_reset
        self.rewards = tuple(float(x) for x in rewards)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.capacity = _round_up(self.n_envs, CAP_ALIGN)
        self.pos_bytes = self._lib.ts_pos_bytes(self.n_tiles)
        self.board_bytes = self._lib.ts_board_bytes(self.size)
        self.board_stride = self._lib.ts_board_stride(self.size)
        self.pos_stride = self._lib.ts_pos_stride(self.size)
        self.count_bytes = 1 if self.max_steps <= (255 if self.auto_reset else 254) else 4
        cap, dev = self.capacity, self.device
        u8 = torch.uint8
        self.wide = self.size > 8
        self._walls = torch.zeros(cap * self._lib.ts_walls_bytes(self.size), dtype=u8, device=dev)
        tbytes = self.pos_bytes if self.goal_mode == GOAL_ORDERED else self._lib.ts_target_board_bytes(self.size)
        self._targets = torch.zeros(cap * tbytes, dtype=u8, device=dev)
        self._init = torch.zeros(cap, self.pos_bytes, dtype=u8, device=dev)
        # host_io (the single-env adapter): positions, actions and the step's outputs live in pinned HOST
        # memory that the kernels address directly (unified addressing), so a step is launches + one
        # stream synchronisation, with no copy calls; meant for tiny batches, where latency is all
        self.host_io = bool(host_io)

        def io(*shape, dtype=u8):
            return torch.zeros(*shape, dtype=dtype, pin_memory=True) if self.host_io else torch.zeros(*shape, dtype=dtype, device=dev)
        self._pos = io(cap, self.pos_bytes)
        self._count = torch.zeros(cap, dtype=u8 if self.count_bytes == 1 else torch.int32, device=dev)
        self._actions = io(cap)
        self._reward = io(cap, dtype=torch.float32)
        self._done = io(cap)
        self._flags = io(cap)
        self._terminal = torch.zeros(cap, self.pos_bytes, dtype=u8, device=dev) if track_terminal else None
        self._obs_targets = None      # ordered targets of a never_win batch, for observe() / target_positions()
        self._scratch_count = None
        self._scratch_flags = None
        self._scratch_reward = None
        self._host_ctx = None
        self._cached_args = None
        self._cached_out = None
        self._cached_obs_args = None
        self._io_np = None
        self._loaded = False

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_puzzles(cls, puzzles: Sequence, **kw) -> "BatchedTilerSliderEnv":
        """Batch from Puzzle objects (or anything with the ImageProcessed fields,
        dataloader.py:21-27) of one shape; boards are validated then encoded by K1."""
        ps = [as_puzzle(p).validate() for p in puzzles]
        S, T, NT, multi = uniform_shape(ps)
        n = len(ps)
        blocked = np.zeros((n, S * S), np.uint8)
        tiles = np.zeros((n, T, 2), np.uint8)
        targets = np.zeros((n, NT, 2), np.uint8)
        for i, p in enumerate(ps):
            for r, c in p.blocked_locations:
                blocked[i, int(r) * S + int(c)] = 1
            tiles[i] = np.asarray(p.initial_locations, np.uint8).reshape(T, 2)
            targets[i] = np.asarray(p.target_locations, np.uint8).reshape(NT, 2)
        return cls.from_arrays(S, blocked, tiles, targets, multi, **kw)

    @classmethod
    def from_arrays(cls, size: int, blocked, tiles, targets, multi_color: bool = False, **kw) -> "BatchedTilerSliderEnv":
        """blocked u8[N,S*S] (nonzero = wall), tiles u8[N,T,2], targets u8[N,NT,2] (row, col);
        numpy arrays or tensors on any device."""
        blocked = torch.as_tensor(blocked, dtype=torch.uint8)
        tiles = torch.as_tensor(tiles, dtype=torch.uint8)
        targets = torch.as_tensor(targets, dtype=torch.uint8)
        n, T, NT = blocked.shape[0], tiles.shape[1], targets.shape[1]
        if blocked.shape != (n, size * size) or tiles.shape != (n, T, 2) or targets.shape != (n, NT, 2):
            raise ValueError("expected blocked[N,S*S], tiles[N,T,2], targets[N,NT,2]")
        env = cls(size, T, n, multi_color, n_targets=NT, **kw)
        env.encode(blocked, tiles, targets)
        return env

    @classmethod
    def synthetic(cls, n_envs: int, size: int, n_tiles: int, n_walls: int, multi_color: bool = False, *,
                  seed: int = 0, env_index_base: int = 0, **kw) -> "BatchedTilerSliderEnv":
        """Random well-formed puzzles generated on the device by K0 (the create_simple_env
        recipe, environment.py:221-226).  `env_index_base` = global index of env 0, so a
        shard of a larger job draws the same puzzles whatever the GPU count."""
        env = cls(size, n_tiles, n_envs, multi_color, **kw)
        a = SynthArgs(size=size, n_tiles=n_tiles, n_walls=n_walls, goal_mode=env.goal_mode,
                      first_env=0, n_envs=n_envs, capacity=env.capacity, env_index_base=env_index_base,
                      seed=seed & (2 ** 64 - 1), d_walls=_ptr(env._walls), d_targets_packed=_ptr(env._targets),
                      d_init=_ptr(env._init), d_pos=_ptr(env._pos))
        with torch.cuda.device(env.device):
            check(env._lib.ts_synth(C.byref(a), env._stream()), "ts_synth")
        env._loaded = True
        env.reset()
        return env

    def encode(self, blocked: torch.Tensor, tiles: torch.Tensor, targets: torch.Tensor) -> None:
        """K1: dense description -> packed planes (GameState.__init__, state.py:61-73)."""
        dev = self.device
        b = blocked.to(dev, torch.uint8).contiguous()
        t = tiles.to(dev, torch.uint8).contiguous()
        g = targets.to(dev, torch.uint8).contiguous()
        nt = g.shape[1]
        a = EncodeArgs(size=self.size, n_tiles=self.n_tiles, n_targets=nt, goal_mode=self.goal_mode,
                       first_env=0, n_envs=self.n_envs, capacity=self.capacity,
                       d_blocked=_ptr(b), d_tiles=_ptr(t), d_targets=_ptr(g) if nt else None,
                       d_walls=_ptr(self._walls), d_targets_packed=_ptr(self._targets),
                       d_init=_ptr(self._init), d_pos=_ptr(self._pos))
        with torch.cuda.device(dev):
            check(self._lib.ts_encode(C.byref(a), self._stream()), "ts_encode")
            torch.cuda.current_stream().synchronize()   # b, t, g are temporaries
        if self.never_win:
            pw = self._lib.ts_pos_bytes(nt)
            self._obs_targets = torch.zeros(self.capacity, pw, dtype=torch.uint8, device=dev)
            self._obs_targets[: self.n_envs, :nt] = g[..., 0] * self.pos_stride + g[..., 1]
        self._loaded = True
        self.reset()

    # ------------------------------------------------------------------ episode API
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self) -> torch.Tensor:
        """TilerSliderEnv.reset (environment.py:89-97) for every env: positions <- initial,
        step_count <- 0, done <- False.  Returns the packed positions uint8[N, pos_bytes]."""
        self._require_loaded()
        self._pos.copy_(self._init)
        self._count.zero_()
        self._flags.zero_()
        self._done.zero_()
        return self._pos[: self.n_envs]

    def _step_args(self, actions_ptr: int, *, count=None, flags=None, max_steps=None, count_bytes=None,
                   auto_reset=None, raw: bool = False) -> StepArgs:
        return StepArgs(size=self.size, n_tiles=self.n_tiles, goal_mode=self.goal_mode, never_win=int(self.never_win),
                        first_env=0, n_envs=self.n_envs, capacity=self.capacity,
                        d_walls=_ptr(self._walls), d_targets_packed=_ptr(self._targets), d_init=_ptr(self._init),
                        d_pos=_ptr(self._pos), d_step_count=_ptr(self._count if count is None else count),
                        count_bytes=self.count_bytes if count_bytes is None else count_bytes,
                        max_steps=self.max_steps if max_steps is None else max_steps,
                        auto_reset=int(self.auto_reset if auto_reset is None else auto_reset),
                        d_actions=actions_ptr, r_win=self.rewards[0], r_step=self.rewards[1], r_invalid=self.rewards[2],
                        d_reward=_ptr(self._reward), d_done=None if raw else _ptr(self._done),
                        d_flags=_ptr(self._flags if flags is None else flags) if (self.track_flags or flags is not None) else None,
                        d_terminal_pos=None if raw else _ptr(self._terminal))

    def _stage_actions(self, actions) -> int:
        n = self.n_envs
        if isinstance(actions, torch.Tensor) and actions.dtype == torch.uint8 and actions.is_cuda \
                and actions.device == self.device and actions.is_contiguous() and actions.numel() >= self.capacity \
                and actions.data_ptr() % 16 == 0:   # kernels may read (and ignore) the padding envs up to capacity
            return actions.data_ptr()
        a = torch.as_tensor(actions)
        if a.numel() != n:
            raise ValueError(f"expected {n} actions, got {a.numel()}")
        self._actions[:n].copy_(a.reshape(n), non_blocking=True)
        return self._actions.data_ptr()

    def step(self, actions) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One step for every env: K2 (slide + goal check + bookkeeping + reward [+ reset]).

        actions: uint8[N] (0 UP, 1 DOWN, 2 LEFT, 3 RIGHT; state.py:31-34), ideally already on
        the device.  Returns (state, reward, done): the packed positions uint8[N,pos_bytes],
        float32[N], bool[N] -- views of buffers that the next step overwrites.  `flags`
        holds is_won / invalid_move / timeout; with auto_reset the state of a finished env is
        already its reset state (its last positions are in `terminal_pos` if tracked)."""
        self._require_loaded()
        ptr = self._stage_actions(actions)
        a = self._cached_args
        if a is None:
            a = self._cached_args = self._step_args(ptr)
            n = self.n_envs
            self._cached_out = (self._pos[:n], self._reward[:n], self._done[:n].view(torch.bool))
        a.d_actions = ptr
        if torch.cuda.current_device() == self.device.index:      # skip the device-guard round trip (host-bound small batches)
            rc = self._lib.ts_step(C.byref(a), torch.cuda.current_stream().cuda_stream)
        else:
            with torch.cuda.device(self.device):
                rc = self._lib.ts_step(C.byref(a), self._stream())
        if rc:
            check(rc, "ts_step")
        if self.host_io:          # the outputs are host memory: they are valid once the stream has drained
            torch.cuda.current_stream(self.device).synchronize()
        return self._cached_out

    def step_host_io(self, action: int, obs_out: torch.Tensor | None = None) -> None:
        """host_io batches: every env takes `action`; K2 (and K3 into the pinned `obs_out`) are launched
        and the stream is synchronised -- afterwards pos / reward / done / flags (and obs_out) can be
        read on the host as they are.  No tensor is created on this path."""
        if not self.host_io:
            raise RuntimeError("step_host_io needs a batch built with host_io=True")
        self._require_loaded()
        if self._io_np is None:
            self._io_np = self._actions.numpy()
        self._io_np[: self.n_envs] = action
        a = self._cached_args
        if a is None:
            a = self._cached_args = self._step_args(self._actions.data_ptr())
            n = self.n_envs
            self._cached_out = (self._pos[:n], self._reward[:n], self._done[:n].view(torch.bool))
        a.d_actions = self._actions.data_ptr()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream()
            rc = self._lib.ts_step(C.byref(a), stream.cuda_stream)
            if rc:
                check(rc, "ts_step")
            if obs_out is not None:
                o = self._cached_obs_args
                if o is None or o.d_obs != obs_out.data_ptr():
                    o = self._cached_obs_args = self._observe_args(obs_out)
                rc = self._lib.ts_observe(C.byref(o), stream.cuda_stream)
                if rc:
                    check(rc, "ts_observe")
            stream.synchronize()

    def capture_steps(self, action_rows: torch.Tensor) -> "torch.cuda.CUDAGraph":
        """Capture `len(action_rows)` consecutive steps (row k = the actions of step k, uint8
        [R, capacity] on this device) into one CUDA graph; `graph.replay()` then advances the
        batch R steps with a single host call.  For launch-bound batches (about 1M envs and
        below) the per-step host cost of Python + ctypes otherwise exceeds the kernel time."""
        self._require_loaded()
        if action_rows.dtype != torch.uint8 or action_rows.device != self.device or action_rows.dim() != 2 \
                or action_rows.shape[1] < self.capacity or not action_rows.is_contiguous():
            raise ValueError("action_rows must be a contiguous uint8 [R, >= capacity] tensor on the env's device")
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self.step(action_rows[0])                      # warm-up outside capture (lazy init)
            torch.cuda.current_stream(self.device).wait_stream(side)
            self.reset()
            with torch.cuda.graph(graph):
                for k in range(action_rows.shape[0]):
                    self.step(action_rows[k])
        return graph

    def raw_move(self, actions) -> torch.Tensor:
        """GameState.move (state.py:120-170) without episode bookkeeping: slides the tiles and
        returns the flags (WON / INVALID bits); step_count, done state and `flags` of the
        episode are untouched."""
        self._require_loaded()
        if self._scratch_count is None:
            self._scratch_count = torch.zeros(self.capacity, dtype=torch.int32, device=self.device)
            self._scratch_flags = torch.zeros(self.capacity, dtype=torch.uint8, device=self.device)
        self._scratch_count.zero_()
        self._scratch_flags.zero_()
        a = self._step_args(self._stage_actions(actions), count=self._scratch_count, flags=self._scratch_flags,
                            max_steps=2 ** 31 - 1, count_bytes=4, auto_reset=False, raw=True)
        if self._scratch_reward is None:
            self._scratch_reward = torch.zeros(self.capacity, dtype=torch.float32, device=self.device)
        a.d_reward = self._scratch_reward.data_ptr()          # the episode's reward buffer stays untouched
        with torch.cuda.device(self.device):
            check(self._lib.ts_step(C.byref(a), self._stream()), "ts_step")
        return self._scratch_flags[: self.n_envs]

    # ------------------------------------------------------------------ host-buffer path
    def step_host(self, h_actions: torch.Tensor, h_reward: torch.Tensor | None = None, h_done: torch.Tensor | None = None,
                  h_flags: torch.Tensor | None = None, chunk_envs: int | None = None, n_streams: int = 4) -> None:
        """The same step driven from pinned HOST buffers (ts_step_host): uploads the actions,
        runs K2 and downloads reward + done (and / or the status byte `flags`), pipelined in
        chunks; returns when everything has landed.  With only `h_flags` given, 1 byte per env
        comes back instead of 5: bit 0 is done, and the reward follows from the WON / INVALID
        bits (`rewards`)."""
        self._require_loaded()
        if h_flags is None and (h_reward is None or h_done is None):
            raise ValueError("step_host needs h_reward and h_done, or h_flags")
        if h_flags is not None and not self.track_flags:
            raise ValueError("this batch was built with track_flags=False: there is no status byte to download")
        for t, dt in ((h_actions, torch.uint8), (h_reward, torch.float32), (h_done, torch.uint8), (h_flags, torch.uint8)):
            if t is not None and (t.is_cuda or t.dtype != dt or t.numel() < self.n_envs or not t.is_contiguous()):
                raise ValueError("step_host needs contiguous host tensors: uint8 actions, float32 reward, uint8 done / flags")
        if chunk_envs is None:      # 2M-env chunks keep PCIe busy (profiles/r1_pcie_e2e.json); small batches: one chunk per stream
            chunk_envs = min(1 << 21, max(1 << 16, -(-self.n_envs // n_streams)))
        if self._host_ctx is None:
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                check(self._lib.ts_host_ctx_create(C.byref(h), n_streams), "ts_host_ctx_create")
            self._host_ctx = h
        a = self._step_args(self._actions.data_ptr())
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()
            check(self._lib.ts_step_host(self._host_ctx, C.byref(a), h_actions.data_ptr(), _ptr(h_reward), _ptr(h_done),
                                         _ptr(h_flags), _round_up(chunk_envs, CAP_ALIGN)), "ts_step_host")

    def __del__(self):
        h, self._host_ctx = getattr(self, "_host_ctx", None), None
        if h is not None:
            try:
                self._lib.ts_host_ctx_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ views and queries
    def _require_loaded(self):
        if not self._loaded:
            raise RuntimeError("no puzzles loaded: use from_puzzles / from_arrays / synthetic / encode")

    @property
    def pos(self) -> torch.Tensor:
        return self._pos[: self.n_envs]

    @property
    def flags(self) -> torch.Tensor:
        """uint8[N]: F_DONE | F_WON | F_INVALID | F_TIMEOUT | F_STALE of the last step."""
        if not self.track_flags:
            raise RuntimeError("this batch was built with track_flags=False: the step stores reward and done only")
        return self._flags[: self.n_envs]

    @property
    def done(self) -> torch.Tensor:
        return self._done[: self.n_envs].view(torch.bool)

    @property
    def reward(self) -> torch.Tensor:
        return self._reward[: self.n_envs]

    @property
    def step_count(self) -> torch.Tensor:
        return self._count[: self.n_envs]

    @property
    def terminal_pos(self) -> torch.Tensor | None:
        return None if self._terminal is None else self._terminal[: self.n_envs]

    def is_won(self) -> torch.Tensor:
        """bool[N]: WON bit of the last step (the reference evaluates the goal only after a
        move, environment.py:123; nothing is won straight after reset)."""
        if not self.track_flags:      # reward = r_win exactly when the step won
            if self.rewards[0] in self.rewards[1:]:
                raise RuntimeError("track_flags=False and r_win equals another reward: is_won() cannot be told from the reward")
            return self.reward == self.rewards[0]
        return (self.flags & F_WON) != 0

    def goal_check(self) -> torch.Tensor:
        """bool[N]: GameState.is_won (state.py:172-186) of the CURRENT positions, without moving."""
        self._require_loaded()
        won = torch.zeros(self.capacity, dtype=torch.uint8, device=self.device)
        a = GoalArgs(size=self.size, n_tiles=self.n_tiles, goal_mode=self.goal_mode, never_win=int(self.never_win),
                     first_env=0, n_envs=self.n_envs, capacity=self.capacity,
                     d_targets_packed=_ptr(self._targets), d_pos=_ptr(self._pos), d_won=_ptr(won))
        with torch.cuda.device(self.device):
            check(self._lib.ts_goal_check(C.byref(a), self._stream()), "ts_goal_check")
        return won[: self.n_envs].view(torch.bool)

    def positions(self, packed: torch.Tensor | None = None) -> torch.Tensor:
        """Decode packed position words to uint8[N, T, 2] (row, col)."""
        p = self.pos if packed is None else packed
        p = p[:, : self.n_tiles]
        ps = self.pos_stride
        return torch.stack((p // ps, p % ps), dim=-1)

    def _board_cells(self, buf: torch.Tensor, walls: bool = False) -> torch.Tensor:
        """Unpack a plane-layout bitboard buffer to bool[N, S*S] (load-time / debugging aid)."""
        nb, cap, n = self.board_bytes, self.capacity, self.n_envs
        if self.wide:   # u16 lines: walls [axis][capacity][S rounded up to even] (last plane = rows), targets [capacity][16]
            per_env = (self.size + 1) // 2 * 2 if walls else 17    # targets: 16 rows + the distinct-target count
            lines = buf.view(torch.int16).view(-1, cap, per_env)[-1, :n, :16 if not walls else per_env].to(torch.int32) & 0xFFFF
            lead = 1 if walls and self.size <= 14 else 0     # wall lines of S <= 14 start with an edge sentinel
            bits = (lines.unsqueeze(-1) >> (torch.arange(16, device=buf.device) + lead)) & 1
            return bits[:, : self.size, : self.size].reshape(n, self.size * self.size).bool()
        cols = []
        for k in range(self._lib.ts_plane_count(nb)):
            w, off = self._lib.ts_plane_width(nb, k), self._lib.ts_plane_offset(nb, k)
            cols.append(buf[off * cap: (off + w) * cap].view(cap, w)[:n])
        by = torch.cat(cols, dim=1)                                          # [N, nb] bytes, little endian
        bits = (by.unsqueeze(-1) >> torch.arange(8, device=by.device, dtype=torch.uint8)) & 1
        S, bs = self.size, self.board_stride
        return bits.reshape(n, nb * 8)[:, : S * bs].reshape(n, S, bs)[:, :, :S].reshape(n, S * S).bool()

    def blocked_cells(self) -> torch.Tensor:
        return self._board_cells(self._walls, walls=True)

    def target_positions(self) -> torch.Tensor:
        """Ordered mode: uint8[N,T,2].  Set mode: bool[N,S*S] target cells."""
        if self.goal_mode == GOAL_ORDERED:
            if self._obs_targets is not None:       # target count != tile count (never_win)
                t = self._obs_targets[: self.n_envs, : self.n_targets]
            else:
                t = self._targets.view(self.capacity, self.pos_bytes)[: self.n_envs, : self.n_tiles]
            return torch.stack((t // self.pos_stride, t % self.pos_stride), dim=-1)
        return self._board_cells(self._targets)

    def observe(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """K3: float32[N,S,S,3] observation, GameState.get_state_array (state.py:188-211)."""
        self._require_loaded()
        S, n = self.size, self.n_envs
        if out is None:
            out = torch.empty(n, S, S, 3, dtype=torch.float32, device=self.device)
        # single colour with one tile is stored as an ordered batch: index+1 == 1, same values
        a = self._observe_args(out)
        with torch.cuda.device(self.device):
            check(self._lib.ts_observe(C.byref(a), self._stream()), "ts_observe")
        return out

    def _observe_args(self, out: torch.Tensor) -> ObserveArgs:
        # a never_win batch draws its ordered targets from their own buffer (count != tile count)
        nt = 0 if self._obs_targets is None else (self.n_targets or -1)
        tg = self._targets if self._obs_targets is None else self._obs_targets
        return ObserveArgs(size=self.size, n_tiles=self.n_tiles, goal_mode=self.goal_mode, n_targets=nt,
                           first_env=0, n_envs=self.n_envs, capacity=self.capacity, d_walls=_ptr(self._walls),
                           d_targets_packed=_ptr(tg), d_pos=_ptr(self._pos), d_obs=_ptr(out))

    def valid_moves(self) -> torch.Tensor:
        """uint8[N]: bit d set when move d changes the state (get_valid_moves,
        environment.py:149-171)."""
        self._require_loaded()
        mask = torch.zeros(self.capacity, dtype=torch.uint8, device=self.device)
        a = ValidArgs(size=self.size, n_tiles=self.n_tiles, first_env=0, n_envs=self.n_envs, capacity=self.capacity,
                      d_walls=_ptr(self._walls), d_pos=_ptr(self._pos), d_mask=_ptr(mask))
        with torch.cuda.device(self.device):
            check(self._lib.ts_valid_moves(C.byref(a), self._stream()), "ts_valid_moves")
        return mask[: self.n_envs]

    def set_positions(self, tiles) -> None:
        """Overwrite current positions from uint8[N,T,2] (row, col) -- e.g. to expand a search
        frontier; validity is the caller's responsibility."""
        t = torch.as_tensor(tiles, dtype=torch.uint8).to(self.device)
        packed = t[..., 0] * self.pos_stride + t[..., 1]
        self._pos[: self.n_envs, : self.n_tiles] = packed

    def puzzle(self, i: int) -> Puzzle:
        """Decode env i back to a Puzzle (host copy)."""
        S = self.size
        blocked = [(int(c) // S, int(c) % S) for c in torch.nonzero(self.blocked_cells()[i]).flatten().tolist()]
        init = self._init[i, : self.n_tiles].cpu()
        tiles = [divmod(int(b), self.pos_stride) for b in init.tolist()]
        if self.goal_mode == GOAL_ORDERED:
            tg = [(int(r), int(c)) for r, c in self.target_positions()[i].cpu().tolist()]
        else:
            tg = [(int(c) // S, int(c) % S) for c in torch.nonzero(self.target_positions()[i]).flatten().tolist()]
        return Puzzle(S, blocked, tiles, tg, self.multi_color)


def shard_range(n_total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous env-index shard [lo, hi) of rank `rank`: envs never interact,
    so the step path needs no collective; the remainder goes to the first ranks."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
