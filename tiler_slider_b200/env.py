"""Single-environment drop-in surface over the CUDA path (batch size 1).

Mirrors, name for name, the reference's
  GameState               explainrl/environment/state.py:18-222
  TilerSliderEnv          explainrl/environment/environment.py:14-194
  TilerSliderEnvFactory   explainrl/environment/environment.py:197-288
so that code written against the reference (its TextRender, its tests) runs unchanged while
every move, goal check, observation and valid-move probe is computed by the sm_100a kernels
through the C-ABI.  Only the Python exceptions of step() (RuntimeError after done,
TypeError on a non-enum action; environment.py:113-117) and the info dict are assembled on
the host, from the flag byte the step kernel returns.

Limits (ValueError): board size <= 16, 0..32 tiles, well-formed puzzles only (distinct tiles,
none on a blocked cell; outside that domain the reference itself is erratic).
Boards without tiles (won iff there are no targets either) and multi-colour boards whose target
count differs from their tile count (they play normally and never win, state.py:183-184) behave
as in the reference.  Nothing is ever computed on the host: there is no CPU fallback behind
these classes.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from ._lib import F_INVALID, F_TIMEOUT, F_WON
from .batch_env import BatchedTilerSliderEnv
from .moves import Move
from .puzzle import Puzzle, parse_board_text


class GameState:
    """One board on the GPU.  Same constructor, attributes and methods as the reference's
    GameState (state.py:47-222)."""

    Move = Move

    def __init__(self, size: int, blocked_locations, initial_locations, target_locations,
                 multi_color: bool = False, *, device: str | torch.device = "cuda", _max_steps: int = 100):
        self.size = size
        self.target_locations = copy.copy(target_locations)
        self.multi_color = multi_color
        self._blocked = [(int(r), int(c)) for r, c in blocked_locations]
        self._device = device
        self._max_steps = _max_steps
        self.is_blocked = np.zeros((size, size), dtype=bool)
        for i, j in self._blocked:
            self.is_blocked[i, j] = True
        self._move_to = None
        self._locs = copy.copy(initial_locations)
        self._dirty = False
        self._initial = copy.copy(initial_locations)
        self._batch_obj: Optional[BatchedTilerSliderEnv] = None

    @property
    def _batch(self) -> BatchedTilerSliderEnv:
        """The batch-of-one CUDA environment, built on first use: constructing a GameState only
        records the board (like the reference's constructor, it accepts any size and tile count);
        shapes the kernels do not cover raise ValueError when the first move / goal check /
        observation is asked for."""
        if self._batch_obj is None:
            p = Puzzle(self.size, self._blocked, [(int(r), int(c)) for r, c in self._initial],
                       [(int(r), int(c)) for r, c in self.target_locations], bool(self.multi_color))
            # host_io: positions and step outputs in pinned host memory the kernels write directly -- a
            # step of this batch of one is two launches and a stream synchronisation, no copies
            b = self._batch_obj = BatchedTilerSliderEnv.from_puzzles([p], max_steps=self._max_steps, auto_reset=False,
                                                                     device=self._device, host_io=True)
            self._np_pos, self._np_flags, self._np_reward = b._pos.numpy(), b._flags.numpy(), b._reward.numpy()
            self._obs_pinned = torch.zeros(1, self.size, self.size, 3, dtype=torch.float32, pin_memory=True)
            self._np_obs = self._obs_pinned.numpy()
            self._obs_fresh = False
            if self._locs != self._initial:      # positions assigned before the first use
                b.set_positions(torch.tensor([[list(map(int, rc)) for rc in self._locs]], dtype=torch.uint8))
        return self._batch_obj

    # -- positions ------------------------------------------------------------------------
    @property
    def current_locations(self):
        """List of (row, col) per tile (the reference's become numpy ints after a move;
        compare by value)."""
        if self._dirty:
            b = self._batch
            torch.cuda.current_stream(b.device).synchronize()      # the kernels write positions straight into host memory
            ps = b.pos_stride
            self._locs = [(int(v) // ps, int(v) % ps) for v in self._np_pos[0, : b.n_tiles]]
            self._dirty = False
        return self._locs

    @current_locations.setter
    def current_locations(self, locs):
        self._locs = list(locs)
        self._dirty = False
        if self._batch_obj is not None:
            self._obs_fresh = False
            self._batch_obj.set_positions(torch.tensor([[list(map(int, rc)) for rc in locs]], dtype=torch.uint8))

    # -- the move path ----------------------------------------------------------------------
    def move(self, move: Move) -> bool:
        """state.py:120-170 on the GPU; returns is_won()."""
        flags = self._batch.raw_move(torch.tensor([move.value], dtype=torch.uint8))
        self._dirty, self._obs_fresh = True, False
        return bool(int(flags[0]) & F_WON)

    def _env_step(self, move: Move) -> int:
        """One bookkept step (K2 with the env's max_steps) plus the observation of the new state (K3),
        both written by the kernels into pinned host memory; returns the flag byte."""
        self._batch.step_host_io(move.value, self._obs_pinned)
        self._dirty, self._obs_fresh = True, True
        return int(self._np_flags[0])

    def is_won(self) -> bool:
        """state.py:172-186 on the current positions (ts_goal_check)."""
        return bool(self._batch.goal_check()[0])

    def get_state_array(self) -> np.ndarray:
        """state.py:188-211: float32[S,S,3] from K3."""
        b = self._batch
        if not self._obs_fresh:
            b.observe(self._obs_pinned)
            torch.cuda.current_stream(b.device).synchronize()
            self._obs_fresh = True
        return self._np_obs[0].copy()

    @property
    def move_to(self) -> np.ndarray:
        """Slide table int[S,S,4,2] (state.py:75-118), produced by the step kernel itself: a
        batch of S*S single-tile boards, one per start cell, moved once per direction."""
        if self._move_to is None:
            S = self.size
            n = S * S
            blocked = np.tile(self.is_blocked.reshape(1, n).astype(np.uint8), (n, 1))
            cells = np.arange(n)
            tiles = np.stack([cells // S, cells % S], axis=-1).astype(np.uint8).reshape(n, 1, 2)
            # the table is defined for blocked start cells too: the reference's sweeps never look at
            # the start cell itself (state.py:85-118), so probe i runs on a board whose cell i is open
            blocked[cells, cells] = 0
            probe = BatchedTilerSliderEnv.from_arrays(S, blocked, tiles, tiles.copy(), False, device=self._device)
            table = np.zeros((S, S, 4, 2), dtype=int)
            for d in range(4):
                probe.reset()
                probe.raw_move(torch.full((n,), d, dtype=torch.uint8))
                table[:, :, d, :] = probe.positions()[:, 0].cpu().numpy().reshape(S, S, 2)
            self._move_to = table
        return self._move_to

    def valid_moves(self) -> list[Move]:
        mask = int(self._batch.valid_moves()[0])
        return [m for m in Move if mask >> m.value & 1]

    def copy(self) -> "GameState":
        """state.py:213-222."""
        return GameState(self.size, list(self._blocked), copy.copy(self.current_locations),
                         copy.copy(self.target_locations), self.multi_color, device=self._device,
                         _max_steps=self._max_steps)


class TilerSliderEnv:
    """Drop-in for the reference's TilerSliderEnv (environment.py:14-194)."""

    def __init__(self, size: int = None, blocked_locations: list = None, initial_locations: list = None,
                 target_locations: list = None, multi_color: bool = False, max_steps: int = 100,
                 device: str | torch.device = "cuda"):
        self.size = size
        self.blocked_locations = blocked_locations or []
        self.initial_locations = initial_locations or []
        self.target_locations = target_locations or []
        self.multi_color = multi_color
        self.max_steps = max_steps
        self.device = device
        self.state: Optional[GameState] = None
        self.step_count = 0
        self.done = False
        self.observation_shape = (size, size, 3) if size else None

    @classmethod
    def from_level(cls, level, max_steps: int = 100, **kw):
        """environment.py:61-80; `level` has the ImageProcessed fields (dataloader.py:21-27)."""
        return cls(size=level.size, blocked_locations=level.blocked_locations,
                   initial_locations=level.initial_locations, target_locations=level.target_locations,
                   multi_color=level.multiple_colors, max_steps=max_steps, **kw)

    def reset(self) -> np.ndarray:
        """environment.py:82-98."""
        self.state = GameState(self.size, self.blocked_locations, self.initial_locations, self.target_locations,
                               self.multi_color, device=self.device, _max_steps=self.max_steps)
        self.step_count = 0
        self.done = False
        return self.state.get_state_array()

    def step(self, move: Move) -> Tuple[np.ndarray, bool, Dict[str, Any]]:
        """environment.py:100-143: returns (observation, done, info) -- no reward in the
        reference's tuple; the kernel's reward is exposed as `last_reward`."""
        if self.done:
            raise RuntimeError("Episode is done. Call reset() to start a new episode.")
        if not isinstance(move, Move):
            raise TypeError(f"Action must be a GameState.Move enum, got {type(move)}")
        flags = self.state._env_step(move)
        won, invalid, timeout = bool(flags & F_WON), bool(flags & F_INVALID), bool(flags & F_TIMEOUT)
        self.last_reward = float(self.state._np_reward[0])
        info = {"is_won": won, "step_count": self.step_count, "invalid_move": invalid}
        if won:
            self.done = True
            info["success"] = True
        self.step_count += 1
        if timeout:
            self.done = True
            info["timeout"] = True
        return self.state.get_state_array(), self.done, info

    def close(self):
        self.state = None

    def get_valid_moves(self) -> list[Move]:
        """environment.py:149-171, computed by ts_valid_moves."""
        if self.state is None:
            return []
        return self.state.valid_moves()

    def get_info(self) -> Dict[str, Any]:
        """environment.py:173-194."""
        if self.state is None:
            return {"initialized": False}
        return {"initialized": True, "size": self.size, "step_count": self.step_count,
                "max_steps": self.max_steps, "done": self.done, "is_won": self.state.is_won(),
                "num_tiles": len(self.state.current_locations), "num_targets": len(self.state.target_locations),
                "multi_color": self.multi_color, "valid_moves": self.get_valid_moves()}


class TilerSliderEnvFactory:
    """environment.py:197-288."""

    @staticmethod
    def create_simple_env(size: int = 5, num_tiles: int = 2, num_obstacles: int = 3, seed: int = None,
                          **kw) -> TilerSliderEnv:
        """Seeded random puzzle, same draw as the reference (environment.py:217-226): the global
        legacy numpy RNG is reseeded and shuffles the row-major cell list; first
        `num_obstacles` blocked, next `num_tiles` tiles, next `num_tiles` targets."""
        if seed is not None:
            np.random.seed(seed)
        cells = [(i, j) for i in range(size) for j in range(size)]
        np.random.shuffle(cells)
        return TilerSliderEnv(size=size, blocked_locations=cells[:num_obstacles],
                              initial_locations=cells[num_obstacles:num_obstacles + num_tiles],
                              target_locations=cells[num_obstacles + num_tiles:num_obstacles + 2 * num_tiles],
                              multi_color=False, **kw)

    @staticmethod
    def create_from_string(board_str: str, multi_color: bool = False, **kw) -> TilerSliderEnv:
        """Text grid -> env (environment.py:236-288); grammar in puzzle.parse_board_text."""
        p = parse_board_text(board_str, multi_color)
        return TilerSliderEnv(size=p.size, blocked_locations=p.blocked_locations,
                              initial_locations=p.initial_locations, target_locations=p.target_locations,
                              multi_color=multi_color, **kw)
