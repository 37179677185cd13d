"""Pure-Python/numpy restatement of the reference's step loop.  TEST INFRASTRUCTURE ONLY.

Purpose: (1) a second, independent oracle for small cases, cross-checked against
ts_oracle.c and the reference-generated fixtures in tests/golden/; (2) the `cpu_baseline`
/ `--impl reference` leg of bench.py -- it keeps the reference's own data structures
(numpy int slide table, np.argsort on a Python list, tuple positions, a Python set for the
cells taken this move, float32 HWC observation built per step), so that its speed on the
GPU box's host cores is representative of the reference's Python loop, which cannot travel
to that box.  Only tests/, __graft_entry__.smoke() and bench.py may import it.

Reference map (paths relative to the reference checkout):
  PortState.__init__/_slide_table  explainrl/environment/state.py:47-118
  PortState.move                   explainrl/environment/state.py:120-170
  PortState.is_won                 explainrl/environment/state.py:172-186
  PortState.observation            explainrl/environment/state.py:188-211
  PortEnv.reset / step             explainrl/environment/environment.py:82-143
"""
from __future__ import annotations

import numpy as np

UP, DOWN, LEFT, RIGHT = 0, 1, 2, 3
# (d_row, d_col) of one sliding step per action
_DELTA = {UP: (-1, 0), DOWN: (1, 0), LEFT: (0, -1), RIGHT: (0, 1)}


class PortState:
    def __init__(self, size, blocked, initial, targets, multi_color=False):
        self.size = size
        self.current_locations = list(initial)
        self.target_locations = list(targets)
        self.multi_color = multi_color
        self.is_blocked = np.zeros((size, size), dtype=bool)
        for r, c in blocked:
            self.is_blocked[r, c] = True
        self._slide_table()

    def _slide_table(self):
        """state.py:75-118: each cell inherits the destination of its neighbour in the
        move direction when that neighbour is in bounds and not blocked."""
        n = self.size
        self.move_to = np.full((n, n, 4, 2), -1, dtype=int)
        for d, (dr, dc) in _DELTA.items():
            rows = range(n) if dr <= 0 else reversed(range(n))
            for r in rows:
                cols = range(n) if dc <= 0 else reversed(range(n))
                for c in cols:
                    nr, nc = r + dr, c + dc
                    if 0 <= nr < n and 0 <= nc < n and not self.is_blocked[nr, nc]:
                        self.move_to[r, c, d] = self.move_to[nr, nc, d]
                    else:
                        self.move_to[r, c, d] = (r, c)

    def move(self, d):
        """state.py:137-170: nearest-to-the-wall first, back off while the cell is taken."""
        locs = self.current_locations
        if d == UP:
            order = np.argsort([r for r, _ in locs])
        elif d == DOWN:
            order = np.argsort([-r for r, _ in locs])
        elif d == LEFT:
            order = np.argsort([c for _, c in locs])
        else:
            order = np.argsort([-c for _, c in locs])
        dr, dc = _DELTA[d]
        taken = set()
        for i in order:
            here = tuple(self.move_to[locs[i][0], locs[i][1], d])
            while here in taken:
                here = (here[0] - dr, here[1] - dc)
            locs[i] = here
            taken.add(here)
        return self.is_won()

    def is_won(self):
        if self.multi_color:
            return self.current_locations == self.target_locations
        return set(self.current_locations) == set(self.target_locations)

    def observation(self):
        obs = np.zeros((self.size, self.size, 3), dtype=np.float32)
        obs[:, :, 0] = self.is_blocked.astype(np.float32)
        for k, (r, c) in enumerate(self.current_locations):
            obs[r, c, 1] = k + 1 if self.multi_color else 1
        for k, (r, c) in enumerate(self.target_locations):
            obs[r, c, 2] = k + 1 if self.multi_color else 1
        return obs


class PortEnv:
    """environment.py:33-143 without the enum type check (actions are ints 0..3 here)."""

    def __init__(self, size, blocked, initial, targets, multi_color=False, max_steps=100):
        self.size, self.blocked, self.initial, self.targets = size, blocked, initial, targets
        self.multi_color, self.max_steps = multi_color, max_steps
        self.state, self.step_count, self.done = None, 0, False

    def reset(self):
        self.state = PortState(self.size, self.blocked, self.initial, self.targets, self.multi_color)
        self.step_count, self.done = 0, False
        return self.state.observation()

    def step(self, d):
        if self.done:
            raise RuntimeError("Episode is done. Call reset() to start a new episode.")
        before = self.state.current_locations.copy()
        won = self.state.move(d)
        info = {"is_won": won, "step_count": self.step_count,
                "invalid_move": before == self.state.current_locations}
        if won:
            self.done = True
            info["success"] = True
        self.step_count += 1
        if self.step_count >= self.max_steps:
            self.done = True
            info["timeout"] = True
        return self.state.observation(), self.done, info


def run_loop(puzzles, actions, max_steps=100):
    """The CPU baseline loop of SURVEY 8(d): `obs,done,info = env.step(a); if done: env.reset()`.

    puzzles: list of (size, blocked, tiles, targets, multi_color); actions: int array [K, n].
    Returns (env_steps_done, final positions per env, number of wins)."""
    envs = [PortEnv(s, b, i, t, m, max_steps) for (s, b, i, t, m) in puzzles]
    for e in envs:
        e.reset()
    wins = 0
    K = len(actions)
    for k in range(K):
        row = actions[k]
        for j, e in enumerate(envs):
            _, done, info = e.step(int(row[j]))
            if done:
                wins += bool(info["is_won"])
                e.reset()
    return K * len(envs), [list(e.state.current_locations) for e in envs], wins
