// ts_generic.cu -- the step / goal kernels for the corners the specialised kernels do not cover:
//   (a) the ragged end of a range on bitboard-class boards (S <= 8): step_kernel works on whole
//       4-env groups, so ts_step hands the last n_envs % 4 envs of a call to generic_step_kernel --
//       nothing outside [first_env, first_env + n_envs) is ever read-modified-written;
//   (b) boards without tiles (T = 0, any S): the reference accepts them -- nothing moves, every
//       move is invalid, and the board is won iff it has no targets either
//       (explainrl/environment/state.py:183-186, tests/test_state.py:40-52,
//       tests/test_environment.py:569-580 of the reference).
// Thread = one env, runtime S and T, plain loops: this is the closed form of GameState.move
// (state.py:120-170; ts_common.cuh) written out cell by cell -- a tile advances by the number of
// EMPTY cells between itself and the first wall / edge ahead -- and the bookkeeping of
// TilerSliderEnv.step (explainrl/environment/environment.py:119-143) and reset (:89-97) exactly as
// step_group (ts_step.cuh) does it.  Not a fast path: at most 3 envs per call in case (a).
#include "ts_common.cuh"
#include "../../include/tiler_slider.h"

namespace ts {

__global__ void __launch_bounds__(128) generic_step_kernel(const ts_step_args a) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= a.n_envs) return;
    const size_t env = (size_t)(a.first_env + i), cap = (size_t)a.capacity;
    const int S = a.size, T = a.n_tiles, pw = pos_bytes(T), ps = pos_stride(S), bs = board_stride(S), nb = board_bytes(S);
    const uint32_t action = a.d_actions[env] & 3u;
    const int dr = action == 0 ? -1 : action == 1 ? 1 : 0, dc = action == 2 ? -1 : action == 3 ? 1 : 0;   // state.py:31-34

    uint8_t p0[MAX_TILES], p[MAX_TILES];
    for (int t = 0; t < T; ++t) p[t] = p0[t] = a.d_pos[env * pw + t];
    bool moved = false;
    if (T > 0) {                                   // bitboard classes only (ts_step never sends wide boards with tiles here)
        uint64_t walls = 0, occ = 0;               // the whole board in one word (independent byte loads), then register bit tests
        for (int b = 0; b < nb; ++b) walls |= (uint64_t)a.d_walls[board_byte_addr(nb, cap, env, b)] << (8 * b);
        for (int t = 0; t < T; ++t) occ |= 1ull << ((p0[t] / ps) * bs + p0[t] % ps);
        for (int t = 0; t < T; ++t) {
            const int r = p0[t] / ps, c = p0[t] % ps;
            int n = 0;
            for (int rr = r + dr, cc = c + dc; rr >= 0 && rr < S && cc >= 0 && cc < S; rr += dr, cc += dc) {
                const int bit = rr * bs + cc;
                if ((walls >> bit) & 1ull) break;
                if (!((occ >> bit) & 1ull)) ++n;
            }
            p[t] = (uint8_t)((r + n * dr) * ps + (c + n * dc));
            moved |= n != 0;
        }
    }
    // goal (state.py:183-186)
    bool won = a.never_win == 0;
    if (a.goal_mode == TS_GOAL_ORDERED) {
        for (int t = 0; t < T; ++t) won &= p[t] == a.d_targets_packed[env * pw + t];
    } else if (wide_board(S)) {                    // T == 0 here: won iff the board has no target cell
        won &= reinterpret_cast<const uint16_t*>(a.d_targets_packed)[env * WIDE_TARGET_WORDS + 16] == (uint16_t)T;
    } else {
        uint64_t occ = 0, tb = 0;
        for (int t = 0; t < T; ++t) occ |= 1ull << ((p[t] / ps) * bs + p[t] % ps);
        for (int b = 0; b < nb; ++b) tb |= (uint64_t)a.d_targets_packed[board_byte_addr(nb, cap, env, b)] << (8 * b);
        won &= occ == tb;
    }
    // bookkeeping (environment.py:126-141)
    const bool narrow = a.count_bytes == 1;
    uint32_t count = narrow ? (uint32_t)reinterpret_cast<const uint8_t*>(a.d_step_count)[env]
                            : reinterpret_cast<const uint32_t*>(a.d_step_count)[env];
    count += 1u;
    const bool timeout = (int)count >= a.max_steps;
    bool done = won || timeout;
    uint32_t flags = (done ? F_DONE : 0u) | (won ? F_WON : 0u) | (moved ? 0u : F_INVALID) | (timeout ? F_TIMEOUT : 0u);
    float reward = won ? a.r_win : (moved ? a.r_step : a.r_invalid);
    if (!a.auto_reset && (a.d_flags[env] & F_DONE)) {   // frozen: the reference raises here (environment.py:113-114)
        flags = F_DONE | F_STALE;
        reward = 0.0f;
        count -= 1u;
        done = true;
        for (int t = 0; t < T; ++t) p[t] = p0[t];
    }
    if (done) {
        if (a.d_terminal_pos) for (int t = 0; t < pw; ++t) a.d_terminal_pos[env * pw + t] = t < T ? p[t] : 0;
        if (a.auto_reset) {                              // environment.py:89-97
            for (int t = 0; t < T; ++t) p[t] = a.d_init[env * pw + t];
            count = 0;
        }
    }
    for (int t = 0; t < pw; ++t) a.d_pos[env * pw + t] = t < T ? p[t] : 0;
    if (narrow) reinterpret_cast<uint8_t*>(a.d_step_count)[env] = (uint8_t)count;
    else reinterpret_cast<uint32_t*>(a.d_step_count)[env] = count;
    a.d_reward[env] = reward;
    if (a.d_done) a.d_done[env] = done ? 1 : 0;
    if (a.d_flags) a.d_flags[env] = (uint8_t)flags;
}

// GameState.is_won of a board without tiles: won iff it has no targets (and the caller's never_win,
// which carries "ordered mode with a non-empty target list", is clear)
__global__ void __launch_bounds__(128) empty_goal_kernel(const ts_goal_args a) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= a.n_envs) return;
    const size_t env = (size_t)(a.first_env + i), cap = (size_t)a.capacity;
    bool won = a.never_win == 0;
    if (a.goal_mode == TS_GOAL_SET) {
        if (wide_board(a.size)) {
            won &= reinterpret_cast<const uint16_t*>(a.d_targets_packed)[env * WIDE_TARGET_WORDS + 16] == 0;
        } else {
            const int nb = board_bytes(a.size);
            for (int b = 0; b < nb; ++b) won &= a.d_targets_packed[board_byte_addr(nb, cap, env, b)] == 0;
        }
    }
    a.d_won[env] = won ? 1 : 0;
}

cudaError_t generic_step_dispatch(const ts_step_args& a, cudaStream_t st) {
    if (a.n_envs <= 0) return cudaSuccess;
    generic_step_kernel<<<(unsigned)((a.n_envs + 127) / 128), 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t empty_goal_dispatch(const ts_goal_args& a, cudaStream_t st) {
    if (a.n_envs <= 0) return cudaSuccess;
    empty_goal_kernel<<<(unsigned)((a.n_envs + 127) / 128), 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ts
