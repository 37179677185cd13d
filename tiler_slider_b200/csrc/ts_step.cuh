// ts_step.cuh -- K2, the fused step kernel (bitboard variant, S*S <= 64).
//
// One launch advances every env of the range by one externally supplied action and fuses
//   GameState.move            explainrl/environment/state.py:120-170   (slide_env)
//   GameState.is_won          explainrl/environment/state.py:172-186   (goal compare)
//   TilerSliderEnv.step       explainrl/environment/environment.py:119-143
//                             invalid_move, done, step counter, timeout
//   TilerSliderEnv.reset      explainrl/environment/environment.py:89-97 (optional auto-reset)
// plus the repo-defined reward.  Thread = 4 consecutive envs; all traffic is 32/64/128-bit
// coalesced, algorithmic bytes per env-step are 3T + ceil(S^2/8) + 8.
#pragma once
#include "ts_common.cuh"
#include "../../include/tiler_slider.h"

namespace ts {

constexpr int STEP_THREADS = 256;

template <int S, int T, int GOAL>
__global__ void __launch_bounds__(STEP_THREADS) step_kernel(const __grid_constant__ ts_step_args a) {
    using BT = BoardTraits<S>;
    using board_t = typename BT::board_t;
    constexpr int PW = pos_bytes(T), PR = (T + 3) / 4, NB = board_bytes(S);
    constexpr int NWORDS = (NB + 3) / 4;

    const size_t n_groups = (size_t)((a.n_envs + GROUP - 1) / GROUP);
    size_t g = (size_t)blockIdx.x * STEP_THREADS + threadIdx.x;
    if (g >= n_groups) return;
    g += (size_t)a.first_env / GROUP;
    const size_t cap = (size_t)a.capacity;
    const size_t e0 = g * GROUP;

    // ---- loads (all issued before first use) ------------------------------------------------
    uint32_t praw[PW];
    ld_words<PW>(a.d_pos + e0 * PW, praw);
    BoardGroup<NB> walls;
    walls.load(a.d_walls, cap, g);
    uint32_t traw[PW];
    BoardGroup<NB> tboard;
    if constexpr (GOAL == TS_GOAL_ORDERED) ld_words<PW>(a.d_targets_packed + e0 * PW, traw);
    else tboard.load(a.d_targets_packed, cap, g);
    const uint32_t act4 = __ldcs(reinterpret_cast<const unsigned int*>(a.d_actions + e0));
    const bool wide_count = a.count_bytes == 4;
    uint32_t cnt[GROUP];
    if (wide_count) {
        const uint4 c = __ldcs(reinterpret_cast<const uint4*>(a.d_step_count) + g);
        cnt[0] = c.x; cnt[1] = c.y; cnt[2] = c.z; cnt[3] = c.w;
    } else {
        const uint32_t c = __ldcs(reinterpret_cast<const unsigned int*>(a.d_step_count) + g);
        cnt[0] = c & 0xFF; cnt[1] = (c >> 8) & 0xFF; cnt[2] = (c >> 16) & 0xFF; cnt[3] = c >> 24;
    }
    uint32_t prev_flags = 0;
    if (!a.auto_reset) prev_flags = __ldcs(reinterpret_cast<const unsigned int*>(a.d_flags + e0));

    uint32_t pnew[PW];
#pragma unroll
    for (int j = 0; j < PW; ++j) pnew[j] = praw[j];
    uint32_t flags4 = 0, done4 = 0;
    float rew[GROUP];

#pragma unroll
    for (int e = 0; e < GROUP; ++e) {
        uint32_t q[PR], q0[PR], bw[NWORDS];
        group_elem<PW>(praw, e, q0);
#pragma unroll
        for (int w = 0; w < PR; ++w) q[w] = q0[w];
        walls.get(e, bw);
        board_t wb;
        if constexpr (BT::WIDE) wb = (uint64_t)bw[0] | ((uint64_t)bw[1] << 32);
        else wb = bw[0];
        const uint32_t action = (act4 >> (8 * e)) & 3u;

        slide_env<S, T>(q, wb, action);

        bool moved = false;
#pragma unroll
        for (int w = 0; w < PR; ++w) moved |= q[w] != q0[w];
        bool won;
        if constexpr (GOAL == TS_GOAL_ORDERED) {
            uint32_t tq[PR];
            group_elem<PW>(traw, e, tq);
            won = true;
#pragma unroll
            for (int w = 0; w < PR; ++w) won &= q[w] == tq[w];
        } else {
            uint32_t tw[NWORDS];
            tboard.get(e, tw);
            board_t tb;
            if constexpr (BT::WIDE) tb = (uint64_t)tw[0] | ((uint64_t)tw[1] << 32);
            else tb = tw[0];
            won = occupancy<S, T>(q) == tb;
        }
        won = won && !a.never_win;
        const uint32_t c1 = cnt[e] + 1u;
        const bool timeout = (int)c1 >= a.max_steps;
        const bool done = won || timeout;
        uint32_t f = (done ? F_DONE : 0u) | (won ? F_WON : 0u) | (moved ? 0u : F_INVALID) | (timeout ? F_TIMEOUT : 0u);
        float r = won ? a.r_win : (moved ? a.r_step : a.r_invalid);
        uint32_t cn = c1;
        const bool stale = ((prev_flags >> (8 * e)) & F_DONE) != 0;   // only ever set when !auto_reset
        if (stale) {
            f = F_DONE | F_STALE;
            r = 0.0f;
            cn = cnt[e];
#pragma unroll
            for (int w = 0; w < PR; ++w) q[w] = q0[w];
        }
        cnt[e] = cn;
        rew[e] = r;
        flags4 |= f << (8 * e);
        done4 |= (f & F_DONE) << (8 * e);
        group_set<PW>(pnew, e, q);
    }

    // ---- auto-reset (environment.py:89-97) and the optional terminal snapshot ----------------
    if (done4 != 0 && a.auto_reset) {
        if (a.d_terminal_pos) st_words<PW>(a.d_terminal_pos + e0 * PW, pnew);
        uint32_t iraw[PW];
        ld_words<PW>(a.d_init + e0 * PW, iraw);
#pragma unroll
        for (int e = 0; e < GROUP; ++e) {
            if ((done4 >> (8 * e)) & 1u) {
                uint32_t q[PR];
                group_elem<PW>(iraw, e, q);
                group_set<PW>(pnew, e, q);
                cnt[e] = 0;
            }
        }
    } else if (done4 != 0 && a.d_terminal_pos) {
        st_words<PW>(a.d_terminal_pos + e0 * PW, pnew);
    }

    // ---- stores --------------------------------------------------------------------------------
    st_words<PW>(a.d_pos + e0 * PW, pnew);
    if (wide_count) __stcs(reinterpret_cast<uint4*>(a.d_step_count) + g, make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]));
    else __stcs(reinterpret_cast<unsigned int*>(a.d_step_count) + g, cnt[0] | (cnt[1] << 8) | (cnt[2] << 16) | (cnt[3] << 24));
    __stcs(reinterpret_cast<float4*>(a.d_reward) + g, make_float4(rew[0], rew[1], rew[2], rew[3]));
    if (a.d_done) __stcs(reinterpret_cast<unsigned int*>(a.d_done + e0), done4);
    if (a.d_flags) __stcs(reinterpret_cast<unsigned int*>(a.d_flags + e0), flags4);
}

template <int S, int T>
inline cudaError_t launch_step(const ts_step_args& a, cudaStream_t stream) {
    const size_t n_groups = (size_t)((a.n_envs + GROUP - 1) / GROUP);
    const unsigned blocks = (unsigned)((n_groups + STEP_THREADS - 1) / STEP_THREADS);
    if (blocks == 0) return cudaSuccess;
    if (a.goal_mode == TS_GOAL_ORDERED) step_kernel<S, T, TS_GOAL_ORDERED><<<blocks, STEP_THREADS, 0, stream>>>(a);
    else step_kernel<S, T, TS_GOAL_SET><<<blocks, STEP_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace ts
