// ts_valid.cuh -- per-env mask of the moves that change the state
// (TilerSliderEnv.get_valid_moves, explainrl/environment/environment.py:149-171), computed
// from the occupancy and wall bitboards with four shifts (no slide is run).
#pragma once
#include "ts_common.cuh"
#include "../../include/tiler_slider.h"

namespace ts {

// A move changes the state iff SOME tile has an empty cell right ahead of it: a tile whose next cell
// is a wall or the edge stays, a tile behind another tile moves exactly when that one does, and the
// chain ends at a wall -- so "nothing moves" == "no tile faces an empty cell".  Four shifts of the
// occupancy board instead of four slides (get_valid_moves itself copies the state and tries the
// four moves, environment.py:162-169; same answers, held by the parity tests).
template <int S> __device__ __forceinline__ uint32_t valid_mask_of(uint64_t occ, uint64_t walls) {
    constexpr int BS = board_stride(S);
    constexpr uint64_t CELLS = [] { uint64_t m = 0; for (int r = 0; r < S; ++r) for (int c = 0; c < S; ++c) m |= 1ull << (r * BS + c); return m; }();
    constexpr uint64_t COL0 = [] { uint64_t m = 0; for (int r = 0; r < S; ++r) m |= 1ull << (r * BS); return m; }();
    const uint64_t open = CELLS & ~walls & ~occ;               // cells a tile can enter (sentinels and bits past the board excluded)
    uint64_t left = occ >> 1, right = occ << 1;
    if constexpr (!padded_board(S)) {                          // compact boards have no sentinel column between the rows
        left = (occ & ~COL0) >> 1;
        right = (occ & ~(COL0 << (S - 1))) << 1;
    }
    return ((occ >> BS) & open ? 1u : 0u) | ((occ << BS) & open ? 2u : 0u) | (left & open ? 4u : 0u) | (right & open ? 8u : 0u);
}

template <int S, int T>
__global__ void __launch_bounds__(256) valid_kernel(const __grid_constant__ ts_valid_args a) {
    constexpr int PW = pos_bytes(T), PR = (T + 3) / 4, NB = board_bytes(S), NWORDS = (NB + 3) / 4;
    const size_t n_groups = (size_t)((a.n_envs + GROUP - 1) / GROUP);
    size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= n_groups) return;
    g += (size_t)a.first_env / GROUP;
    const size_t e0 = g * GROUP;
    uint32_t praw[PW];
    ld_words<PW>(a.d_pos + e0 * PW, praw);
    BoardGroup<NB> walls;
    walls.load(a.d_walls, (size_t)a.capacity, g);
    uint32_t mask4 = 0;
#pragma unroll
    for (int e = 0; e < GROUP; ++e) {
        uint32_t q0[PR], bw[NWORDS];
        group_elem<PW>(praw, e, q0);
        walls.get(e, bw);
        mask4 |= valid_mask_of<S>(occupancy<S, T>(q0), board64(bw)) << (8 * e);
    }
    __stcs(reinterpret_cast<unsigned int*>(a.d_mask + e0), mask4);
}

template <int S, int T>
inline cudaError_t launch_valid(const ts_valid_args& a, cudaStream_t stream) {
    const size_t n_groups = (size_t)((a.n_envs + GROUP - 1) / GROUP);
    const unsigned blocks = (unsigned)((n_groups + 255) / 256);
    if (blocks == 0) return cudaSuccess;
    valid_kernel<S, T><<<blocks, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

// ---- standalone goal check (GameState.is_won, state.py:172-186) -------------------------------
template <int S, int T>
__global__ void __launch_bounds__(256) goal_kernel(const __grid_constant__ ts_goal_args a) {
    constexpr int PW = pos_bytes(T), PR = (T + 3) / 4, NB = board_bytes(S), NWORDS = (NB + 3) / 4;
    const size_t n_groups = (size_t)((a.n_envs + GROUP - 1) / GROUP);
    size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= n_groups) return;
    g += (size_t)a.first_env / GROUP;
    const size_t e0 = g * GROUP;
    uint32_t praw[PW];
    ld_words<PW>(a.d_pos + e0 * PW, praw);
    uint32_t won4 = 0;
    if (a.goal_mode == TS_GOAL_ORDERED) {
        uint32_t traw[PW];
        ld_words<PW>(a.d_targets_packed + e0 * PW, traw);
#pragma unroll
        for (int e = 0; e < GROUP; ++e) {
            uint32_t q[PR], t[PR];
            group_elem<PW>(praw, e, q);
            group_elem<PW>(traw, e, t);
            bool won = true;
#pragma unroll
            for (int w = 0; w < PR; ++w) won &= q[w] == t[w];
            won4 |= (won ? 1u : 0u) << (8 * e);
        }
    } else {
        BoardGroup<NB> tboard;
        tboard.load(a.d_targets_packed, (size_t)a.capacity, g);
#pragma unroll
        for (int e = 0; e < GROUP; ++e) {
            uint32_t q[PR], tw[NWORDS];
            group_elem<PW>(praw, e, q);
            tboard.get(e, tw);
            const uint64_t tb = board64(tw);
            won4 |= (occupancy<S, T>(q) == tb ? 1u : 0u) << (8 * e);
        }
    }
    if (a.never_win) won4 = 0;
    __stcs(reinterpret_cast<unsigned int*>(a.d_won + e0), won4);
}

template <int S, int T>
inline cudaError_t launch_goal(const ts_goal_args& a, cudaStream_t stream) {
    const size_t n_groups = (size_t)((a.n_envs + GROUP - 1) / GROUP);
    const unsigned blocks = (unsigned)((n_groups + 255) / 256);
    if (blocks == 0) return cudaSuccess;
    goal_kernel<S, T><<<blocks, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace ts
