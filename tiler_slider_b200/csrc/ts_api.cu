// ts_api.cu -- C-ABI entry points (include/tiler_slider.h), argument validation, and the
// load-time kernels: K1 ts_encode, K0 ts_synth (K3 ts_observe lives in ts_observe.cu).
//
// Reference map (paths relative to the reference checkout):
//   ts_encode   explainrl/environment/state.py:61-73        GameState.__init__ (is_blocked + lists)
//   ts_synth    explainrl/environment/environment.py:221-226 create_simple_env recipe
//   ts_observe  explainrl/environment/state.py:188-211      get_state_array (kernel in ts_observe.cu)
//   ts_step     see ts_step.cuh;  ts_valid_moves see ts_valid.cuh
#include <cstdarg>
#include <cstdio>
#include "ts_common.cuh"
#include "../../include/tiler_slider.h"

namespace ts {
#define TS_DECL(S)                                                              \
    cudaError_t step_dispatch_s##S(const ts_step_args&, cudaStream_t);          \
    cudaError_t valid_dispatch_s##S(const ts_valid_args&, cudaStream_t);        \
    cudaError_t goal_dispatch_s##S(const ts_goal_args&, cudaStream_t);
TS_DECL(1) TS_DECL(2) TS_DECL(3) TS_DECL(4) TS_DECL(5) TS_DECL(6) TS_DECL(7) TS_DECL(8)
#undef TS_DECL
cudaError_t observe_dispatch(const ts_observe_args&, cudaStream_t);
cudaError_t wide_step_dispatch(const ts_step_args&, cudaStream_t);
cudaError_t wide_valid_dispatch(const ts_valid_args&, cudaStream_t);
cudaError_t wide_goal_dispatch(const ts_goal_args&, cudaStream_t);
cudaError_t generic_step_dispatch(const ts_step_args&, cudaStream_t);
cudaError_t generic_goal_dispatch(const ts_goal_args&, cudaStream_t);
cudaError_t generic_valid_dispatch(const ts_valid_args&, cudaStream_t);

static thread_local char g_err[256] = "ok";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_result(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_shape(int size, int n_tiles, int64_t first, int64_t n, int64_t cap) {
    if (size < 1 || size > MAX_SIZE) return fail(TS_E_BAD_SIZE, "size %d outside 1..%d", size, MAX_SIZE);
    if (n_tiles < 0 || n_tiles > MAX_TILES_ANY) return fail(TS_E_BAD_TILES, "n_tiles %d outside 0..%d", n_tiles, MAX_TILES_ANY);
    if (cap >= (int64_t)1 << 33) return fail(TS_E_BAD_CAPACITY, "capacity %lld too large (max 2^33 - 128 envs per call)", (long long)cap);
    if (cap <= 0 || cap % CAP_ALIGN != 0) return fail(TS_E_BAD_CAPACITY, "capacity %lld is not a positive multiple of %d", (long long)cap, CAP_ALIGN);
    if (first < 0 || n < 0 || first % GROUP != 0 || first + n > cap)
        return fail(TS_E_BAD_RANGE, "env range [%lld, %lld) invalid for capacity %lld (first_env must be a multiple of %d)",
                    (long long)first, (long long)(first + n), (long long)cap, GROUP);
    return 0;
}

// value of bit `bit` of a stored WALL board: blocked cells, plus (padded boards) the sentinel
// column S of every row and every bit past the last row
__device__ __forceinline__ int wall_bit(const uint8_t* blocked_cells, int S, int bit) {
    const int bs = board_stride(S);
    const int r = bit / bs, c = bit % bs;
    if (r >= S) return padded_board(S) ? 1 : 0;
    if (c >= S) return 1;
    return blocked_cells[r * S + c] ? 1 : 0;
}

// Write the wall board(s) of one env from a 0/1 cell map (all three board classes).
__device__ void store_walls(uint8_t* d_walls, size_t cap, size_t env, int S, const uint8_t* cellmap) {
    if (wide_board(S)) {
        const int lines = 2 * wide_line_words(S);      // S rounded up to even
        uint16_t* w = reinterpret_cast<uint16_t*>(d_walls) + env * lines;   // [axis][env][line]
        const size_t plane = cap * lines;
        const int lead = wide_line_lead(S);           // S <= 14: cells at bits 1..S between two sentinels
        for (int l = 0; l < lines; ++l) {
            uint32_t col = 0, row = 0;
            if (l < S && lead) col = row = 1u | (1u << (S + 1));   // edge sentinels (ts_wide.cu)
            if (l < S)
                for (int k = 0; k < S; ++k) {
                    if (cellmap[k * S + l]) col |= 1u << (k + lead);    // column l, bit = row
                    if (cellmap[l * S + k]) row |= 1u << (k + lead);    // row l, bit = column
                }
            w[0 * plane + l] = (uint16_t)col;
            w[1 * plane + l] = (uint16_t)row;
        }
        return;
    }
    const int nb = board_bytes(S);
    for (int b = 0; b < nb; ++b) {
        uint32_t v = 0;
        for (int k = 0; k < 8; ++k) v |= (uint32_t)wall_bit(cellmap, S, 8 * b + k) << k;
        d_walls[board_byte_addr(nb, cap, env, b)] = (uint8_t)v;
    }
}

// Write the set-goal target board of one env from its target cells (row*S+col each).
__device__ void store_target_board(uint8_t* d_tb, size_t cap, size_t env, int S, int T, const int* cells, int n) {
    if (wide_board(S)) {
        uint16_t rows[16];
        for (int r = 0; r < 16; ++r) rows[r] = 0;
        for (int t = 0; t < n; ++t) rows[cells[t] / S] |= (uint16_t)(1u << (cells[t] % S));
        int distinct = 0;
        for (int r = 0; r < 16; ++r) distinct += __popc((uint32_t)rows[r]);
        // the wide kernels test "every tile on a target cell"; that is set equality only when
        // there are exactly T distinct targets (state.py:185-186): the count travels in word 16
        uint16_t* o = reinterpret_cast<uint16_t*>(d_tb) + env * WIDE_TARGET_WORDS;
        for (int r = 0; r < 16; ++r) o[r] = rows[r];
        o[16] = (uint16_t)distinct;
        return;
    }
    const int nb = board_bytes(S), bs = board_stride(S);
    for (int b = 0; b < nb; ++b) {
        uint32_t v = 0;
        for (int t = 0; t < n; ++t) {
            const int bit = (cells[t] / S) * bs + (cells[t] % S);
            if ((bit >> 3) == b) v |= 1u << (bit & 7);
        }
        d_tb[board_byte_addr(nb, cap, env, b)] = (uint8_t)v;
    }
}

// ---- K1 encode --------------------------------------------------------------------------------
__global__ void encode_kernel(const ts_encode_args a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_envs) return;
    const size_t env = (size_t)(a.first_env + i), cap = (size_t)a.capacity;
    const int S = a.size, T = a.n_tiles, NT = a.n_targets, pw = pos_bytes(T), ps = pos_stride(S);
    store_walls(a.d_walls, cap, env, S, a.d_blocked + (size_t)i * S * S);
    const uint8_t* tl = a.d_tiles + (size_t)i * T * 2;
    for (int t = 0; t < pw; ++t) {
        const uint8_t v = t < T ? (uint8_t)(tl[2 * t] * ps + tl[2 * t + 1]) : 0;
        a.d_init[env * pw + t] = v;
        a.d_pos[env * pw + t] = v;
    }
    const uint8_t* tg = a.d_targets + (size_t)i * NT * 2;
    if (a.goal_mode == TS_GOAL_ORDERED) {
        for (int t = 0; t < pw; ++t)
            a.d_targets_packed[env * pw + t] = (t < NT && t < T) ? (uint8_t)(tg[2 * t] * ps + tg[2 * t + 1]) : 0;
    } else {
        int cells[MAX_SIZE * MAX_SIZE];
        const int n = NT < MAX_SIZE * MAX_SIZE ? NT : MAX_SIZE * MAX_SIZE;
        for (int t = 0; t < n; ++t) cells[t] = tg[2 * t] * S + tg[2 * t + 1];
        store_target_board(a.d_targets_packed, cap, env, S, T, cells, n);
    }
}

// ---- K0 synth -----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void synth_kernel(const ts_synth_args a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_envs) return;
    const size_t env = (size_t)(a.first_env + i), cap = (size_t)a.capacity;
    const int S = a.size, T = a.n_tiles, W = a.n_walls, pw = pos_bytes(T);
    const int cells = S * S, draws = W + 2 * T;
    uint64_t rng = a.seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(a.env_index_base + i + 1));
    splitmix64(rng);
    uint8_t perm[MAX_SIZE * MAX_SIZE];
    for (int c = 0; c < cells; ++c) perm[c] = (uint8_t)c;
    for (int d = 0; d < draws; ++d) {  // partial Fisher-Yates: prefix of a uniform permutation
        const uint32_t r = (uint32_t)(splitmix64(rng) >> 32);
        const int j = d + (int)(((uint64_t)r * (uint32_t)(cells - d)) >> 32);
        const uint8_t t = perm[d];
        perm[d] = perm[j];
        perm[j] = t;
    }
    const int ps = pos_stride(S);
    uint8_t cellmap[MAX_SIZE * MAX_SIZE];
    for (int c = 0; c < cells; ++c) cellmap[c] = 0;
    for (int d = 0; d < W; ++d) cellmap[perm[d]] = 1;
    store_walls(a.d_walls, cap, env, S, cellmap);
    for (int t = 0; t < pw; ++t) {
        uint8_t v = 0;
        if (t < T) { const int c = perm[W + t]; v = (uint8_t)((c / S) * ps + (c % S)); }
        a.d_init[env * pw + t] = v;
        a.d_pos[env * pw + t] = v;
    }
    if (a.goal_mode == TS_GOAL_ORDERED) {
        for (int t = 0; t < pw; ++t) {
            uint8_t v = 0;
            if (t < T) { const int c = perm[W + T + t]; v = (uint8_t)((c / S) * ps + (c % S)); }
            a.d_targets_packed[env * pw + t] = v;
        }
    } else {
        int tcells[MAX_TILES_ANY];
        for (int t = 0; t < T; ++t) tcells[t] = perm[W + T + t];
        store_target_board(a.d_targets_packed, cap, env, S, T, tcells, T);
    }
}

}  // namespace ts

using namespace ts;

extern "C" {

int ts_version(void) { return TS_VERSION; }
const char* ts_last_error_string(void) { return g_err; }
int ts_pos_bytes(int n_tiles) { return pos_bytes(n_tiles); }
int ts_board_bytes(int size) { return board_bytes(size); }
int ts_board_stride(int size) { return board_stride(size); }
int ts_pos_stride(int size) { return pos_stride(size); }
int ts_plane_count(int n_bytes) { return plane_count(n_bytes); }
int ts_plane_width(int n_bytes, int k) { return plane_width(n_bytes, k); }
int ts_plane_offset(int n_bytes, int k) { return plane_offset(n_bytes, k); }
int ts_walls_bytes(int size) { return walls_bytes(size); }
int ts_target_board_bytes(int size) { return target_board_bytes(size); }
int ts_supported(int size, int n_tiles) { return size >= 1 && size <= MAX_SIZE && n_tiles >= 0 && n_tiles <= MAX_TILES_ANY; }

int ts_encode(const ts_encode_args* a, void* stream) {
    if (!a) return fail(TS_E_NULL_POINTER, "null args");
    if (int rc = check_shape(a->size, a->n_tiles, a->first_env, a->n_envs, a->capacity)) return rc;
    if (a->goal_mode != TS_GOAL_ORDERED && a->goal_mode != TS_GOAL_SET) return fail(TS_E_BAD_ARGUMENT, "goal_mode %d", a->goal_mode);
    if (a->n_targets < 0 || a->n_targets > MAX_SIZE * MAX_SIZE) return fail(TS_E_BAD_TILES, "n_targets %d", a->n_targets);
    if (!a->d_blocked || (!a->d_tiles && a->n_tiles) || (!a->d_targets && a->n_targets) || !a->d_walls || !a->d_targets_packed || !a->d_init || !a->d_pos)
        return fail(TS_E_NULL_POINTER, "null device pointer");
    if (a->n_envs == 0) return 0;
    const unsigned blocks = (unsigned)((a->n_envs + 127) / 128);
    encode_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(*a);
    return cuda_result(cudaGetLastError(), "ts_encode launch");
}

int ts_synth(const ts_synth_args* a, void* stream) {
    if (!a) return fail(TS_E_NULL_POINTER, "null args");
    if (int rc = check_shape(a->size, a->n_tiles, a->first_env, a->n_envs, a->capacity)) return rc;
    if (a->goal_mode != TS_GOAL_ORDERED && a->goal_mode != TS_GOAL_SET) return fail(TS_E_BAD_ARGUMENT, "goal_mode %d", a->goal_mode);
    if (a->n_walls < 0 || a->n_walls + 2 * a->n_tiles > a->size * a->size)
        return fail(TS_E_BAD_ARGUMENT, "n_walls + 2*n_tiles = %d exceeds the %d cells", a->n_walls + 2 * a->n_tiles, a->size * a->size);
    if (!a->d_walls || !a->d_targets_packed || !a->d_init || !a->d_pos) return fail(TS_E_NULL_POINTER, "null device pointer");
    if (a->n_envs == 0) return 0;
    const unsigned blocks = (unsigned)((a->n_envs + 127) / 128);
    synth_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(*a);
    return cuda_result(cudaGetLastError(), "ts_synth launch");
}

int ts_step(const ts_step_args* a, void* stream) {
    if (!a) return fail(TS_E_NULL_POINTER, "null args");
    if (int rc = check_shape(a->size, a->n_tiles, a->first_env, a->n_envs, a->capacity)) return rc;
    if (!ts_supported(a->size, a->n_tiles)) return fail(TS_E_UNSUPPORTED, "no step kernel for size %d, n_tiles %d", a->size, a->n_tiles);
    if (a->goal_mode != TS_GOAL_ORDERED && a->goal_mode != TS_GOAL_SET) return fail(TS_E_BAD_ARGUMENT, "goal_mode %d", a->goal_mode);
    if (a->count_bytes != 1 && a->count_bytes != 4) return fail(TS_E_BAD_ARGUMENT, "count_bytes %d (1 or 4)", a->count_bytes);
    // 1-byte counters are bumped SWAR, four to a word: a frozen env (auto_reset = 0) resting at 255 would carry into its neighbour
    if (a->count_bytes == 1 && a->max_steps > (a->auto_reset ? 255 : 254)) return fail(TS_E_BAD_ARGUMENT, "max_steps %d needs count_bytes 4", a->max_steps);
    if (!a->d_walls || !a->d_targets_packed || !a->d_init || !a->d_pos || !a->d_step_count || !a->d_actions || !a->d_reward)
        return fail(TS_E_NULL_POINTER, "null device pointer");
    if (!a->d_done && !a->d_flags) return fail(TS_E_NULL_POINTER, "need d_done or d_flags");
    if (!a->auto_reset && !a->d_flags) return fail(TS_E_NULL_POINTER, "auto_reset=0 needs d_flags (done state is carried there)");
    const void* ptrs[] = {a->d_walls, a->d_targets_packed, a->d_init, a->d_pos, a->d_step_count, a->d_actions,
                          a->d_reward, a->d_done, a->d_flags, a->d_terminal_pos};
    for (const void* p : ptrs)
        if (p && !aligned16(p)) return fail(TS_E_MISALIGNED, "device pointer %p is not 16-byte aligned", p);
    if (a->n_envs == 0) return 0;
    ts_step_args args = *a;
    if (args.max_steps < 1) args.max_steps = 1;   // step_count >= max_steps holds on the first step either way
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_tiles == 0 || a->n_tiles > MAX_TILES)      // nothing to slide / more tiles than the register kernels hold (ts_generic.cu)
        return cuda_result(generic_step_dispatch(args, st), "ts_step launch");
    // the bitboard kernels own whole 4-env groups: the ragged end of the range goes to the generic kernel
    const int64_t ragged = wide_board(a->size) ? 0 : a->n_envs % GROUP;
    args.n_envs -= ragged;
    cudaError_t e = cudaSuccess;
    if (args.n_envs > 0) {
        switch (a->size) {
            case 1: e = step_dispatch_s1(args, st); break;
            case 2: e = step_dispatch_s2(args, st); break;
            case 3: e = step_dispatch_s3(args, st); break;
            case 4: e = step_dispatch_s4(args, st); break;
            case 5: e = step_dispatch_s5(args, st); break;
            case 6: e = step_dispatch_s6(args, st); break;
            case 7: e = step_dispatch_s7(args, st); break;
            case 8: e = step_dispatch_s8(args, st); break;
            default: e = wide_step_dispatch(args, st); break;
        }
    }
    if (e == cudaSuccess && ragged) {
        args.first_env += args.n_envs;
        args.n_envs = ragged;
        e = generic_step_dispatch(args, st);
    }
    return cuda_result(e, "ts_step launch");
}

int ts_observe(const ts_observe_args* a, void* stream) {
    if (!a) return fail(TS_E_NULL_POINTER, "null args");
    if (int rc = check_shape(a->size, a->n_tiles, a->first_env, a->n_envs, a->capacity)) return rc;
    if (!a->d_walls || !a->d_targets_packed || !a->d_pos || !a->d_obs) return fail(TS_E_NULL_POINTER, "null device pointer");
    if (a->n_envs == 0) return 0;
    if (!aligned16(a->d_obs)) return fail(TS_E_MISALIGNED, "d_obs %p is not 16-byte aligned", (const void*)a->d_obs);
    return cuda_result(observe_dispatch(*a, (cudaStream_t)stream), "ts_observe launch");
}

int ts_valid_moves(const ts_valid_args* a, void* stream) {
    if (!a) return fail(TS_E_NULL_POINTER, "null args");
    if (int rc = check_shape(a->size, a->n_tiles, a->first_env, a->n_envs, a->capacity)) return rc;
    if (!ts_supported(a->size, a->n_tiles)) return fail(TS_E_UNSUPPORTED, "no kernel for size %d, n_tiles %d", a->size, a->n_tiles);
    if (!a->d_walls || !a->d_pos || !a->d_mask) return fail(TS_E_NULL_POINTER, "null device pointer");
    if (a->n_envs == 0) return 0;
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_tiles == 0)      // no tile, no move that changes anything
        return cuda_result(cudaMemsetAsync(a->d_mask + a->first_env, 0, (size_t)a->n_envs, st), "ts_valid_moves memset");
    if (a->n_tiles > MAX_TILES) return cuda_result(generic_valid_dispatch(*a, st), "ts_valid_moves launch");
    switch (a->size) {
        case 1: e = valid_dispatch_s1(*a, st); break;
        case 2: e = valid_dispatch_s2(*a, st); break;
        case 3: e = valid_dispatch_s3(*a, st); break;
        case 4: e = valid_dispatch_s4(*a, st); break;
        case 5: e = valid_dispatch_s5(*a, st); break;
        case 6: e = valid_dispatch_s6(*a, st); break;
        case 7: e = valid_dispatch_s7(*a, st); break;
        case 8: e = valid_dispatch_s8(*a, st); break;
        default: e = wide_valid_dispatch(*a, st); break;
    }
    return cuda_result(e, "ts_valid_moves launch");
}

int ts_goal_check(const ts_goal_args* a, void* stream) {
    if (!a) return fail(TS_E_NULL_POINTER, "null args");
    if (int rc = check_shape(a->size, a->n_tiles, a->first_env, a->n_envs, a->capacity)) return rc;
    if (!ts_supported(a->size, a->n_tiles)) return fail(TS_E_UNSUPPORTED, "no kernel for size %d, n_tiles %d", a->size, a->n_tiles);
    if (a->goal_mode != TS_GOAL_ORDERED && a->goal_mode != TS_GOAL_SET) return fail(TS_E_BAD_ARGUMENT, "goal_mode %d", a->goal_mode);
    if (!a->d_targets_packed || !a->d_pos || !a->d_won) return fail(TS_E_NULL_POINTER, "null device pointer");
    if (a->n_envs == 0) return 0;
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_tiles == 0 || a->n_tiles > MAX_TILES) return cuda_result(generic_goal_dispatch(*a, st), "ts_goal_check launch");
    switch (a->size) {
        case 1: e = goal_dispatch_s1(*a, st); break;
        case 2: e = goal_dispatch_s2(*a, st); break;
        case 3: e = goal_dispatch_s3(*a, st); break;
        case 4: e = goal_dispatch_s4(*a, st); break;
        case 5: e = goal_dispatch_s5(*a, st); break;
        case 6: e = goal_dispatch_s6(*a, st); break;
        case 7: e = goal_dispatch_s7(*a, st); break;
        case 8: e = goal_dispatch_s8(*a, st); break;
        default: e = wide_goal_dispatch(*a, st); break;
    }
    return cuda_result(e, "ts_goal_check launch");
}

}  // extern "C"
