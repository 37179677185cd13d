// ts_bfs.cu -- breadth-first state-space expansion with hash-partitioned dedup (BASELINE
// config 5).  The reference has no solver; the nearest thing is get_valid_moves
// (explainrl/environment/environment.py:149-171), which tries the four moves on copies.  The
// successor function here is exactly GameState.move (explainrl/environment/state.py:120-170)
// through slide_env<S,T>, and the goal test is is_won (state.py:172-186).  BFS itself is
// repo-defined (parity unpinned); its level histograms are pinned to a plain BFS over the
// reference's move (tests/golden/misc.json).
//
// State key (u64):  bits 0..31 (T <= 4) or 0..62 (T > 4, single puzzle): the canonical position
// word -- ordered goal: the position bytes as they are; set goal: the bytes sorted ascending
// (tiles are interchangeable); bits 32..62: puzzle id (T <= 4); bit 63: "this successor meets
// the goal" (carried through the exchange, ignored by dedup).  TS_BFS_NONE (all ones) marks a
// move that changed nothing.
//
// Kernels:  bfs_seed (initial keys), K4 bfs_expand (frontier x 4 moves -> successor keys),
// bfs_partition_count / bfs_partition_scatter (bucket successors by owner rank =
// hash(key) % n_ranks, ahead of the NCCL all-to-all done by the Python driver), K5
// bfs_hash_insert (open-addressing visited table, atomicCAS; emits the new keys = next
// frontier; optionally per-puzzle tallies, parent links and the goal successor of every puzzle),
// K4x bfs_expand_exchange (expand + bucket + store into the owners' inboxes over NVLink peer
// memory, one kernel), bfs_traceback / bfs_trace_step (shortest move strings from the parent links:
// on one rank / one step at a time when the links are spread over the owners).
#include "ts_common.cuh"
#include "../../include/tiler_slider.h"

namespace ts {

constexpr uint64_t BFS_NONE = ~0ull;
constexpr unsigned BFS_PERSISTENT_BLOCKS = 148 * 8;   // one wave of 256-thread CTAs on a B200
constexpr uint64_t BFS_WON_BIT = 1ull << 63;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// sort the T position bytes ascending (odd-even transposition network; T <= 8)
template <int T> __device__ __forceinline__ void sort_bytes(uint32_t (&q)[(T + 3) / 4]) {
    uint32_t b[T];
    static_for<0, T>([&](auto I) { constexpr int i = decltype(I)::value; b[i] = byte_of<i % 4>(q[i / 4]); });
#pragma unroll
    for (int round = 0; round < T; ++round)
#pragma unroll
        for (int i = round & 1; i + 1 < T; i += 2) {
            const uint32_t lo = min(b[i], b[i + 1]), hi = max(b[i], b[i + 1]);
            b[i] = lo; b[i + 1] = hi;
        }
#pragma unroll
    for (int w = 0; w < (T + 3) / 4; ++w) q[w] = 0;
    static_for<0, T>([&](auto I) { constexpr int i = decltype(I)::value; q[i / 4] |= b[i] << (8 * (i % 4)); });
}

template <int T> __device__ __forceinline__ uint64_t make_key(const uint32_t (&q)[(T + 3) / 4], uint64_t pid) {
    if constexpr (T <= 4) return (uint64_t)q[0] | (pid << 32);
    else return (uint64_t)q[0] | ((uint64_t)q[1] << 32);
}
template <int T> __device__ __forceinline__ void split_key(uint64_t key, uint32_t (&q)[(T + 3) / 4], uint64_t& pid) {
    q[0] = (uint32_t)key;
    if constexpr (T <= 4) pid = (key >> 32) & 0x7FFFFFFFull;
    else { q[1] = (uint32_t)(key >> 32) & 0x7FFFFFFFu; pid = 0; }
}

template <int S, int T>
__global__ void __launch_bounds__(256) bfs_seed_kernel(const ts_bfs_args a) {
    constexpr int PW = pos_bytes(T), PR = (T + 3) / 4;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n_items) return;
    uint32_t q[PR] = {};
    const uint8_t* p = a.d_init + (size_t)i * PW;
    static_for<0, T>([&](auto I) { constexpr int t = decltype(I)::value; q[t / 4] |= (uint32_t)p[t] << (8 * (t % 4)); });
    if (a.goal_mode == TS_GOAL_SET) sort_bytes<T>(q);
    a.d_out_keys[i] = make_key<T>(q, (uint64_t)i);
}

// the four successor keys of one frontier key (TS_BFS_NONE where the move changed nothing)
template <int S, int T>
__device__ __forceinline__ void successors(const ts_bfs_args& a, uint64_t in_key, uint64_t (&out)[4]) {
    constexpr int PW = pos_bytes(T), PR = (T + 3) / 4, NB = board_bytes(S);
    uint32_t q0[PR];
    uint64_t pid;
    split_key<T>(in_key, q0, pid);
    const size_t cap = (size_t)a.puzzle_capacity;
    const uint64_t walls = load_board_elem<NB>(a.d_walls, cap, (size_t)pid);
    uint64_t tboard = 0;
    uint32_t tq[PR] = {};
    if (a.goal_mode == TS_GOAL_SET) tboard = load_board_elem<NB>(a.d_targets_packed, cap, (size_t)pid);
    else {
        const uint8_t* p = a.d_targets_packed + (size_t)pid * PW;
        static_for<0, T>([&](auto I) { constexpr int t = decltype(I)::value; tq[t / 4] |= (uint32_t)p[t] << (8 * (t % 4)); });
    }
#pragma unroll
    for (uint32_t d = 0; d < 4; ++d) {
        uint32_t q[PR];
#pragma unroll
        for (int w = 0; w < PR; ++w) q[w] = q0[w];
        slide_env<S, T>(q, walls, d >> 1, (d & 1u) ^ 1u);
        bool won = a.never_win == 0;
        if (a.goal_mode == TS_GOAL_SET) {
            won &= occupancy<S, T>(q) == tboard;
            sort_bytes<T>(q);     // canonical form: tiles are interchangeable (state.py:185-186)
        } else {
#pragma unroll
            for (int w = 0; w < PR; ++w) won &= q[w] == tq[w];
        }
        bool same = true;
#pragma unroll
        for (int w = 0; w < PR; ++w) same &= q[w] == q0[w];
        uint64_t key = make_key<T>(q, pid) | (won ? BFS_WON_BIT : 0ull);
        if (same && !won) key = BFS_NONE;
        out[d] = key;
    }
}

// number of items of a launch: the host's n_items, or (device-driven levels) what an earlier
// kernel left in *d_n_items
__device__ __forceinline__ int64_t item_count(const ts_bfs_args& a) {
    if (!a.d_n_items) return a.n_items;
    const int64_t n = *a.d_n_items * a.n_items_scale;
    return n < a.n_items ? n : a.n_items;
}

// K4: thread = one frontier state, four successors (grid-stride: the grid may be smaller than
// the frontier when its size is only known on the device)
template <int S, int T>
__global__ void __launch_bounds__(256) bfs_expand_kernel(const ts_bfs_args a) {
    const int64_t n = item_count(a);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        uint64_t key[4];
        successors<S, T>(a, a.d_in_keys[i], key);
#pragma unroll
        for (int d = 0; d < 4; ++d) a.d_out_keys[4 * i + d] = key[d];
    }
}

// owner rank of a key (won bit excluded)
__device__ __forceinline__ uint32_t key_owner(uint64_t key, uint32_t n_ranks) {
    return (uint32_t)((mix64(key & ~BFS_WON_BIT) >> 40) % n_ranks);
}

// Warp-aggregated bucket bookkeeping: lanes holding keys of the same owner elect a leader
// (__match_any_sync) that bumps the block's shared counter once for all of them; with 2-8 owners
// and a billion keys a per-key atomic on 2-8 addresses is what the exchange would wait for.
// Returns this lane's slot inside the block's share of its owner's bucket.
__device__ __forceinline__ uint32_t block_bucket_slot(unsigned int* hist, bool live, uint32_t owner) {
    const unsigned active = __ballot_sync(0xFFFFFFFFu, live);
    uint32_t slot = 0;
    if (live) {
        const unsigned peers = __match_any_sync(active, owner);
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&hist[owner], (unsigned)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        slot = base + (uint32_t)__popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
    }
    return slot;
}

__global__ void __launch_bounds__(256) bfs_partition_count_kernel(const ts_bfs_args a) {
    __shared__ unsigned int hist[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const uint64_t key = i < a.n_items ? a.d_in_keys[i] : BFS_NONE;
    const bool live = key != BFS_NONE;
    block_bucket_slot(hist, live, live ? key_owner(key, (uint32_t)a.n_ranks) : 0u);
    __syncthreads();
    if (threadIdx.x < a.n_ranks && hist[threadIdx.x]) atomicAdd((unsigned long long*)&a.d_counts[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

// d_counts holds, on entry, the write cursor (= exclusive prefix offset) of every owner bucket;
// a block reserves its share of each bucket with one global atomic per owner
__global__ void __launch_bounds__(256) bfs_partition_scatter_kernel(const ts_bfs_args a) {
    __shared__ unsigned int hist[64];
    __shared__ unsigned long long base[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const uint64_t key = i < a.n_items ? a.d_in_keys[i] : BFS_NONE;
    const bool live = key != BFS_NONE;
    const uint32_t owner = live ? key_owner(key, (uint32_t)a.n_ranks) : 0u;
    const uint32_t slot = block_bucket_slot(hist, live, owner);
    __syncthreads();
    if (threadIdx.x < a.n_ranks && hist[threadIdx.x])
        base[threadIdx.x] = atomicAdd((unsigned long long*)&a.d_counts[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
    __syncthreads();
    if (live) {
        a.d_out_keys[base[owner] + slot] = key;
        // with parents: the key it was expanded from travels to the owner beside it
        if (a.d_out_parents) a.d_out_parents[base[owner] + slot] = a.d_parent_keys[i >> 2] & ~BFS_WON_BIT;
    }
}

// K4x: expand + bucket + exchange over NVLink peer memory in one kernel.  A block keeps its 1024
// successors in registers, builds the per-owner histogram in shared memory (warp-aggregated),
// reserves its share of every owner's inbox with ONE system-scope atomic per owner on that rank's
// arrival cursor, and stores the keys straight into the owners' inboxes -- for a remote owner
// these are posted writes through the NVLink mapping.  No partition pass over the successors, no
// size exchange, no all-to-all; nothing here waits on another GPU.
constexpr int XCHG_STATES = 4;     // frontier states per thread and round of K4x

template <int S, int T>
__global__ void __launch_bounds__(256) bfs_expand_exchange_kernel(const ts_bfs_args a) {
    __shared__ unsigned int hist[64];
    __shared__ unsigned long long base[64];
    const int64_t n_items = item_count(a);
    constexpr int64_t ROUND = 256 * XCHG_STATES;
    // grid-stride over rounds of 1,024 frontier keys = 4,096 successors per block (the frontier size
    // may only be known on the device, ts_bfs_args.d_n_items).  One reservation per owner and round:
    // the arrival cursors are single addresses that every block of every rank adds to, so the fewer
    // and larger the reservations the better (at 256 keys per round an 8-GPU search spent more time
    // queueing on them than inserting).
    for (int64_t first = (int64_t)blockIdx.x * ROUND; first < n_items; first += (int64_t)gridDim.x * ROUND) {
        if (threadIdx.x < 64) hist[threadIdx.x] = 0;
        __syncthreads();
        uint64_t key[XCHG_STATES][4];
        uint32_t where[XCHG_STATES][4];                    // owner | slot << 8
#pragma unroll
        for (int s = 0; s < XCHG_STATES; ++s) {
            const int64_t i = first + s * 256 + threadIdx.x;
#pragma unroll
            for (int d = 0; d < 4; ++d) key[s][d] = BFS_NONE;
            if (i < n_items) successors<S, T>(a, a.d_in_keys[i], key[s]);
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const bool live = key[s][d] != BFS_NONE;
                const uint32_t owner = live ? key_owner(key[s][d], (uint32_t)a.n_ranks) : 0u;
                where[s][d] = owner | (block_bucket_slot(hist, live, owner) << 8);
            }
        }
        __syncthreads();
        if (threadIdx.x < a.n_ranks) {
            const unsigned n = hist[threadIdx.x];
            unsigned long long b = ~0ull;
            if (n) {
                uint64_t* peer = a.d_peer_bufs[threadIdx.x];
                b = atomicAdd_system((unsigned long long*)(peer + a.parity), (unsigned long long)n);
                if (b + n > (unsigned long long)a.inbox_capacity) {      // inbox full: drop and report
                    peer[2] = 1;
                    a.d_counts[2] = 1;
                    b = ~0ull;
                }
                atomicAdd((unsigned long long*)&a.d_counts[3], (unsigned long long)n);
            }
            base[threadIdx.x] = b;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < XCHG_STATES; ++s)
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (key[s][d] == BFS_NONE) continue;
                const uint32_t owner = where[s][d] & 0xFFu;
                const unsigned long long b = base[owner];
                if (b == ~0ull) continue;
                uint64_t* inbox = a.d_peer_bufs[owner] + TS_BFS_XHDR + (int64_t)a.parity * a.inbox_capacity;
                inbox[b + (where[s][d] >> 8)] = key[s][d];
            }
        __syncthreads();                                                  // hist / base are reused by the next round
    }
    __threadfence_system();
}

// K5: insert keys into the open-addressing visited table (EMPTY = all ones).  New keys are
// appended to d_out_keys through the cursor d_counts[0]; d_counts[1] counts won successors seen,
// d_counts[2] is set when the table is full.
__global__ void __launch_bounds__(256) bfs_hash_insert_kernel(const ts_bfs_args a) {
    const int64_t n_items = item_count(a);
    const unsigned lane = threadIdx.x & 31u;
    // grid-stride over whole blocks of 256 items, so every warp runs the same number of rounds
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n_items; base += (int64_t)gridDim.x * 256) {
    const int64_t i = base + threadIdx.x;
    const uint64_t raw = i < n_items ? a.d_in_keys[i] : BFS_NONE;
    const bool live = raw != BFS_NONE;
    const uint64_t key = raw & ~BFS_WON_BIT;
    bool is_new = false, full = false;
    if (live) {
        const uint64_t mask = (uint64_t)a.table_capacity - 1;
        uint64_t slot = mix64(key) & mask;
        int64_t probe = 0;
        for (; probe < a.table_capacity; ++probe) {
            // Three successors in four are revisits: look before you CAS.  A slot never changes once
            // it holds a key, so a plain (L2) load that sees a key is final; one that sees EMPTY may
            // be out of date, and the CAS that follows settles it.
            unsigned long long old = __ldcg((const unsigned long long*)&a.d_table[slot]);
            if (old == BFS_NONE)
                old = atomicCAS((unsigned long long*)&a.d_table[slot], (unsigned long long)BFS_NONE, (unsigned long long)key);
            else if (old != key) { slot = (slot + 1) & mask; continue; }
            if (old == BFS_NONE) {
                if (a.d_table_parent) a.d_table_parent[slot] = a.d_parent_keys ? (a.d_parent_keys[a.parent_per_item ? i : (i >> 2)] & ~BFS_WON_BIT) : BFS_NONE;
                is_new = true;
                break;
            }
            if (old == key) break;
            slot = (slot + 1) & mask;
        }
        full = probe == a.table_capacity;
    }
    // one cursor bump per warp, not per key: a frontier of 1e8 new keys would otherwise queue
    // 1e8 atomics on one address.  Keys of a warp stay in lane order in the output.
    const bool won = live && (raw & BFS_WON_BIT) != 0;
    const unsigned new_mask = __ballot_sync(0xFFFFFFFFu, is_new), won_mask = __ballot_sync(0xFFFFFFFFu, won);
    if (new_mask) {
        const int leader = __ffs(new_mask) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd((unsigned long long*)&a.d_counts[0], (unsigned long long)__popc(new_mask));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (is_new) {
            const unsigned long long pos = base + (unsigned long long)__popc(new_mask & ((1u << lane) - 1u));
            if ((int64_t)pos < a.out_capacity) a.d_out_keys[pos] = raw;
            else full = true;
        }
    }
    // optional per-puzzle tallies; a frontier is roughly in puzzle order, so the new keys of a warp
    // mostly share a puzzle and one lane adds for all of them
    const uint32_t pid = a.n_tiles <= 4 ? (uint32_t)(key >> 32) & 0x7FFFFFFFu : 0u;
    if (a.d_states_per_puzzle && is_new) {
        const unsigned peers = __match_any_sync(new_mask, pid);
        if ((int)lane == __ffs(peers) - 1)
            atomicAdd((unsigned long long*)&a.d_states_per_puzzle[pid], (unsigned long long)__popc(peers));
    }
    if (a.d_solve_depth && won) {
        const int before = atomicMin(&a.d_solve_depth[pid], a.depth);
        if (before > a.depth && a.d_goal_keys) {
            a.d_goal_keys[pid] = raw;
            if (a.d_goal_parents && a.d_parent_keys)
                a.d_goal_parents[pid] = a.d_parent_keys[a.parent_per_item ? i : (i >> 2)] & ~BFS_WON_BIT;
        }
    }
    if (won_mask) {
        const int leader = __ffs(won_mask) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd((unsigned long long*)&a.d_counts[1], (unsigned long long)__popc(won_mask));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (won) {
            const unsigned long long w = base + (unsigned long long)__popc(won_mask & ((1u << lane) - 1u));
            if (a.d_won_keys && (int64_t)w < a.won_capacity) a.d_won_keys[w] = raw;
        }
    }
    if (full) a.d_counts[2] = 1;
    }
}

// slot of `key` in the visited table, or -1
__device__ __forceinline__ int64_t table_find(const uint64_t* table, int64_t capacity, uint64_t key) {
    const uint64_t mask = (uint64_t)capacity - 1;
    uint64_t slot = mix64(key) & mask;
    for (int64_t probe = 0; probe < capacity; ++probe) {
        const uint64_t v = table[slot];
        if (v == key) return (int64_t)slot;
        if (v == BFS_NONE) return -1;
        slot = (slot + 1) & mask;
    }
    return -1;
}

// the move that leads from `parent` to `key` (both without the goal bit): the smallest move
// index whose slide of the parent reproduces the child; -1 if none does
template <int S, int T>
__device__ __forceinline__ int move_between(const ts_bfs_args& a, uint64_t parent, uint64_t key) {
    constexpr int PR = (T + 3) / 4, NB = board_bytes(S);
    uint32_t q0[PR];
    uint64_t pid;
    split_key<T>(parent, q0, pid);
    const uint64_t walls = load_board_elem<NB>(a.d_walls, (size_t)a.puzzle_capacity, (size_t)pid);
    int move = -1;
    for (uint32_t d = 0; d < 4 && move < 0; ++d) {
        uint32_t q[PR];
#pragma unroll
        for (int w = 0; w < PR; ++w) q[w] = q0[w];
        slide_env<S, T>(q, walls, d >> 1, (d & 1u) ^ 1u);
        if (a.goal_mode == TS_GOAL_SET) sort_bytes<T>(q);
        if (make_key<T>(q, pid) == key) move = (int)d;
    }
    return move;
}

// shortest move string of each goal state: walk the parent chain, recover every move by
// re-sliding the parent, reverse at the end
template <int S, int T>
__global__ void __launch_bounds__(128) bfs_traceback_kernel(const ts_bfs_args a) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= a.n_items) return;
    uint8_t* out = a.d_moves + (size_t)i * a.max_moves;
    uint64_t key = a.d_in_keys[i];
    if (key == BFS_NONE) { a.d_lengths[i] = -1; return; }
    key &= ~BFS_WON_BIT;
    int len = 0;
    bool ok = true;
    {   // the last link is given, not looked up: the goal state's table entry is that of its FIRST
        // visit, which for a puzzle that starts on its goal is the root (no parent)
        const uint64_t from = a.d_parent_keys[i];
        const int move = from == BFS_NONE ? -1 : move_between<S, T>(a, from & ~BFS_WON_BIT, key);
        if (move < 0) { a.d_lengths[i] = -1; return; }
        out[len++] = (uint8_t)move;
        key = from & ~BFS_WON_BIT;
    }
    while (ok) {
        const int64_t slot = table_find(a.d_table, a.table_capacity, key);
        if (slot < 0) { ok = false; break; }
        const uint64_t parent = a.d_table_parent[slot];
        if (parent == BFS_NONE) break;                     // reached a root
        if (len >= a.max_moves) { ok = false; break; }
        const int move = move_between<S, T>(a, parent, key);
        if (move < 0) { ok = false; break; }
        out[len++] = (uint8_t)move;
        key = parent;
    }
    if (!ok) { a.d_lengths[i] = -1; return; }
    for (int l = 0, r = len - 1; l < r; ++l, --r) { const uint8_t t = out[l]; out[l] = out[r]; out[r] = t; }
    a.d_lengths[i] = len;
}

// one traceback step over a visited set that is spread over several owners: this rank answers
// for the keys it owns, every other item gets the smallest int64 so that a MAX all-reduce over
// the owners assembles the step (see ts_bfs_trace_step in the header)
constexpr uint64_t TRACE_NOT_MINE = 0x8000000000000000ull, TRACE_BROKEN = 0xFFFFFFFFFFFFFFFEull;
template <int S, int T>
__global__ void __launch_bounds__(128) bfs_trace_step_kernel(const ts_bfs_args a) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= a.n_items) return;
    uint64_t key = a.d_in_keys[i], out = TRACE_NOT_MINE;
    int move = 0;
    if (key != BFS_NONE) {
        key &= ~BFS_WON_BIT;
        if (a.d_parent_keys) {                       // the last link of a solution: given, on every rank
            out = a.d_parent_keys[i];
            move = out == BFS_NONE ? -1 : move_between<S, T>(a, out & ~BFS_WON_BIT, key);
            out = move < 0 ? TRACE_BROKEN : (out & ~BFS_WON_BIT);
            move = move < 0 ? 0 : move;
        } else if (a.n_ranks <= 1 || key_owner(key, (uint32_t)a.n_ranks) == (uint32_t)a.rank) {
            const int64_t slot = table_find(a.d_table, a.table_capacity, key);
            out = slot < 0 ? TRACE_BROKEN : a.d_table_parent[slot];
            if (slot >= 0 && out != BFS_NONE) {
                move = move_between<S, T>(a, out, key);
                if (move < 0) { out = TRACE_BROKEN; move = 0; }
            }
        }
    }
    a.d_out_keys[i] = out;
    a.d_moves[i] = (uint8_t)move;
}

template <int S> static cudaError_t bfs_dispatch_T(int op, const ts_bfs_args& a, cudaStream_t st) {
    unsigned blocks = (unsigned)((a.n_items + 255) / 256);
    if (op == 3) blocks = (unsigned)((a.n_items + 256 * XCHG_STATES - 1) / (256 * XCHG_STATES));
    if (a.d_n_items && blocks > BFS_PERSISTENT_BLOCKS) blocks = BFS_PERSISTENT_BLOCKS;   // size unknown on the host: grid-stride
#define TS_BFS_CASE(T)                                                                  \
    case T:                                                                             \
        if (op == 0) bfs_seed_kernel<S, T><<<blocks, 256, 0, st>>>(a);                  \
        else if (op == 1) bfs_expand_kernel<S, T><<<blocks, 256, 0, st>>>(a);           \
        else if (op == 3) bfs_expand_exchange_kernel<S, T><<<blocks, 256, 0, st>>>(a);  \
        else if (op == 4) bfs_trace_step_kernel<S, T><<<(unsigned)((a.n_items + 127) / 128), 128, 0, st>>>(a); \
        else bfs_traceback_kernel<S, T><<<(unsigned)((a.n_items + 127) / 128), 128, 0, st>>>(a); \
        break;
    switch (a.n_tiles) {
        TS_BFS_CASE(1) TS_BFS_CASE(2) TS_BFS_CASE(3) TS_BFS_CASE(4)
        TS_BFS_CASE(5) TS_BFS_CASE(6) TS_BFS_CASE(7) TS_BFS_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef TS_BFS_CASE
    return cudaGetLastError();
}

static cudaError_t bfs_dispatch(int op, const ts_bfs_args& a, cudaStream_t st) {
    switch (a.size) {
        case 1: return bfs_dispatch_T<1>(op, a, st);
        case 2: return bfs_dispatch_T<2>(op, a, st);
        case 3: return bfs_dispatch_T<3>(op, a, st);
        case 4: return bfs_dispatch_T<4>(op, a, st);
        case 5: return bfs_dispatch_T<5>(op, a, st);
        case 6: return bfs_dispatch_T<6>(op, a, st);
        case 7: return bfs_dispatch_T<7>(op, a, st);
        case 8: return bfs_dispatch_T<8>(op, a, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ts

using namespace ts;

static int bfs_check(const ts_bfs_args* a, bool needs_shape) {
    if (!a) return TS_E_NULL_POINTER;
    if (a->n_items < 0) return TS_E_BAD_RANGE;
    if (needs_shape) {
        if (a->size < 1 || a->size > 8) return TS_E_UNSUPPORTED;       // bitboard classes only
        if (a->n_tiles < 1 || a->n_tiles > MAX_TILES) return TS_E_BAD_TILES;
        if (a->puzzle_capacity <= 0 || a->puzzle_capacity % CAP_ALIGN) return TS_E_BAD_CAPACITY;
        if (a->n_tiles > 4 && a->puzzle_capacity > CAP_ALIGN) return TS_E_UNSUPPORTED;   // T > 4: single puzzle, no id bits
    }
    return 0;
}

extern "C" {

int ts_bfs_seed(const ts_bfs_args* a, void* stream) {
    if (int rc = bfs_check(a, true)) return rc;
    if (a->n_items == 0) return 0;
    if (!a->d_init || !a->d_out_keys) return TS_E_NULL_POINTER;
    return (int)bfs_dispatch(0, *a, (cudaStream_t)stream);
}

int ts_bfs_expand(const ts_bfs_args* a, void* stream) {
    if (int rc = bfs_check(a, true)) return rc;
    if (a->n_items == 0) return 0;   // an empty local frontier is normal on a multi-rank search
    if (!a->d_walls || !a->d_targets_packed || !a->d_in_keys || !a->d_out_keys) return TS_E_NULL_POINTER;
    return (int)bfs_dispatch(1, *a, (cudaStream_t)stream);
}

int ts_bfs_expand_exchange(const ts_bfs_args* a, void* stream) {
    if (int rc = bfs_check(a, true)) return rc;
    if (a->n_ranks < 1 || a->n_ranks > 64 || a->inbox_capacity < 1 || (a->parity != 0 && a->parity != 1)) return TS_E_BAD_ARGUMENT;
    if (a->n_items == 0) return 0;
    if (!a->d_walls || !a->d_targets_packed || !a->d_in_keys || !a->d_counts || !a->d_peer_bufs) return TS_E_NULL_POINTER;
    return (int)bfs_dispatch(3, *a, (cudaStream_t)stream);
}

int ts_bfs_levels(const ts_bfs_args* a, int32_t first_depth, int32_t n_levels, uint64_t* front0, uint64_t* front1,
                  uint64_t* succ, int64_t* lvl, int64_t frontier_capacity, int64_t known_frontier, void* stream) {
    if (int rc = bfs_check(a, true)) return rc;
    if (!front0 || !front1 || !succ || !lvl || !a->d_table || !a->d_walls || !a->d_targets_packed) return TS_E_NULL_POINTER;
    if (first_depth < 0 || n_levels < 0 || frontier_capacity < 1) return TS_E_BAD_ARGUMENT;
    if (known_frontier >= 0 && (n_levels != 1 || known_frontier > frontier_capacity)) return TS_E_BAD_ARGUMENT;
    const bool exact = known_frontier >= 0;
    if (a->table_capacity < 2 || (a->table_capacity & (a->table_capacity - 1))) return TS_E_BAD_ARGUMENT;
    for (int32_t d = first_depth; d < first_depth + n_levels; ++d) {
        uint64_t* in = (d & 1) ? front1 : front0;
        uint64_t* out = (d & 1) ? front0 : front1;
        ts_bfs_args e = *a;                       // K4: frontier -> 4 successors each
        e.n_items = exact ? known_frontier : frontier_capacity;
        e.d_n_items = exact ? nullptr : lvl + 4 * (int64_t)d;
        e.n_items_scale = 1;
        e.d_in_keys = in;
        e.d_out_keys = succ;
        if (int rc = ts_bfs_expand(&e, stream)) return rc;
        ts_bfs_args k = *a;                       // K5: successors -> table, new keys = next frontier
        k.n_items = 4 * (exact ? known_frontier : frontier_capacity);
        k.d_n_items = exact ? nullptr : lvl + 4 * (int64_t)d;
        k.n_items_scale = 4;
        k.d_in_keys = succ;
        k.d_out_keys = out;
        k.out_capacity = frontier_capacity;
        k.d_counts = reinterpret_cast<uint64_t*>(lvl + 4 * (int64_t)(d + 1));
        k.d_parent_keys = a->d_table_parent ? in : nullptr;
        k.depth = d + 1;
        k.d_won_keys = nullptr;
        if (int rc = ts_bfs_hash_insert(&k, stream)) return rc;
    }
    return 0;
}

int ts_bfs_traceback(const ts_bfs_args* a, void* stream) {
    if (int rc = bfs_check(a, true)) return rc;
    if (a->n_items == 0) return 0;
    if (!a->d_walls || !a->d_in_keys || !a->d_parent_keys || !a->d_table || !a->d_table_parent || !a->d_moves || !a->d_lengths) return TS_E_NULL_POINTER;
    if (a->table_capacity < 2 || (a->table_capacity & (a->table_capacity - 1)) || a->max_moves < 1) return TS_E_BAD_ARGUMENT;
    return (int)bfs_dispatch(2, *a, (cudaStream_t)stream);
}

int ts_bfs_trace_step(const ts_bfs_args* a, void* stream) {
    if (int rc = bfs_check(a, true)) return rc;
    if (a->n_items == 0) return 0;
    if (!a->d_walls || !a->d_in_keys || !a->d_out_keys || !a->d_table || !a->d_table_parent || !a->d_moves) return TS_E_NULL_POINTER;
    if (a->table_capacity < 2 || (a->table_capacity & (a->table_capacity - 1))) return TS_E_BAD_ARGUMENT;
    if (a->n_ranks < 1 || a->n_ranks > 64 || a->rank < 0 || a->rank >= a->n_ranks) return TS_E_BAD_ARGUMENT;
    return (int)bfs_dispatch(4, *a, (cudaStream_t)stream);
}

int ts_bfs_partition_count(const ts_bfs_args* a, void* stream) {
    if (int rc = bfs_check(a, false)) return rc;
    if (a->n_ranks < 1 || a->n_ranks > 64) return TS_E_BAD_ARGUMENT;
    if (a->n_items == 0) return 0;
    if (!a->d_in_keys || !a->d_counts) return TS_E_NULL_POINTER;
    bfs_partition_count_kernel<<<(unsigned)((a->n_items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

int ts_bfs_partition_scatter(const ts_bfs_args* a, void* stream) {
    if (int rc = bfs_check(a, false)) return rc;
    if (a->n_ranks < 1 || a->n_ranks > 64) return TS_E_BAD_ARGUMENT;
    if (a->n_items == 0) return 0;
    if (!a->d_in_keys || !a->d_counts) return TS_E_NULL_POINTER;   // d_out_keys may be NULL when every key is NONE
    if (a->d_out_parents && !a->d_parent_keys) return TS_E_NULL_POINTER;
    bfs_partition_scatter_kernel<<<(unsigned)((a->n_items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

int ts_bfs_hash_insert(const ts_bfs_args* a, void* stream) {
    if (int rc = bfs_check(a, false)) return rc;
    if (a->table_capacity < 2 || (a->table_capacity & (a->table_capacity - 1))) return TS_E_BAD_ARGUMENT;
    if (a->n_items == 0) return 0;
    if (!a->d_in_keys || !a->d_counts || !a->d_out_keys || !a->d_table) return TS_E_NULL_POINTER;
    unsigned blocks = (unsigned)((a->n_items + 255) / 256);
    if (a->d_n_items && blocks > BFS_PERSISTENT_BLOCKS) blocks = BFS_PERSISTENT_BLOCKS;
    bfs_hash_insert_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

}  // extern "C"
