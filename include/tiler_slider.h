/*
 * tiler_slider.h -- C-ABI of libtiler_slider.so, the B200 (sm_100a) batched Tiler-Slider
 * step path.
 *
 * The reference (AnimeshSinha1309/tiler-slider) is pure Python and has no FFI boundary; the
 * path sits behind two classes, GameState (explainrl/environment/state.py:18-222) and
 * TilerSliderEnv (explainrl/environment/environment.py:14-194).  Each entry point below
 * names the reference code whose results it reproduces; INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *  - Every pointer named d_* (or inside a ts_*_args struct, unless marked host) is a DEVICE
 *    pointer into caller-owned memory (PyTorch tensors in this repo).  The library
 *    allocates nothing persistent except inside ts_host_ctx.
 *  - Calls are asynchronous on the given cudaStream_t (void*; 0 = legacy default stream),
 *    except ts_step_host which synchronises before it returns.
 *  - Return value: 0 = success; negative = argument error (TS_E_*); positive = cudaError_t
 *    from the launch.  ts_last_error_string() describes the last failure of the calling
 *    thread.  No global mutable state: re-entrant across streams and devices (the caller
 *    selects the device).
 *  - There is no CPU fallback anywhere in this library.
 *  - Domain: well-formed boards -- tiles on distinct cells, none on a blocked cell (what the
 *    reference's loaders produce; outside it the reference itself is erratic).  The
 *    kernels do not check this (tiler_slider_b200.Puzzle.validate does, on the host); a tile placed
 *    on a blocked cell slides out of it on boards up to 8x8 and 15x15 / 16x16, and stays put on
 *    9x9 .. 14x14 (the wide slide treats the own cell as wall-free).  GameState.move_to, which is
 *    defined for blocked start cells, therefore probes on a board whose start cell is open.
 *
 * Packed layout (struct-of-arrays, environment index innermost):
 *  - capacity: allocation stride in environments, a multiple of TS_CAP_ALIGN.  EVERY per-env
 *    array handed to the library (state, actions, reward, done, flags, ...) must hold
 *    `capacity` elements.  ts_step reads and writes exactly the envs of its range
 *    [first_env, first_env + n_envs): whole 4-env groups go through the vectorised kernel, the
 *    ragged end (n_envs % 4 envs) through a per-env kernel.  The pure queries ts_valid_moves and
 *    ts_goal_check may also WRITE their (correct) result for the up-to-3 envs that share the last
 *    4-env group of the range.
 *  - position word: ts_pos_bytes(T) in {1,2,4,8,16,32} bytes per env, byte i = row*PS + col of tile
 *    i (PS = ts_pos_stride(S)), unused bytes zero.  Arrays: pos (in/out), init, targets
 *    (ordered mode).
 *  - board: bitboard of ts_board_bytes(S) bytes per env, bit row*BS + col (BS =
 *    ts_board_stride(S)) set = blocked (walls) or target cell (targets, set mode), little
 *    endian, split into byte planes of width 16 (repeated), 8, 4, 2, 1 -- widest first; plane
 *    k of a buffer starts at byte offset ts_plane_offset(nb,k)*capacity and holds one element
 *    per env.  S <= 6: BS = PS = S+1 and a WALL board also has column S of every row and all
 *    bits past the last row set (sentinels that end a slide; ts_encode / ts_synth write
 *    them).  S = 7, 8: BS = S, PS = 16, no sentinels.
 *  - wide boards (S >= 9): walls = u16 [axis][capacity][L lines], L = S rounded up to even:
 *    plane 1 = rows (line r, one bit per column), used by LEFT/RIGHT, plane 0 = columns (line
 *    c, one bit per row), used by UP/DOWN; ts_walls_bytes(S) = 4*L bytes per env, of which a
 *    step reads the 2*L bytes of one axis.  S <= 14: cell k of a
 *    line is bit k+1, and bit 0 and bit S+1 of every line are set (edge sentinels); S = 15, 16:
 *    cell k is bit k, no sentinel.  Set-goal target record = u16 [capacity][17]: 16 rows, cell
 *    (r,c) = bit c of word r, then the number of distinct target cells (the step kernels test
 *    "every tile on a target cell" plus "that number == n_tiles", which is set equality).  PS = 16.
 */
#ifndef TILER_SLIDER_H
#define TILER_SLIDER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TS_VERSION 201          /* 0.2.1 */
#define TS_CAP_ALIGN 128
#define TS_MAX_SIZE 16
#define TS_MAX_TILES 32         /* 0..8: register kernels; 9..32: per-env generic kernels (csrc/ts_generic.cu) */

/* flag bits of the per-env status byte */
#define TS_F_DONE 1u      /* environment.py:133-141  done = is_won or step_count >= max_steps */
#define TS_F_WON 2u       /* info['is_won'] / info['success']  (state.py:172-186)            */
#define TS_F_INVALID 4u   /* info['invalid_move']: no tile moved (environment.py:129)         */
#define TS_F_TIMEOUT 8u   /* info['timeout'] (environment.py:139-141)                         */
#define TS_F_STALE 16u    /* stepped while already done without auto_reset: frozen, no-op
                             (the reference raises RuntimeError, environment.py:113-114)      */

/* goal modes */
#define TS_GOAL_ORDERED 0 /* multi_color=True : tile i on target i (state.py:183-184); targets
                             are packed position words                                        */
#define TS_GOAL_SET 1     /* multi_color=False: occupancy == target set (state.py:185-186);
                             targets are a bitboard in the plane layout                       */

/* error codes */
#define TS_E_BAD_SIZE (-1)
#define TS_E_BAD_TILES (-2)
#define TS_E_BAD_CAPACITY (-3)
#define TS_E_NULL_POINTER (-4)
#define TS_E_MISALIGNED (-5)
#define TS_E_BAD_RANGE (-6)
#define TS_E_UNSUPPORTED (-7)
#define TS_E_BAD_ARGUMENT (-8)

int ts_version(void);
const char *ts_last_error_string(void);

/* layout queries (pure host arithmetic) */
int ts_pos_bytes(int n_tiles);   /* n_tiles = 0 (a board without tiles): 1, the byte is zero */
int ts_board_bytes(int size);
int ts_board_stride(int size);   /* BS: bit index of cell (r,c) in a board = r*BS + c */
int ts_pos_stride(int size);     /* PS: position byte of a tile at (r,c) = r*PS + c */
int ts_walls_bytes(int size);    /* bytes per env of the walls buffer (its size = this * capacity) */
int ts_target_board_bytes(int size); /* bytes per env of the set-goal target board buffer */
int ts_plane_count(int n_bytes);
int ts_plane_width(int n_bytes, int k);
int ts_plane_offset(int n_bytes, int k);
/* 1 if ts_step covers (size, n_tiles): 1 <= size <= 16, 0 <= n_tiles <= 32, else 0 */
int ts_supported(int size, int n_tiles);

/* ---------------------------------------------------------------------------------------
 * K1 ts_encode: dense per-env description -> packed layout.
 * Reproduces GameState.__init__ (state.py:61-73): is_blocked grid + location lists.
 *   d_blocked  u8 [n_envs][size*size]  nonzero = blocked cell
 *   d_tiles    u8 [n_envs][n_tiles][2] (row, col)
 *   d_targets  u8 [n_envs][n_targets][2]
 * Writes walls planes, init + pos position words, targets (ordered: position word of the first
 * min(n_targets, n_tiles) targets -- with n_targets != n_tiles the goal can never be met
 * (state.py:183-184) and the caller passes never_win = 1 to ts_step / ts_goal_check; set:
 * bitboard planes, any n_targets).  n_tiles may be 0 (d_tiles is then not read).
 * Environments are written at [first_env, first_env + n_envs).
 * ------------------------------------------------------------------------------------- */
typedef struct ts_encode_args {
    int32_t size, n_tiles, n_targets, goal_mode;
    int64_t first_env, n_envs, capacity;
    const uint8_t *d_blocked, *d_tiles, *d_targets;
    uint8_t *d_walls, *d_targets_packed, *d_init, *d_pos;
} ts_encode_args;
int ts_encode(const ts_encode_args *a, void *stream);

/* ---------------------------------------------------------------------------------------
 * K0 ts_synth: synthetic puzzles generated on the device in the packed layout.
 * Same recipe as TilerSliderEnvFactory.create_simple_env (environment.py:221-226): a
 * uniformly random arrangement of distinct cells, the first n_walls blocked, the next
 * n_tiles tiles, the next n_tiles targets (always well formed).  The random stream is a
 * counter-based hash of (seed, global env index = env_index_base + local index), NOT
 * numpy's Mersenne Twister, so puzzles differ from create_simple_env(seed=...) but are
 * independent of how environments are sharded across GPUs.
 * ------------------------------------------------------------------------------------- */
typedef struct ts_synth_args {
    int32_t size, n_tiles, n_walls, goal_mode;
    int64_t first_env, n_envs, capacity, env_index_base;
    uint64_t seed;
    uint8_t *d_walls, *d_targets_packed, *d_init, *d_pos;
} ts_synth_args;
int ts_synth(const ts_synth_args *a, void *stream);

/* ---------------------------------------------------------------------------------------
 * K2 ts_step: one environment step for envs [first_env, first_env+n_envs) (first_env a
 * multiple of 4).  Fuses GameState.move (state.py:120-170), is_won (state.py:172-186) and
 * the bookkeeping of TilerSliderEnv.step (environment.py:119-143): invalid_move, win ->
 * done, step counter, timeout; plus (new in this repo) reward and optional auto-reset
 * (TilerSliderEnv.reset, environment.py:89-97: positions <- init, step_count <- 0).
 *
 *   reward  = won ? r_win : invalid_move ? r_invalid : r_step      (stale envs: 0)
 *   d_flags : TS_F_* bits.  With auto_reset = 0 it is read first: an env whose DONE bit is
 *             set is frozen and reports DONE|STALE.  With auto_reset = 1 it is write-only
 *             and may be NULL.
 *   d_done  : 0/1 byte per env (may be NULL if d_flags is given).
 *   d_terminal_pos : optional position words, written for a 4-env group whenever one of its
 *             envs finished this step: the positions after the move, before the reset.
 *   never_win: 1 when the goal can never be met (ordered mode with len(targets) !=
 *             len(tiles), state.py:183-184).
 *   n_tiles = 0: nothing moves (every step is INVALID); won iff the board has no targets either
 *             (ordered mode: the caller says so through never_win).
 *   count_bytes = 1 needs max_steps <= 255 with auto_reset, <= 254 without.
 * ------------------------------------------------------------------------------------- */
typedef struct ts_step_args {
    int32_t size, n_tiles, goal_mode, never_win;
    int64_t first_env, n_envs, capacity;
    const uint8_t *d_walls, *d_targets_packed, *d_init;
    uint8_t *d_pos;
    void *d_step_count;
    int32_t count_bytes, max_steps, auto_reset, reserved0;
    const uint8_t *d_actions;
    float r_win, r_step, r_invalid, reserved1;
    float *d_reward;
    uint8_t *d_done, *d_flags, *d_terminal_pos;
} ts_step_args;
int ts_step(const ts_step_args *a, void *stream);

/* ---------------------------------------------------------------------------------------
 * K3 ts_observe: dense observation, GameState.get_state_array (state.py:188-211).
 *   d_obs f32 [n_envs][size][size][3] (HWC): ch0 blocked, ch1 tile index+1 (ordered mode) or
 *   1 (set mode), ch2 target index+1 or 1.  In set mode targets come from the bitboard.
 *   n_targets (ordered mode): 0 = as many targets as tiles, d_targets_packed as ts_step reads it;
 *   k > 0 = k targets in words of ts_pos_bytes(k) bytes per env (a board whose target count
 *   differs from its tile count, which ts_step's packed word cannot hold); -1 = no targets.
 * ------------------------------------------------------------------------------------- */
typedef struct ts_observe_args {
    int32_t size, n_tiles, goal_mode, n_targets;
    int64_t first_env, n_envs, capacity;
    const uint8_t *d_walls, *d_targets_packed, *d_pos;
    float *d_obs;
} ts_observe_args;
int ts_observe(const ts_observe_args *a, void *stream);

/* ---------------------------------------------------------------------------------------
 * ts_valid_moves: bit d of d_mask[env] set when move d changes any tile position
 * (TilerSliderEnv.get_valid_moves, environment.py:149-171).
 * ------------------------------------------------------------------------------------- */
typedef struct ts_valid_args {
    int32_t size, n_tiles;
    int64_t first_env, n_envs, capacity;
    const uint8_t *d_walls, *d_pos;
    uint8_t *d_mask;
} ts_valid_args;
int ts_valid_moves(const ts_valid_args *a, void *stream);

/* ---------------------------------------------------------------------------------------
 * ts_goal_check: d_won[env] = 1 when the CURRENT positions meet the goal
 * (GameState.is_won, state.py:172-186), without moving.
 * ------------------------------------------------------------------------------------- */
typedef struct ts_goal_args {
    int32_t size, n_tiles, goal_mode, never_win;
    int64_t first_env, n_envs, capacity;
    const uint8_t *d_targets_packed, *d_pos;
    uint8_t *d_won;
} ts_goal_args;
int ts_goal_check(const ts_goal_args *a, void *stream);

/* ---------------------------------------------------------------------------------------
 * Breadth-first state-space expansion (BASELINE config 5; no reference counterpart -- the
 * successor function is GameState.move, state.py:120-170, the goal test is_won,
 * state.py:172-186; boards of size <= 8).
 *
 * State key (uint64): canonical position word in the low bits (ordered goal: the position
 * bytes; set goal: the bytes sorted ascending), puzzle id in bits 32..62 when n_tiles <= 4
 * (n_tiles > 4: one puzzle only), bit 63 = "meets the goal" (ignored by dedup).
 * TS_BFS_NONE marks a move that changed nothing.  The puzzle table is a packed batch
 * (walls / targets / init of `puzzle_capacity` stride, as written by ts_encode).
 *
 *   ts_bfs_seed              d_out_keys[i] = key of puzzle i's initial state, i < n_items
 *   ts_bfs_expand        K4  d_out_keys[4*i+d] = successor of d_in_keys[i] under move d
 *   ts_bfs_partition_count   d_counts[r] += #keys of d_in_keys owned by rank r = hash(key) % n_ranks
 *   ts_bfs_partition_scatter d_out_keys bucketed by owner; d_counts[r] = write cursor of bucket r
 *                            (initialise to the exclusive prefix sum of the counts)
 *   ts_bfs_hash_insert   K5  insert d_in_keys into the open-addressing table d_table
 *                            (table_capacity a power of two, empty = TS_BFS_NONE); keys not
 *                            seen before are appended to d_out_keys; d_counts[0] = append
 *                            cursor, d_counts[1] += #goal successors seen, d_counts[2] = 1 on
 *                            overflow of the table or of out_capacity
 *   ts_bfs_traceback         follow the recorded parents from a goal state back to the root and
 *                            emit the move string (the move between parent and child is found by
 *                            re-running the four slides on the parent)
 * The exchange between partition and insert is an NCCL all-to-all done by the caller.
 * ------------------------------------------------------------------------------------- */
#define TS_BFS_NONE 0xFFFFFFFFFFFFFFFFull
#define TS_BFS_WON_BIT 0x8000000000000000ull
typedef struct ts_bfs_args {
    int32_t size, n_tiles, goal_mode, never_win, n_ranks;
    int32_t rank;                       /* ts_bfs_trace_step: the caller's rank among n_ranks owners */
    int64_t n_items, puzzle_capacity, table_capacity, out_capacity;
    const uint8_t *d_walls, *d_targets_packed, *d_init;
    const uint64_t *d_in_keys;
    uint64_t *d_out_keys, *d_table, *d_counts;
    /* optional parent tracking: with d_table_parent given, ts_bfs_hash_insert records for every
     * new key the key it was expanded from; d_parent_keys = NULL marks roots.
     *   parent_per_item = 0 (single rank): d_in_keys must be exactly the output of ts_bfs_expand
     *     on d_parent_keys, so that the parent of d_in_keys[i] is d_parent_keys[i / 4];
     *   parent_per_item = 1 (several ranks): the parent of d_in_keys[i] is d_parent_keys[i] --
     *     the array ts_bfs_partition_scatter filled next to the keys (d_out_parents) and the
     *     caller exchanged with them. */
    const uint64_t *d_parent_keys;
    uint64_t *d_table_parent;
    /* ts_bfs_traceback: d_in_keys[i] = goal state of item i (TS_BFS_NONE: none), d_parent_keys[i]
     * = the state it was generated from (d_goal_parents of the inserts); writes the shortest move
     * string of item i to d_moves[i * max_moves ...] (0..3, root first) and its length to
     * d_lengths[i] (-1: no goal / longer than max_moves / broken chain) */
    uint8_t *d_moves;
    int32_t *d_lengths;
    int64_t max_moves;
    /* ts_bfs_hash_insert, optional: the first won_capacity goal successors seen (duplicates
     * included) are also copied to d_won_keys, in the order of d_counts[1] */
    uint64_t *d_won_keys;
    int64_t won_capacity;
    /* ts_bfs_hash_insert, optional per-puzzle tallies (puzzle id = key bits 32..62; 0 when
     * n_tiles > 4): d_states_per_puzzle[pid] += 1 for every new key; for every goal successor
     * d_solve_depth[pid] = min(itself, depth), and the first one to lower it also stores its key
     * in d_goal_keys[pid] (optional) -- a goal state at the smallest depth */
    int64_t *d_states_per_puzzle;
    int32_t *d_solve_depth;
    uint64_t *d_goal_keys;
    int32_t depth, reserved2;
    /* ts_bfs_expand_exchange (multi-GPU, peer memory): d_peer_bufs[r] = rank r's exchange buffer
     * as mapped into THIS process (NVLink peer mapping, e.g. torch symmetric memory):
     *   word 0, 1          arrival cursors of inbox 0 / inbox 1 (keys received so far)
     *   word 2             set to 1 when an inbox overflowed
     *   word TS_BFS_XHDR + p * inbox_capacity ...   inbox p (u64 keys)
     * The successors of d_in_keys are written straight into inbox `parity` of their owner rank
     * (hash(key) % n_ranks); d_counts[3] += number of keys sent. */
    uint64_t *const *d_peer_bufs;
    int64_t inbox_capacity;
    int32_t parity, parent_per_item;
    /* device-driven levels (ts_bfs_expand, ts_bfs_hash_insert), optional: the item count is read
     * on the DEVICE as min(n_items, *d_n_items * n_items_scale) -- n_items is then only the
     * bound the grid is sized for -- so consecutive levels can be launched back to back
     * without the host learning the frontier sizes in between: expand level d with
     * d_n_items = &new_count[d], insert with the same pointer, scale 4 and d_counts =
     * the counter block of level d+1. */
    const int64_t *d_n_items;
    int64_t n_items_scale;
    /* ts_bfs_partition_scatter, optional (with d_parent_keys = the frontier d_in_keys was expanded
     * from): d_out_parents[j] = parent of the key written to d_out_keys[j] */
    uint64_t *d_out_parents;
    /* ts_bfs_hash_insert, optional (with d_goal_keys and d_parent_keys): d_goal_parents[pid] = the
     * state the goal successor stored in d_goal_keys[pid] was generated from.  A shortest solution
     * is the chain root -> that state plus the move from it to the goal; the goal state's own
     * table entry must not be used for this, because a puzzle whose initial state already meets
     * the goal is won by its first step (a root is never a goal successor), and the root has no
     * parent. */
    uint64_t *d_goal_parents;
} ts_bfs_args;
#define TS_BFS_XHDR 16
int ts_bfs_seed(const ts_bfs_args *a, void *stream);
int ts_bfs_expand(const ts_bfs_args *a, void *stream);
int ts_bfs_partition_count(const ts_bfs_args *a, void *stream);
int ts_bfs_partition_scatter(const ts_bfs_args *a, void *stream);
int ts_bfs_hash_insert(const ts_bfs_args *a, void *stream);
/* expand + bucket + exchange in one kernel: no partition pass, no size exchange, no all-to-all;
 * the caller separates levels with any collective that orders the ranks' streams (an all-reduce
 * of the sent counts doubles as the termination test) and alternates `parity` level by level. */
int ts_bfs_expand_exchange(const ts_bfs_args *a, void *stream);
int ts_bfs_traceback(const ts_bfs_args *a, void *stream);
/* One step of a traceback over a visited set that is spread over n_ranks owners (hash-partitioned
 * search with parents): for every item i whose key d_in_keys[i] this rank owns, look the key up and
 * write its parent to d_out_keys[i] and the move that leads from the parent to it to d_moves[i]
 * (one byte per item).  Read as int64: parent >= 0; -1 (TS_BFS_NONE) the key is a root; -2 the chain
 * is broken (key or move not found); INT64_MIN (with move 0) for TS_BFS_NONE items and for keys of
 * other ranks -- so a MAX all-reduce of both arrays over the owners yields the step for every
 * item.  n_ranks = 1: every key is this rank's.  With d_parent_keys given (the last link of a
 * solution, d_goal_parents) nothing is looked up: the parent of item i is d_parent_keys[i], on
 * every rank. */
int ts_bfs_trace_step(const ts_bfs_args *a, void *stream);
/* n_levels device-driven BFS levels launched back to back (2 kernels per level, no host sync):
 * level d = first_depth .. first_depth + n_levels - 1 expands the lvl[4*d] keys of
 * front[d & 1] into succ (4 * frontier_capacity keys) and inserts them as level d+1 into
 * front[(d+1) & 1], counters in lvl[4*(d+1) ..] = (new keys, goal successors, overflow, -),
 * which must be zero on entry.  `a` supplies the puzzle tables, d_table / table_capacity, the
 * optional parent table (d_table_parent; parents are the frontier keys) and per-puzzle tallies.
 * known_frontier >= 0: the host knows the size of level first_depth (n_levels must be 1): that
 * level is launched with a grid of its own size instead of one persistent wave. */
int ts_bfs_levels(const ts_bfs_args *a, int32_t first_depth, int32_t n_levels, uint64_t *front0, uint64_t *front1,
                  uint64_t *succ, int64_t *lvl, int64_t frontier_capacity, int64_t known_frontier, void *stream);

/* ---------------------------------------------------------------------------------------
 * K6 ts_bfs_local: breadth-first search of many small puzzles, one CTA per puzzle, on chip
 * (csrc/ts_bfs_local.cu).  Same successor function, goal test, canonical states and result
 * definitions as the hash-partitioned search above; boards of size <= 8 with 1..4 tiles.
 * The visited set of a puzzle is a bitmap in shared memory over a perfect hash of the state (the
 * arrangement number of its T tiles among the F free cells of the puzzle: F!/(F-T)! bits,
 * bitmap_words u32 words, a multiple of 4); the frontier queue is a ring of `queue_smem` states
 * (a power of two >= 32) in shared memory -- the level being expanded and the one being appended --
 * and spills to d_spill[cta][spill_per_cta], indexed by discovery order.  Persistent grid: CTAs
 * draw puzzles from the ticket counter.
 *   d_puzzle_ids       optional: search puzzles d_puzzle_ids[0 .. n_puzzles) instead of 0 .. n_puzzles-1
 *   d_status[pid]      0 searched; 1 state space exceeds bitmap_words*32 bits; 2 spill slab exhausted; 3 deeper
 *                      than 255 levels -- for != 0 nothing else of the puzzle is valid (search it
 *                      with the hash-partitioned path)
 *   d_states_per_puzzle[pid], d_solve_depth[pid] (-1: no goal within max_depth)
 *   d_levels[d]        += new states at depth d over the searched puzzles (n_levels >= 256 entries)
 *   d_counters         [0] ticket (zero on entry), [1] += successors generated, [3] = max(depth reached)
 *   d_lengths / d_moves / d_parent_scratch  optional shortest move string per puzzle (0..3, root
 *                      first; length -1: unsolved or longer than max_moves); d_parent_scratch holds
 *                      grid * spill_per_cta words (spill_per_cta then bounds the states of a puzzle)
 * ts_bfs_local_smem_bytes: dynamic shared memory of a launch with these arguments;
 * ts_bfs_local_ctas_per_sm: resident CTAs per SM and the SM count, to size the grid.
 * ------------------------------------------------------------------------------------- */
typedef struct ts_bfs_local_args {
    int32_t size, n_tiles, goal_mode, never_win;
    int64_t n_puzzles, puzzle_capacity;
    const uint8_t *d_walls, *d_targets_packed, *d_init;
    const int32_t *d_puzzle_ids;
    int32_t max_depth, bitmap_words, queue_smem, n_levels;
    uint32_t *d_spill;
    int64_t spill_per_cta;
    uint32_t *d_parent_scratch;
    int64_t *d_states_per_puzzle;
    int32_t *d_solve_depth, *d_status;
    int64_t *d_levels;
    uint64_t *d_counters;
    uint8_t *d_moves;
    int32_t *d_lengths;
    int64_t max_moves;
} ts_bfs_local_args;
int ts_bfs_local_smem_bytes(const ts_bfs_local_args *a);
int ts_bfs_local_ctas_per_sm(const ts_bfs_local_args *a, int *ctas_per_sm, int *n_sm);
int ts_bfs_local(const ts_bfs_local_args *a, int grid, void *stream);

/* ---------------------------------------------------------------------------------------
 * ts_step_host: the same step through HOST buffers (the call a host-side driver makes):
 * h_actions (pinned) -> device, ts_step, reward/done -> h_reward/h_done (pinned), pipelined in
 * chunks over the context's streams; returns after everything has landed.
 * a->d_actions/d_reward/d_done must still point at device staging of n_envs elements.
 * h_flags (optional, pinned): also download the TS_F_* status byte.  With h_reward = h_done =
 * NULL only that byte comes back (1 instead of 5 bytes per env over PCIe): it encodes done (bit
 * 0) and, since reward is a function of WON / INVALID, the reward.
 * ------------------------------------------------------------------------------------- */
typedef struct ts_host_ctx ts_host_ctx;
int ts_host_ctx_create(ts_host_ctx **out, int n_streams);
int ts_host_ctx_destroy(ts_host_ctx *ctx);
int ts_step_host(ts_host_ctx *ctx, const ts_step_args *a, const uint8_t *h_actions, float *h_reward,
                 uint8_t *h_done, uint8_t *h_flags, int64_t chunk_envs);

#ifdef __cplusplus
}
#endif
#endif /* TILER_SLIDER_H */
