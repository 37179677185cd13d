// ts_common.cuh -- packed-state layout rules and the closed-form slide core shared by every
// sm_100a kernel of the Tiler-Slider step path.
//
// Results reproduced (reference paths relative to the reference checkout):
//   explainrl/environment/state.py:120-170  GameState.move   -> slide_env<S,T>()
//   explainrl/environment/state.py:172-186  GameState.is_won -> goal compare in the callers
//
// Layout in HBM (struct-of-arrays, environment index innermost):
//   position word   one word of POS_BYTES(T) in {1,2,4,8} bytes per env; byte i = (row<<4)|col of
//                   tile i, unused bytes are zero
//   board planes    a bitboard of ceil(S*S/8) bytes per env, bit r*S+c, little endian, split
//                   into byte planes of width 16/8/4/2/1 (widest first); plane k starts at byte
//                   offset plane_offset(k)*capacity of the buffer and is indexed by env
//   capacity        allocation stride in envs, a multiple of 128 so that every plane start and
//                   every 4-env group is 16-byte aligned
// A thread owns GROUP=4 consecutive envs, so every stream is read with one 32/64/128-bit
// load per plane and a warp touches whole 128-byte lines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace ts {

constexpr int GROUP = 4;           // envs per thread
constexpr int CAP_ALIGN = 128;     // capacity granularity (envs)
constexpr int MAX_SIZE = 16;
constexpr int MAX_TILES = 8;

__host__ __device__ constexpr int pos_bytes(int T) { return T <= 1 ? 1 : T <= 2 ? 2 : T <= 4 ? 4 : 8; }
__host__ __device__ constexpr int board_bytes(int S) { return (S * S + 7) / 8; }

// decomposition of nb bytes into planes of width 16 (repeated), 8, 4, 2, 1
__host__ __device__ constexpr int plane_count(int nb) {
    int k = nb / 16, rem = nb % 16;
    for (int w = 8; w >= 1; w >>= 1)
        if (rem >= w) { rem -= w; ++k; }
    return k;
}
__host__ __device__ constexpr int plane_width(int nb, int i) {
    int k = 0, rem = nb;
    while (rem >= 16) { if (k == i) return 16; rem -= 16; ++k; }
    for (int w = 8; w >= 1; w >>= 1)
        if (rem >= w) { if (k == i) return w; rem -= w; ++k; }
    return 0;
}
__host__ __device__ constexpr int plane_offset(int nb, int i) {
    int off = 0;
    for (int k = 0; k < i; ++k) off += plane_width(nb, k);
    return off;
}

// ---- compile-time loop -----------------------------------------------------------------
template <int I> struct IC { static constexpr int value = I; };
template <int B, int E, class F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) { f(IC<B>{}); static_for<B + 1, E>(f); }
}

// ---- streaming global access ---------------------------------------------------------------
// Every stream is touched once per step and the working set (>=400 MB at 16M envs) exceeds
// the 126 MB L2, so loads/stores carry the evict-first (.cs) policy.
template <int WORDS> __device__ __forceinline__ void ld_words(const void* p, uint32_t (&r)[WORDS]) {
    if constexpr (WORDS == 1) {
        r[0] = __ldcs(reinterpret_cast<const unsigned int*>(p));
    } else if constexpr (WORDS == 2) {
        uint2 v = __ldcs(reinterpret_cast<const uint2*>(p));
        r[0] = v.x; r[1] = v.y;
    } else {
        static_assert(WORDS % 4 == 0, "group loads are 4, 8 or 16*k bytes");
#pragma unroll
        for (int k = 0; k < WORDS / 4; ++k) {
            uint4 v = __ldcs(reinterpret_cast<const uint4*>(p) + k);
            r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
        }
    }
}
template <int WORDS> __device__ __forceinline__ void st_words(void* p, const uint32_t (&r)[WORDS]) {
    if constexpr (WORDS == 1) {
        __stcs(reinterpret_cast<unsigned int*>(p), r[0]);
    } else if constexpr (WORDS == 2) {
        __stcs(reinterpret_cast<uint2*>(p), make_uint2(r[0], r[1]));
    } else {
#pragma unroll
        for (int k = 0; k < WORDS / 4; ++k)
            __stcs(reinterpret_cast<uint4*>(p) + k, make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]));
    }
}

// element e (0..3) of a 4-env group of W-byte elements held in raw[W] words -> out words
template <int W> __device__ __forceinline__ void group_elem(const uint32_t (&raw)[W], int e, uint32_t (&out)[(W + 3) / 4]) {
    if constexpr (W == 1) out[0] = (raw[0] >> (8 * e)) & 0xFFu;
    else if constexpr (W == 2) out[0] = (raw[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
    else {
#pragma unroll
        for (int k = 0; k < W / 4; ++k) out[k] = raw[(W / 4) * e + k];
    }
}
template <int W> __device__ __forceinline__ void group_set(uint32_t (&raw)[W], int e, const uint32_t (&in)[(W + 3) / 4]) {
    if constexpr (W == 1) raw[0] = (raw[0] & ~(0xFFu << (8 * e))) | ((in[0] & 0xFFu) << (8 * e));
    else if constexpr (W == 2) raw[e >> 1] = (raw[e >> 1] & ~(0xFFFFu << (16 * (e & 1)))) | ((in[0] & 0xFFFFu) << (16 * (e & 1)));
    else {
#pragma unroll
        for (int k = 0; k < W / 4; ++k) raw[(W / 4) * e + k] = in[k];
    }
}

// ---- a 4-env group of bitboards held as its byte planes ------------------------------------
template <int NB> struct BoardGroup {
    static constexpr int NP = plane_count(NB);
    static constexpr int NWORDS = (NB + 3) / 4;
    uint32_t raw[NB];  // plane k occupies raw[plane_offset(k) .. +plane_width(k))

    __device__ __forceinline__ void load(const uint8_t* base, size_t capacity, size_t group) {
        static_for<0, NP>([&](auto I) {
            constexpr int k = decltype(I)::value;
            constexpr int w = plane_width(NB, k), off = plane_offset(NB, k);
            uint32_t tmp[w];
            ld_words<w>(base + (size_t)off * capacity + group * (size_t)(GROUP * w), tmp);
#pragma unroll
            for (int j = 0; j < w; ++j) raw[off + j] = tmp[j];
        });
    }
    __device__ __forceinline__ void store(uint8_t* base, size_t capacity, size_t group) const {
        static_for<0, NP>([&](auto I) {
            constexpr int k = decltype(I)::value;
            constexpr int w = plane_width(NB, k), off = plane_offset(NB, k);
            uint32_t tmp[w];
#pragma unroll
            for (int j = 0; j < w; ++j) tmp[j] = raw[off + j];
            st_words<w>(base + (size_t)off * capacity + group * (size_t)(GROUP * w), tmp);
        });
    }
    // board words of env e: word j holds bits 32j..32j+31
    __device__ __forceinline__ void get(int e, uint32_t (&bw)[NWORDS]) const {
#pragma unroll
        for (int j = 0; j < NWORDS; ++j) bw[j] = 0;
        static_for<0, NP>([&](auto I) {
            constexpr int k = decltype(I)::value;
            constexpr int w = plane_width(NB, k), off = plane_offset(NB, k);
            uint32_t tmp[w], el[(w + 3) / 4];
#pragma unroll
            for (int j = 0; j < w; ++j) tmp[j] = raw[off + j];
            group_elem<w>(tmp, e, el);
            if constexpr (w >= 4) {
#pragma unroll
                for (int j = 0; j < w / 4; ++j) bw[off / 4 + j] = el[j];
            } else {
                bw[off / 4] |= el[0] << (8 * (off % 4));
            }
        });
    }
    __device__ __forceinline__ void set(int e, const uint32_t (&bw)[NWORDS]) {
        static_for<0, NP>([&](auto I) {
            constexpr int k = decltype(I)::value;
            constexpr int w = plane_width(NB, k), off = plane_offset(NB, k);
            uint32_t tmp[w], el[(w + 3) / 4];
#pragma unroll
            for (int j = 0; j < w; ++j) tmp[j] = raw[off + j];
            if constexpr (w >= 4) {
#pragma unroll
                for (int j = 0; j < w / 4; ++j) el[j] = bw[off / 4 + j];
            } else {
                el[0] = bw[off / 4] >> (8 * (off % 4));
            }
            group_set<w>(tmp, e, el);
#pragma unroll
            for (int j = 0; j < w; ++j) raw[off + j] = tmp[j];
        });
    }
};

// ---- SWAR helpers on 4 packed position bytes -------------------------------------------------
__device__ __forceinline__ uint32_t swap_nibbles(uint32_t x) {
    return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
}
template <int I> __device__ __forceinline__ uint32_t byte_of(uint32_t x) {
    return __byte_perm(x, 0, 0x4440 + I);
}

// =============================================================================================
// slide_env<S,T>: one simultaneous slide of all T tiles of one env (state.py:137-170).
//
// Closed form (checked against the reference by the parity tests): inside every
// maximal wall-free run of a line, tiles keep their order and pack against the run's end in
// the move direction.  For a tile at offset o whose run ends before offset e (nearest wall
// above o, or the edge), the new offset is  o + #empty cells in (o, e).
//
// The four directions share one code path: the board is rotated by 180 degrees for UP/LEFT
// so that every move is "toward higher offsets", positions are nibble-swapped for vertical
// moves so that (line, offset) = (col, row), and the wall bits of a vertical line are
// gathered from the row-major bitboard with one multiply (bits S*k -> B+k, B=(S-1)^2; all
// 36 partial products land on distinct bits, so there are no carries).
// Bitboard variant: S*S <= 64.
// =============================================================================================
__host__ __device__ constexpr uint64_t make_col0(int S) { uint64_t m = 0; for (int k = 0; k < S; ++k) m |= 1ull << (S * k); return m; }
__host__ __device__ constexpr uint64_t make_magic(int S) { uint64_t m = 0; for (int j = 0; j < S; ++j) m |= 1ull << ((S - 1) * (S - 1) - (S - 1) * j); return m; }

template <int S> struct BoardTraits {
    static constexpr int NBITS = S * S;
    static constexpr bool WIDE = NBITS > 32;                 // board needs 64 bits
    static constexpr bool GATHER32 = S * (S - 1) <= 31;      // column gather fits a 32-bit multiply
    using board_t = typename std::conditional<WIDE, uint64_t, uint32_t>::type;
    static constexpr uint64_t COL0 = make_col0(S);           // bits S*k, k < S: column 0
    static constexpr uint64_t MAGIC = make_magic(S);         // moves bit S*k to bit GSHIFT+k
    static constexpr int GSHIFT = (S - 1) * (S - 1);
};

template <int S> __device__ __forceinline__ typename BoardTraits<S>::board_t rot180(typename BoardTraits<S>::board_t w) {
    if constexpr (BoardTraits<S>::WIDE) return __brevll(w) >> (64 - S * S);
    else return __brev(w) >> (32 - S * S);
}

// q[PR]: packed position words ((row<<4)|col per byte).  Updated in place.  walls: row-major
// bitboard.  action: 0 UP, 1 DOWN, 2 LEFT, 3 RIGHT (state.py:31-34).
template <int S, int T>
__device__ __forceinline__ void slide_env(uint32_t (&q)[(T + 3) / 4], typename BoardTraits<S>::board_t walls, uint32_t action) {
    using BT = BoardTraits<S>;
    using board_t = typename BT::board_t;
    constexpr int PR = (T + 3) / 4;
    constexpr uint32_t KFLIP = 0x11111111u * (uint32_t)(S - 1);
    const bool horiz = (action & 2u) != 0;
    const bool flip = (action & 1u) == 0;

    const board_t wb = flip ? rot180<S>(walls) : walls;
    // vertical: line = column -> read the column through the multiply gather
    const uint32_t line_mul = horiz ? (uint32_t)S : 1u;

    uint32_t Q[PR], LSo[PR], LSw[PR], OFF[PR];
    board_t occ = 0;
#pragma unroll
    for (int w = 0; w < PR; ++w) {
        uint32_t x = horiz ? q[w] : swap_nibbles(q[w]);
        x = flip ? KFLIP - x : x;
        Q[w] = x;
        const uint32_t line = (x >> 4) & 0x0F0F0F0Fu;
        OFF[w] = x & 0x0F0F0F0Fu;
        LSo[w] = line * (uint32_t)S;
        LSw[w] = line * line_mul;
    }
    static_for<0, T>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const uint32_t p = byte_of<i % 4>(LSo[i / 4] + OFF[i / 4]);
        occ |= (board_t)1 << p;
    });
    static_for<0, T>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const uint32_t shw = byte_of<i % 4>(LSw[i / 4]);
        const uint32_t sho = byte_of<i % 4>(LSo[i / 4]);
        const uint32_t o = byte_of<i % 4>(OFF[i / 4]);
        uint32_t wl;
        if constexpr (BT::GATHER32) {
            const uint32_t x = (uint32_t)(wb >> shw);
            const uint32_t g = ((x & (uint32_t)BT::COL0) * (uint32_t)BT::MAGIC) >> BT::GSHIFT;
            wl = horiz ? x : g;
        } else {
            const board_t x = wb >> shw;
            const uint32_t g = (uint32_t)(((x & (board_t)BT::COL0) * (board_t)BT::MAGIC) >> BT::GSHIFT);
            wl = horiz ? (uint32_t)x : g;
        }
        const uint32_t ol = (uint32_t)(occ >> sho);
        const uint32_t above = 0xFFFFFFFEu << o;           // offsets > o
        const uint32_t blk = (wl & above) | (1u << S);     // walls above o, edge sentinel at S
        const uint32_t run = (blk - 1u) & ~blk;            // offsets below the nearest wall
        const uint32_t empty = run & above & ~ol;
        Q[i / 4] += (uint32_t)__popc(empty) << (8 * (i % 4));
    });
#pragma unroll
    for (int w = 0; w < PR; ++w) {
        uint32_t x = flip ? KFLIP - Q[w] : Q[w];
        q[w] = horiz ? x : swap_nibbles(x);
    }
    // unused bytes of the last word were zero and come back zero:
    // KFLIP-(KFLIP-0)=0 and swap_nibbles(0)=0; no count was added to them.
}

// occupancy bitboard (bit r*S+c) of packed positions
template <int S, int T>
__device__ __forceinline__ typename BoardTraits<S>::board_t occupancy(const uint32_t (&q)[(T + 3) / 4]) {
    using board_t = typename BoardTraits<S>::board_t;
    board_t occ = 0;
    uint32_t P[(T + 3) / 4];
#pragma unroll
    for (int w = 0; w < (T + 3) / 4; ++w) P[w] = ((q[w] >> 4) & 0x0F0F0F0Fu) * (uint32_t)S + (q[w] & 0x0F0F0F0Fu);
    static_for<0, T>([&](auto I) {
        constexpr int i = decltype(I)::value;
        occ |= (board_t)1 << byte_of<i % 4>(P[i / 4]);
    });
    return occ;
}

// flag bits (mirrored in include/tiler_slider.h)
constexpr uint32_t F_DONE = 1, F_WON = 2, F_INVALID = 4, F_TIMEOUT = 8, F_STALE = 16;

}  // namespace ts
