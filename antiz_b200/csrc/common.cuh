// antiz_b200 - shared device helpers and host<->device record layouts (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FULL 0xffffffffu
#define ATZ_PAD 512u /* every device buffer that kernels read with unaligned word loads has >= this much zeroed slack */

namespace atz {

// ---------------------------------------------------------------------------------------------
// Records shared by host and kernels
// ---------------------------------------------------------------------------------------------
struct ChainRef {            // bucket lists of one (plaintext, hash_bits): see chains.cu
    const uint32_t *list;    // positions sorted by (hash, position)
    const uint32_t *idx;     // idx[p]  = slot of p in list
    const uint16_t *lsth;    // lsth[s] = hash of list[s]: the chain of p runs down from slot idx[p]-1 while this stays equal
    // optional row table (deflate.cu, build_rows_kernel): 32 bytes per position p < rlen
    const uint4 *rec; uint32_t rlen; uint32_t rbudget;   // rbudget: number of chain candidates the table has looked at
};
#define REC_VALID 0x08000000u
// token map of the ORIGINAL stream (one byte per plaintext position, written by the inflate kernel in produce mode):
// 0 = nothing known; 1 = a literal starts here; 3..250 = a match of that length starts here; TM_LONG = a match of
// 251..258 bytes starts here; TM_INNER + c = inside a match whose length is <= 4 (c = 3), 5 (c = 2), 6 (c = 1), longer (c = 0):
// deflate_fast at level L inserts the inner positions of a match iff its length <= max_insert_length = 3 + L (Z/deflate.c:1680).
#define RES_ABSENT 0x1ffu   /* resolved-table length field: no row here, walk the chain */
#define TM_LONG 251u
#define TM_INNER 252u

struct TrialDesc {
    const uint8_t *in;       // plaintext (any alignment; readable slack of ATZ_PAD bytes behind it and 3 before it)
    const uint8_t *orig;     // original compressed stream to compare with (any alignment) or nullptr
    uint8_t *out;            // store mode: output buffer (4 B aligned) or nullptr
    const uint8_t *tmap;     // token map of the original stream (8 B aligned) or nullptr: the hypothesis of deflate_fast rows, the speculation points of the burst parse
    const uint2 *res;        // levels 4-9: resolved table of this (level, window) for positions < ch.rlen (deflate.cu resolve_rows_kernel) or nullptr
    ChainRef ch;             // unused for level 0
    uint32_t n;              // plaintext length U
    uint32_t c;              // original stream length C
    uint32_t out_cap;        // store mode capacity in bytes (multiple of 4)
    uint32_t adler;          // adler32(plaintext)
    uint8_t level, wbits, memlevel, store;
    uint16_t phase1;         // 1 = stop with TR_PASSED once the --shortcut-len prefix has been compared and accepted
    uint16_t strategy;       // zlib's strategy: 0 Z_DEFAULT_STRATEGY (all the reference ever uses), 1 Z_FILTERED, 2 Z_HUFFMAN_ONLY, 3 Z_RLE, 4 Z_FIXED
};

struct TrialOpts {
    uint32_t shortcut;       // --shortcut-len S; the shortcut applies iff C > S (main.cpp:632)
    uint32_t bail_below;     // bail iff ident over the first min(S, C') bytes < this (main.cpp:649; 0xffffffff = always)
    uint32_t sizediff;       // --sizediff-tresh (main.cpp:671)
    uint32_t cut_mismatch;   // early cut when mismatches exceed this (0xffffffff = never; DESIGN.md "early cut")
    uint32_t compare;        // 1 = search trial (compare with orig), 0 = plain deflate
    uint32_t burst;          // 1 = burst parse where the original's token map allows it (0: test hook ATZ_BURST=0)
};

enum { TR_COMPARED = 0, TR_BAILED = 1, TR_SIZE = 2, TR_CUT = 3, TR_OVERFLOW = 4, TR_PASSED = 5 };
struct TrialResult {
    int32_t status;
    uint32_t in_consumed;    // plaintext bytes parsed when the trial stopped (algorithmic-bytes accounting)
    uint32_t out_len;        // C' (or bytes produced so far if stopped early)
    uint32_t ident;          // equal bytes over min(C', C)
    uint32_t kcycles, kcycles_flush;   // SM kilocycles spent in the trial / in its block flushes (profiling aid)
};

struct InflateJob {          // one candidate stream
    uint64_t off;            // file offset of the zlib header
    uint64_t avail;          // bytes available up to the end of the candidate's chunk
    uint64_t vtotal;         // virtual input length including the following chunks (== avail: no continuation)
    uint64_t out_off;        // offset of the output region in the arena
    uint64_t out_cap;        // size of the output region
    uint64_t tmap_off;       // offset of this stream's token map in the arena, or ~0 for none
    // The reference's chunk reader keeps the wrong byte for the overlap (searchInfile, main.cpp:411-414: `LastByte = rBuffer[f.gcount() - 1]`
    // is the last byte of chunk 0 but the SECOND TO LAST of every later chunk), so the first byte of chunk k >= 2 is file[start_k - 1]:
    //   bit 0      : this candidate starts at such a chunk start - its first input byte is file[off - 1], the rest file[off + 1 ...]
    //   bits 8..15 : how many of the chunk boundaries a continuation crosses still repeat the true last byte (1 for a candidate of chunk 0)
    uint64_t flags;
};
#define INFJ_FIRST_FROM_PREV 1ull
enum { INF_END = 0, INF_NEED_INPUT = 1, INF_DATA_ERROR = 2, INF_NEED_DICT = 3, INF_OUT_FULL = 4 };
struct InflateResult {
    int32_t status; uint32_t adler;
    uint64_t total_in, total_out, in_at_outcap;
};

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Unaligned 32-bit read through the read-only path. Touches up to 3 bytes before and 7 after p: callers
// guarantee the slack (ATZ_PAD, 16 B aligned arenas).
__device__ __forceinline__ uint32_t ldu32(const uint8_t *p) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    uint32_t lo = __ldg(w), hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, (uint32_t)(a & 3) * 8u);
}
// Same for memory written earlier by this kernel (no .nc path).
__device__ __forceinline__ uint32_t ldu32_rw(const uint8_t *p) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    uint32_t lo = w[0], hi = w[1];
    return __funnelshift_r(lo, hi, (uint32_t)(a & 3) * 8u);
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t &total) {
    uint32_t lane = lane_id(), x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(FULL, x, d); if (lane >= (uint32_t)d) x += y; }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}

// zlib's 3-byte rolling hash after three updates (Z/deflate.c:167, hash_shift = (hash_bits+2)/3 Z/deflate.c:291)
__device__ __forceinline__ uint32_t hash3(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t shift, uint32_t mask) {
    return ((((b0 << shift) ^ b1) << shift) ^ b2) & mask;
}

} // namespace atz
