// K1 - candidate scan: find every offset i with (file[i], file[i+1]) one of the 24 zlib headers AntiZ accepts
// (ZBuffSearcher::parseOffsetType main.cpp:168-203: CM=8, CINFO 2..7, FDICT=0, (CMF*256+FLG) % 31 == 0).
// HBM-bound: 16-byte vector loads, a 4-bytes-at-a-time prefilter on the first header byte, two passes over 64 KiB tiles
// (count, exclusive scan of the tile counts, ordered write) so the offsets come out sorted without a sort.
// A single-pass variant (tile tickets + decoupled look-back for the output position) was built and measured in round 2: the same
// 1.5 ms per GB inside the running program (the phase is a kernel of about a millisecond plus a host round trip for the count), so
// the two passes were kept: no spin-waits, exact buffer sizes.
#include "common.cuh"

namespace atz {

#define SCAN_THREADS 256
#define SCAN_TILE (SCAN_THREADS * 16 * 16) /* 64 KiB per CTA */

__device__ __forceinline__ bool is_magic(uint32_t b0, uint32_t b1) {
    // closed form of the 24-way switch (checked exhaustively against it in tests/test_host_logic.py)
    return (b0 & 0x8fu) == 0x08u && b0 >= 0x28u && (b1 & 0x20u) == 0 && ((b0 << 8) | b1) % 31u == 0;
}
// 16 consecutive positions starting at byte `pos` (a multiple of 16); bit k set <=> (pos+k, pos+k+1) is a header, pos+k is in
// [lo, hi) (the part of the file this launch scans: a shard's chunk range, api.cu atz_scan_shard) and pos+k+1 < n
__device__ __forceinline__ uint32_t magic_mask16(const uint8_t *file, uint64_t pos, uint64_t lo, uint64_t hi, uint64_t n) {
    if (pos >= hi || pos + 16 <= lo) return 0;
    const uint4 v = __ldg((const uint4 *)(file + pos));          // buffer is padded: always in bounds
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint64_t A = ((uint64_t)v.y << 32) | v.x, B = ((uint64_t)v.w << 32) | v.z;      // (dynamic byte picks without indexing registers)
    // four bytes at a time: a header's first byte is 0x28, 0x38, ... 0x78 (low nibble 8, top bit clear, >= 0x28): 6 values of 256, so
    // most 16-byte groups have none and the per-position test (FDICT, FCHECK, bounds) runs for the few bytes that qualify
    uint32_t q = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t hit = __vcmpeq4(w[j] & 0x8f8f8f8fu, 0x08080808u) & __vcmpgeu4(w[j], 0x28282828u) & 0x01010101u;
        q |= ((hit & 1u) | ((hit >> 7) & 2u) | ((hit >> 14) & 4u) | ((hit >> 21) & 8u)) << (4 * j);
    }
    if (!q) return 0;
    const uint32_t nxt = __ldg(file + pos + 16);
    uint32_t m = 0;
    while (q) {
        const uint32_t k = (uint32_t)__ffs((int)q) - 1; q &= q - 1;
        const uint32_t b0 = (uint32_t)((k < 8 ? A >> (8 * k) : B >> (8 * (k - 8))) & 0xff);
        const uint32_t b1 = k == 15 ? nxt : (uint32_t)((k < 7 ? A >> (8 * (k + 1)) : B >> (8 * (k - 7))) & 0xff);
        if (is_magic(b0, b1) && pos + k + 1 < n && pos + k >= lo && pos + k < hi) m |= 1u << k;
    }
    return m;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_count_kernel(const uint8_t *file, uint64_t lo, uint64_t hi, uint64_t n, uint32_t *tile_counts) {
    const uint64_t tile0 = (lo & ~(uint64_t)15) + (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t c = 0;
#pragma unroll 4
    for (int it = 0; it < 16; it++) c += __popc(magic_mask16(file, tile0 + ((uint64_t)it * SCAN_THREADS + threadIdx.x) * 16, lo, hi, n));
    c = __reduce_add_sync(FULL, c);
    __shared__ uint32_t ws[SCAN_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < SCAN_THREADS / 32; i++) t += ws[i]; tile_counts[blockIdx.x] = t; }
}

// exclusive scan of the tile counts in place; total -> *total (single CTA; tiles <= 65536 for a 4 GiB file)
__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint32_t *tile_counts, uint32_t ntiles, uint32_t *total) {
    __shared__ uint32_t wsum[32]; __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t b = 0; b < ntiles; b += 1024) {
        uint32_t i = b + threadIdx.x, v = i < ntiles ? tile_counts[i] : 0, tot, ex = warp_excl_scan(v, tot);
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = tot;
        __syncthreads();
        uint32_t woff = 0; for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) woff += wsum[w];
        uint32_t carry = carry_s;
        if (i < ntiles) tile_counts[i] = carry + woff + ex;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_write_kernel(const uint8_t *file, uint64_t lo, uint64_t hi, uint64_t n, const uint32_t *tile_base, uint32_t *cand, uint8_t *ctype, uint32_t cap) {
    const uint64_t tile0 = (lo & ~(uint64_t)15) + (uint64_t)blockIdx.x * SCAN_TILE;
    __shared__ uint32_t wsum[SCAN_THREADS / 32]; __shared__ uint32_t run_s;
    if (threadIdx.x == 0) run_s = tile_base[blockIdx.x];
    __syncthreads();
    for (int it = 0; it < 16; it++) {
        uint64_t pos = tile0 + ((uint64_t)it * SCAN_THREADS + threadIdx.x) * 16;
        uint32_t m = magic_mask16(file, pos, lo, hi, n), c = __popc(m), tot, ex = warp_excl_scan(c, tot);
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = tot;
        __syncthreads();
        uint32_t woff = 0, all = 0;
        for (uint32_t w = 0; w < SCAN_THREADS / 32; w++) { if (w < (threadIdx.x >> 5)) woff += wsum[w]; all += wsum[w]; }
        uint32_t o = run_s + woff + ex;
        while (m) { uint32_t k = __ffs((int)m) - 1; m &= m - 1; if (o < cap) { cand[o] = (uint32_t)(pos + k); uint32_t b0 = __ldg(file + pos + k), b1 = __ldg(file + pos + k + 1); ctype[o] = (uint8_t)(4 * ((b0 >> 4) - 2) + (b1 >> 6)); } o++; }
        __syncthreads();
        if (threadIdx.x == 0) run_s += all;
        __syncthreads();
    }
}

// `file` is the address of file offset 0 (16 B aligned; only [lo & ~15, hi + 16) has to be mapped), positions lo <= i < hi are scanned
uint32_t scan_tiles_for(uint64_t lo, uint64_t hi) { return hi > lo ? (uint32_t)((hi - (lo & ~(uint64_t)15) + SCAN_TILE - 1) / SCAN_TILE) : 0u; }
cudaError_t launch_scan_count(const uint8_t *file, uint64_t lo, uint64_t hi, uint64_t n, uint32_t *tile_counts, uint32_t *total, cudaStream_t s) {
    uint32_t nt = scan_tiles_for(lo, hi);
    if (nt) scan_count_kernel<<<nt, SCAN_THREADS, 0, s>>>(file, lo, hi, n, tile_counts);
    scan_tiles_kernel<<<1, 1024, 0, s>>>(tile_counts, nt, total);
    return cudaGetLastError();
}
cudaError_t launch_scan_write(const uint8_t *file, uint64_t lo, uint64_t hi, uint64_t n, const uint32_t *tile_base, uint32_t *cand, uint8_t *ctype, uint32_t cap, cudaStream_t s) {
    if (scan_tiles_for(lo, hi)) scan_write_kernel<<<scan_tiles_for(lo, hi), SCAN_THREADS, 0, s>>>(file, lo, hi, n, tile_base, cand, ctype, cap);
    return cudaGetLastError();
}

// K4 - diff compaction for one winner: positions i < min(C', C) with out[i] != orig[i], then i in [C', C)
// (main.cpp:699-712).  One warp per job, ordered append by ballot.
struct DiffJob { const uint8_t *out; const uint8_t *orig; uint32_t cprime, c; uint32_t *pos; uint8_t *val; uint32_t cap; uint32_t *count; };
__global__ void __launch_bounds__(128) diff_kernel(const DiffJob *jobs, uint32_t njobs) {
    uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (wid >= njobs) return;
    const DiffJob j = jobs[wid];
    uint32_t smaller = j.cprime < j.c ? j.cprime : j.c, nd = 0;
    for (uint32_t i0 = 0; i0 < j.c; i0 += 32) {
        uint32_t i = i0 + lane; bool d = false; uint8_t ov = 0;
        if (i < j.c) { ov = j.orig[i]; d = i >= smaller || j.out[i] != ov; }
        uint32_t bm = __ballot_sync(FULL, d);
        if (d) { uint32_t o = nd + __popc(bm & ((1u << lane) - 1)); if (o < j.cap) { j.pos[o] = i; j.val[o] = ov; } }
        nd += __popc(bm);
    }
    if (lane == 0) *j.count = nd;
}
cudaError_t launch_diff(const DiffJob *jobs, uint32_t njobs, cudaStream_t s) {
    diff_kernel<<<(njobs + 3) / 4, 128, 0, s>>>(jobs, njobs);
    return cudaGetLastError();
}

// Gather: copy n bytes src -> dst for many (src, dst, n) jobs in one launch (the recompressed streams' plaintext, concatenated
// for a single D2H copy; what writeStreamdesc's per-stream re-inflate produced, main.cpp:824-828).  Sources are 16 B aligned
// with ATZ_PAD slack; destinations have any alignment: head/tail bytes singly, the middle as dst-aligned 32-bit words.
struct CopyJob { const uint8_t *src; uint8_t *dst; uint64_t n; };
__global__ void __launch_bounds__(256) gather_kernel(const CopyJob *jobs, uint32_t njobs) {
    for (uint32_t ji = blockIdx.x; ji < njobs; ji += gridDim.x) {
        const CopyJob j = jobs[ji];
        uint64_t head = (4 - ((uintptr_t)j.dst & 3)) & 3; if (head > j.n) head = j.n;
        const uint64_t words = (j.n - head) >> 2, tail0 = head + 4 * words;
        if (threadIdx.x < head) j.dst[threadIdx.x] = j.src[threadIdx.x];
        uint32_t *dw = (uint32_t *)(j.dst + head);
        for (uint64_t w = threadIdx.x; w < words; w += blockDim.x) dw[w] = ldu32(j.src + head + 4 * w);
        if (tail0 + threadIdx.x < j.n) j.dst[tail0 + threadIdx.x] = j.src[tail0 + threadIdx.x];
    }
}
cudaError_t launch_gather(const CopyJob *jobs, uint32_t njobs, cudaStream_t s) {
    if (!njobs) return cudaSuccess;
    uint32_t ctas = njobs < 148u * 8u ? njobs : 148u * 8u;
    gather_kernel<<<ctas, 256, 0, s>>>(jobs, njobs);
    return cudaGetLastError();
}

} // namespace atz
