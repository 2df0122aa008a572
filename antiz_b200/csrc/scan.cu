// K1 - candidate scan: find every offset i with (file[i], file[i+1]) one of the 24 zlib headers AntiZ accepts
// (ZBuffSearcher::parseOffsetType main.cpp:168-203: CM=8, CINFO 2..7, FDICT=0, (CMF*256+FLG) % 31 == 0).
// HBM-bound: 16-byte vector loads, one byte of look-ahead per thread, ONE pass over 64 KiB tiles (decoupled look-back
// for the output position of a tile's hits), so the offsets come out sorted without a sort.
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
    uint4 v = __ldg((const uint4 *)(file + pos));          // buffer is padded: always in bounds
    uint32_t nxt = __ldg(file + pos + 16);
    uint32_t w[5] = {v.x, v.y, v.z, v.w, nxt};
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        uint32_t b0 = (w[k >> 2] >> (8 * (k & 3))) & 0xff;
        uint32_t b1 = (w[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xff;
        if (is_magic(b0, b1) && pos + k + 1 < n && pos + k >= lo && pos + k < hi) m |= 1u << k;
    }
    return m;
}

// Single pass (the file is read once): a CTA takes the next tile (ticket), finds the headers of its 64 KiB (16 x 16 positions per
// thread, the 16-bit masks stay in registers), publishes its count, gets the number of hits in all tiles before it by looking back
// at the tiles in flight (decoupled look-back: a tile publishes its own count at once and its inclusive prefix as soon as it knows
// it), and writes its hits in order.  Offsets come out sorted without a sort and without a second read.
//   state[t] = status << 32 | value: status 0 = not yet, 1 = value is tile t's own count, 2 = value is the inclusive prefix up to t
//   ctl[0] = tile ticket, ctl[1] = total number of hits (written by the last tile)
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const uint8_t *file, uint64_t lo, uint64_t hi, uint64_t n, uint32_t ntiles,
                                                            unsigned long long *state, uint32_t *ctl, uint32_t *cand, uint8_t *ctype, uint32_t cap) {
    __shared__ uint32_t tile_s, base_s, wsum[SCAN_THREADS / 32];
    if (threadIdx.x == 0) tile_s = atomicAdd(&ctl[0], 1u);
    __syncthreads();
    const uint32_t tile = tile_s;
    if (tile >= ntiles) return;
    const uint64_t tile0 = (lo & ~(uint64_t)15) + (uint64_t)tile * SCAN_TILE;
    uint32_t masks[8], c = 0;
#pragma unroll
    for (int it = 0; it < 16; it += 2) {
        const uint32_t m0 = magic_mask16(file, tile0 + ((uint64_t)it * SCAN_THREADS + threadIdx.x) * 16, lo, hi, n);
        const uint32_t m1 = magic_mask16(file, tile0 + ((uint64_t)(it + 1) * SCAN_THREADS + threadIdx.x) * 16, lo, hi, n);
        masks[it >> 1] = m0 | (m1 << 16); c += __popc(m0) + __popc(m1);
    }
    const uint32_t wtot = __reduce_add_sync(FULL, c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = wtot;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int i = 0; i < SCAN_THREADS / 32; i++) total += wsum[i];
        uint32_t excl = 0;
        if (tile == 0) atomicExch(&state[0], (2ull << 32) | total);
        else {
            atomicExch(&state[tile], (1ull << 32) | total);
            for (int64_t t = (int64_t)tile - 1; t >= 0; t--) {
                unsigned long long v;
                while (((v = atomicAdd(&state[t], 0ull)) >> 32) == 0) __nanosleep(20);
                excl += (uint32_t)v;
                if ((v >> 32) == 2) break;
            }
            atomicExch(&state[tile], (2ull << 32) | (excl + total));
        }
        if (tile == ntiles - 1) ctl[1] = excl + total;
        base_s = excl;
    }
    __syncthreads();
    // ordered write: iteration-major, thread-minor (file order), as the positions were assigned
    uint32_t run = base_s;
#pragma unroll 1
    for (int it = 0; it < 16; it++) {
        uint32_t m = (masks[it >> 1] >> ((it & 1) * 16)) & 0xffffu;
        const uint64_t pos = tile0 + ((uint64_t)it * SCAN_THREADS + threadIdx.x) * 16;
        uint32_t tot, ex = warp_excl_scan(__popc(m), tot);
        __syncthreads();
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = tot;
        __syncthreads();
        uint32_t woff = 0, all = 0;
        for (uint32_t w = 0; w < SCAN_THREADS / 32; w++) { if (w < (threadIdx.x >> 5)) woff += wsum[w]; all += wsum[w]; }
        uint32_t o = run + woff + ex;
        while (m) {
            const uint32_t k = __ffs((int)m) - 1; m &= m - 1;
            if (o < cap) { cand[o] = (uint32_t)(pos + k); const uint32_t b0 = __ldg(file + pos + k), b1 = __ldg(file + pos + k + 1); ctype[o] = (uint8_t)(4 * ((b0 >> 4) - 2) + (b1 >> 6)); }
            o++;
        }
        run += all;
    }
}

// `file` is the address of file offset 0 (16 B aligned; only [lo & ~15, hi + 16) has to be mapped), positions lo <= i < hi are scanned
uint32_t scan_tiles_for(uint64_t lo, uint64_t hi) { return hi > lo ? (uint32_t)((hi - (lo & ~(uint64_t)15) + SCAN_TILE - 1) / SCAN_TILE) : 0u; }
// state: ntiles x 8 bytes, ctl: 2 x 4 bytes; both zeroed here.  *ctl[1] receives the number of hits (may exceed cap: then only the
// first cap were stored and the caller repeats the scan with a larger buffer)
cudaError_t launch_scan(const uint8_t *file, uint64_t lo, uint64_t hi, uint64_t n, unsigned long long *state, uint32_t *ctl, uint32_t *cand, uint8_t *ctype, uint32_t cap, cudaStream_t s) {
    const uint32_t nt = scan_tiles_for(lo, hi);
    cudaMemsetAsync(ctl, 0, 8, s);
    if (!nt) return cudaGetLastError();
    cudaMemsetAsync(state, 0, (size_t)nt * 8, s);
    scan_kernel<<<nt, SCAN_THREADS, 0, s>>>(file, lo, hi, n, nt, state, ctl, cand, ctype, cap);
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
