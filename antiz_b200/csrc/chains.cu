// Bucket lists ("hash chains as arrays") + adler32.
//
// For one plaintext and one hash_bits (= memLevel + 7, Z/deflate.c:288) zlib's head[]/prev[] tables define, for
// every position p, the chain of earlier positions with the same 3-byte hash, most recent first
// (INSERT_STRING Z/deflate.c:186-189).  For levels 4-9 every position is inserted, so that chain is a pure
// function of the data: here it is materialised once as
//     list[]  all positions p (p + 2 < n) sorted by (hash(p), p)
//     lsth[s] hash of list[s] (16 bit)
//     idx[p]  slot of p in list[]
// and shared by all level x window trials of the stream.  The chain of p is list[idx[p]-1], list[idx[p]-2], ... for as
// long as lsth[] stays equal, which K3 reads 32 entries at a time.
// Built for ALL (plaintext, hash_bits) tasks of a wave together by a multi-block LSD radix sort (1 or 2 stable 8-bit
// passes): per pass a count kernel (digit histogram of every 2048-entry chunk), a scan kernel (one CTA per task) and a
// scatter kernel (stable ranks inside a chunk from __match_any_sync and per-warp digit counts).  Chunks are handed out in
// task order, so at any moment the GPU works on a few tasks and their scattered 4-byte writes meet in L2.
#include "common.cuh"

namespace atz {

struct ChainTask {
    const uint8_t *in; uint32_t n; uint32_t hbits;
    uint32_t *list; uint32_t *idx; uint16_t *lsth;
    uint32_t *tmp; uint16_t *tmph;      // pass-1 output (positions by low digit, with their hashes); unused when hbits <= 8
    uint32_t chunk0, nchunks;
};

#define CH_THREADS 256
#define CH_WARPS (CH_THREADS / 32)
#define CH_TILES 8
#define CH_CHUNK (CH_THREADS * CH_TILES)

__device__ __forceinline__ uint32_t hash_at(const uint8_t *in, uint32_t p, uint32_t shift, uint32_t mask) {
    const uint32_t w = ldu32(in + p);
    return hash3(w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, shift, mask);
}
__device__ __forceinline__ uint32_t task_of_chunk(const ChainTask *tasks, uint32_t ntasks, uint32_t ch) {
    uint32_t lo = 0, hi = ntasks - 1;
    while (lo < hi) { uint32_t mid = (lo + hi + 1) >> 1; if (tasks[mid].chunk0 <= ch) lo = mid; else hi = mid - 1; }
    return lo;
}

// digit histogram of every chunk.  HI = second pass (high byte of the hash, entries in pass-1 order)
template <bool HI>
__global__ void __launch_bounds__(CH_THREADS) chain_count_kernel(const ChainTask *tasks, uint32_t ntasks, uint32_t nchunks, uint32_t *hist) {
    __shared__ uint32_t h256[256];
    for (uint32_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const ChainTask t = tasks[task_of_chunk(tasks, ntasks, ch)];
        h256[threadIdx.x] = 0;
        __syncthreads();
        if (!(HI && t.hbits <= 8)) {
            const uint32_t np = t.n >= 3 ? t.n - 2 : 0, mask = (1u << t.hbits) - 1, shift = (t.hbits + 2) / 3;
            const uint32_t s0 = (ch - t.chunk0) * CH_CHUNK;
#pragma unroll
            for (int k = 0; k < CH_TILES; k++) {
                const uint32_t s = s0 + k * CH_THREADS + threadIdx.x;
                if (s < np) { const uint32_t h = HI ? (uint32_t)t.tmph[s] : hash_at(t.in, s, shift, mask); atomicAdd(&h256[HI ? h >> 8 : h & 255u], 1u); }
            }
        }
        __syncthreads();
        hist[(size_t)ch * 256 + threadIdx.x] = h256[threadIdx.x];
        __syncthreads();
    }
}

// per task: hist[chunk][d] -> number of entries with digit d in earlier chunks; dbase[task][d] -> entries with a smaller digit
__global__ void __launch_bounds__(256) chain_scan_kernel(const ChainTask *tasks, uint32_t *hist, uint32_t *dbase) {
    __shared__ uint32_t wtot[8];
    const ChainTask t = tasks[blockIdx.x];
    const uint32_t d = threadIdx.x;
    uint32_t run = 0;
    uint32_t *h = hist + (size_t)t.chunk0 * 256 + d;
#pragma unroll 4
    for (uint32_t c = 0; c < t.nchunks; c++) { const uint32_t v = h[(size_t)c * 256]; h[(size_t)c * 256] = run; run += v; }
    uint32_t tot, ex = warp_excl_scan(run, tot);
    if ((d & 31) == 31) wtot[d >> 5] = tot;
    __syncthreads();
    uint32_t off = 0; for (uint32_t w = 0; w < (d >> 5); w++) off += wtot[w];
    dbase[(size_t)blockIdx.x * 256 + d] = off + ex;
}

struct ScatterSmem { uint32_t base[256]; uint16_t wcnt[CH_WARPS][256]; };

// stable scatter of every chunk by the digit.  FINAL semantics (list/lsth/idx) for the only pass of tasks with hbits <= 8 and
// for the second pass of the others.  Every warp owns 256 consecutive entries of the chunk (8 tiles of 32): it ranks them among
// its own entries with __match_any_sync and a running per-digit count in shared memory, the block turns the per-warp counts into
// offsets (one thread per digit), and every entry then knows its destination: two block barriers per 2048 entries.
template <bool HI>
__global__ void __launch_bounds__(CH_THREADS) chain_scatter_kernel(const ChainTask *tasks, uint32_t ntasks, uint32_t nchunks, const uint32_t *hist, const uint32_t *dbase) {
    __shared__ ScatterSmem sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const uint32_t ti = task_of_chunk(tasks, ntasks, ch);
        const ChainTask t = tasks[ti];
        if (HI && t.hbits <= 8) continue;
        const bool final = HI || t.hbits <= 8;
        const uint32_t np = t.n >= 3 ? t.n - 2 : 0, mask = (1u << t.hbits) - 1, shift = (t.hbits + 2) / 3;
        const uint32_t c0 = (ch - t.chunk0) * CH_CHUNK + warp * (32 * CH_TILES);
        __syncthreads();     // the previous chunk's offsets are no longer read
        sm.base[tid] = hist[(size_t)ch * 256 + tid] + dbase[(size_t)ti * 256 + tid];
        { uint32_t *z = (uint32_t *)sm.wcnt[warp]; for (uint32_t q = lane; q < 128; q += 32) z[q] = 0; }
        __syncwarp();
        uint32_t pp[CH_TILES], hh[CH_TILES], pos[CH_TILES];
#pragma unroll
        for (int k = 0; k < CH_TILES; k++) {
            const uint32_t s = c0 + 32 * k + lane; const bool ok = s < np;
            const uint32_t p = ok ? (HI ? t.tmp[s] : s) : 0;
            const uint32_t h = ok ? (HI ? (uint32_t)t.tmph[s] : hash_at(t.in, p, shift, mask)) : 0;
            const uint32_t d = ok ? (HI ? h >> 8 : h & 255u) : 0xffffffffu;
            const uint32_t peers = __match_any_sync(FULL, d), rank = __popc(peers & ((1u << lane) - 1)), leader = (uint32_t)__ffs((int)peers) - 1;
            uint32_t before = 0;
            if (ok && rank == 0) { before = sm.wcnt[warp][d]; sm.wcnt[warp][d] = (uint16_t)(before + __popc(peers)); }
            before = __shfl_sync(FULL, before, leader);
            __syncwarp();
            pp[k] = p; hh[k] = ok ? h : 0xffffffffu; pos[k] = before + rank;
        }
        __syncthreads();
        {   // digit tid: per-warp counts -> exclusive offsets over the warps, on top of the digit's base
            uint32_t run = sm.base[tid];
#pragma unroll
            for (uint32_t w = 0; w < CH_WARPS; w++) { const uint32_t c = sm.wcnt[w][tid]; sm.wcnt[w][tid] = (uint16_t)(run - sm.base[tid]); run += c; }
        }
        __syncthreads();
        uint32_t *dst = final ? t.list : t.tmp; uint16_t *dsth = final ? t.lsth : t.tmph;
#pragma unroll
        for (int k = 0; k < CH_TILES; k++) {
            if (hh[k] == 0xffffffffu) continue;
            const uint32_t h = hh[k], d = HI ? h >> 8 : h & 255u;
            const uint32_t dest = sm.base[d] + sm.wcnt[warp][d] + pos[k];
            dst[dest] = pp[k]; dsth[dest] = (uint16_t)h; if (final) t.idx[pp[k]] = dest;
        }
    }
}

cudaError_t launch_build_chains(const ChainTask *tasks, uint32_t ntasks, uint32_t nchunks, uint32_t *hist, uint32_t *dbase, bool any_two_pass, cudaStream_t s) {
    const uint32_t grid = nchunks < 148u * 8u ? nchunks : 148u * 8u;
    chain_count_kernel<false><<<grid, CH_THREADS, 0, s>>>(tasks, ntasks, nchunks, hist);
    chain_scan_kernel<<<ntasks, 256, 0, s>>>(tasks, hist, dbase);
    chain_scatter_kernel<false><<<grid, CH_THREADS, 0, s>>>(tasks, ntasks, nchunks, hist, dbase);
    if (any_two_pass) {
        chain_count_kernel<true><<<grid, CH_THREADS, 0, s>>>(tasks, ntasks, nchunks, hist);
        chain_scan_kernel<<<ntasks, 256, 0, s>>>(tasks, hist, dbase);
        chain_scatter_kernel<true><<<grid, CH_THREADS, 0, s>>>(tasks, ntasks, nchunks, hist, dbase);
    }
    return cudaGetLastError();
}
uint32_t chain_chunk_size() { return CH_CHUNK; }

// adler32 (Z/adler32.c:65-133) of n bytes, one CTA per job: per-thread partial (a, b) over a contiguous slice, then a
// weighted tree combine (adler32_combine's identity: b_total = b1 + b2 + len2 * (a1 - 1)).
struct AdlerJob { const uint8_t *in; uint32_t n; uint32_t *out; };
__global__ void __launch_bounds__(256) adler_kernel(const AdlerJob *jobs) {
    const AdlerJob j = jobs[blockIdx.x];
    __shared__ uint32_t sa[256], sb[256], sl[256];
    const uint32_t T = blockDim.x, t = threadIdx.x;
    uint32_t per = (j.n + T - 1) / T; if (per == 0) per = 1;
    uint64_t beg = (uint64_t)t * per, end = beg + per; if (beg > j.n) beg = j.n; if (end > j.n) end = j.n;
    uint32_t a = 0, b = 0, len = (uint32_t)(end - beg);   // partial sums relative to a = 0 (no +1 yet)
    for (uint64_t i = beg; i < end;) {
        uint32_t k = (uint32_t)(end - i); if (k > 3800) k = 3800;   // keep b below 2^32: 3800*3801/2*255 + ... safe
        for (uint32_t e = 0; e < k; e++) { a += __ldg(j.in + i + e); b += a; }
        a %= 65521u; b %= 65521u; i += k;
    }
    sa[t] = a; sb[t] = b; sl[t] = len;
    __syncthreads();
    if (t == 0) {   // serial combine of 256 partials (cheap)
        uint64_t A = 1, B = 0;
        for (uint32_t k = 0; k < T; k++) {
            uint64_t l2 = sl[k] % 65521u;
            B = (B + sb[k] + l2 * A) % 65521u;   // every byte of slice k sees the running a of all earlier slices
            A = (A + sa[k]) % 65521u;
        }
        *j.out = (uint32_t)((B << 16) | A);
    }
}
cudaError_t launch_adler(const AdlerJob *jobs, uint32_t njobs, cudaStream_t s) {
    adler_kernel<<<njobs, 256, 0, s>>>(jobs);
    return cudaGetLastError();
}

} // namespace atz
