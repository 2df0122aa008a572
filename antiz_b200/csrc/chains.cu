// Bucket lists ("hash chains as arrays") + adler32.
//
// For one plaintext and one hash_bits (= memLevel + 7, Z/deflate.c:288) zlib's head[]/prev[] tables define, for
// every position p, the chain of earlier positions with the same 3-byte hash, most recent first
// (INSERT_STRING Z/deflate.c:186-189).  For levels 4-9 every position is inserted, so that chain is a pure
// function of the data: here it is materialised once as
//     list[]  all positions p (p + 2 < n) sorted by (hash(p), p)
//     idx[p]  slot of p in list[]
//     cnt[p]  how many earlier positions share p's bucket (saturating u16)
// and shared by all level x window trials of the stream.  The chain of p is list[idx[p]-1], list[idx[p]-2], ...
// which K3 reads 32 entries at a time.  One warp builds one (plaintext, hash_bits) task: histogram by atomics,
// warp scan, then a stable fill that ranks equal hashes inside each 32-position group with __match_any_sync.
#include "common.cuh"

namespace atz {

struct ChainTask {
    const uint8_t *in; uint32_t n; uint32_t hbits;
    uint32_t *list; uint32_t *idx; uint16_t *cnt;
};

__global__ void __launch_bounds__(128) build_chains_kernel(const ChainTask *tasks, uint32_t ntasks, uint32_t *queue, uint32_t *tab_all) {
    const uint32_t lane = lane_id(), wpc = blockDim.x >> 5, slot = blockIdx.x * wpc + (threadIdx.x >> 5);
    uint32_t *tab = tab_all + (size_t)slot * 65536u;   // bucket cursor table of this warp
    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(queue, 1u);
        ti = __shfl_sync(FULL, ti, 0);
        if (ti >= ntasks) break;
        const ChainTask t = tasks[ti];
        const uint32_t np = t.n >= 3 ? t.n - 2 : 0, hsize = 1u << t.hbits, mask = hsize - 1, shift = (t.hbits + 2) / 3;
        for (uint32_t j = lane; j < hsize; j += 32) tab[j] = 0;
        __syncwarp();
        for (uint32_t p0 = 0; p0 < np; p0 += 32) {
            uint32_t p = p0 + lane;
            if (p < np) { uint32_t w = ldu32(t.in + p); atomicAdd(&tab[hash3(w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, shift, mask)], 1u); }
        }
        __syncwarp();
        uint32_t run = 0;   // exclusive scan of the bucket sizes
        for (uint32_t j0 = 0; j0 < hsize; j0 += 32) {
            uint32_t v = tab[j0 + lane], tot, ex = warp_excl_scan(v, tot);
            tab[j0 + lane] = run + ex; run += tot;
        }
        __syncwarp();
        for (uint32_t p0 = 0; p0 < np; p0 += 32) {   // stable fill, 32 positions per step in position order
            uint32_t p = p0 + lane; bool ok = p < np; uint32_t h = 0xffffffffu;
            if (ok) { uint32_t w = ldu32(t.in + p); h = hash3(w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, shift, mask); }
            uint32_t peers = __match_any_sync(FULL, h);
            if (ok) {
                uint32_t before = __popc(peers & ((1u << lane) - 1));
                uint32_t cur = tab[h];                      // same value for all peers (read before any update)
                uint32_t sl = cur + before;
                // rank inside the bucket = slot - bucket start; recover bucket start lazily: cnt counts earlier peers
                t.list[sl] = p; t.idx[p] = sl;
                __syncwarp(peers);
                if (before == 0) tab[h] = cur + __popc(peers);
            }
            __syncwarp();
        }
        // cnt[p] = idx[p] - (slot of the first entry of p's bucket): second pass over the list, bucket by bucket
        __syncwarp();
        for (uint32_t p0 = 0; p0 < np; p0 += 32) {
            uint32_t p = p0 + lane;
            if (p < np) {
                uint32_t w = ldu32(t.in + p); uint32_t h = hash3(w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, shift, mask);
                // after the fill tab[h] is the END of bucket h; its start is the end of bucket h-1 (or 0)
                uint32_t start = h ? tab[h - 1] : 0;
                uint32_t r = t.idx[p] - start;
                t.cnt[p] = (uint16_t)(r > 65535u ? 65535u : r);
            }
        }
        __syncwarp();
    }
}

cudaError_t launch_build_chains(const ChainTask *tasks, uint32_t ntasks, uint32_t *queue, uint32_t *tab_all, int ctas, int warps_per_cta, cudaStream_t s) {
    build_chains_kernel<<<ctas, warps_per_cta * 32, 0, s>>>(tasks, ntasks, queue, tab_all);
    return cudaGetLastError();
}

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
