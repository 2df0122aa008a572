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
// which K3 reads 32 entries at a time.  One CTA builds one (plaintext, hash_bits) task: an LSD radix sort of the
// positions by hash, one or two stable 8-bit counting passes over 256-entry tiles (ranks inside a tile from
// __match_any_sync and per-warp digit counts in shared memory).
#include "common.cuh"

namespace atz {

struct ChainTask {
    const uint8_t *in; uint32_t n; uint32_t hbits;
    uint32_t *list; uint32_t *idx; uint16_t *cnt;
};

#define CH_THREADS 256
#define CH_WARPS (CH_THREADS / 32)

struct ChainSmem {
    uint32_t base[2][256];            // digit histograms, then running output cursors (low digit, high digit)
    uint32_t start0[256];             // bucket starts when the hash has <= 8 bits
    uint16_t wcnt[CH_WARPS][256];     // per-warp digit counts of the current tile
    uint32_t hs[CH_THREADS];          // hashes of the current tile (cnt sweep)
    uint32_t wtot[CH_WARPS];
    uint32_t task, carry, lasth;
};

__device__ __forceinline__ uint32_t hash_at(const uint8_t *in, uint32_t p, uint32_t shift, uint32_t mask) {
    const uint32_t w = ldu32(in + p);
    return hash3(w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, shift, mask);
}
// exclusive scan of 256 values, one per thread
__device__ __forceinline__ uint32_t cta_excl_scan256(uint32_t v, ChainSmem &sm) {
    uint32_t tot, ex = warp_excl_scan(v, tot);
    if (lane_id() == 31) sm.wtot[threadIdx.x >> 5] = tot;
    __syncthreads();
    uint32_t off = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) off += sm.wtot[w];
    __syncthreads();
    return off + ex;
}
// One stable counting-sort pass over np entries by an 8-bit digit of the hash.  Tiles of 256 entries in order; inside a
// tile the rank of an entry among equal digits = (entries of earlier warps) + (earlier lanes of its own warp).
// Positions travel together with their 16-bit hash (srch/dsth), so no pass ever gathers from the plaintext again.
template <bool HI, bool FINAL>
__device__ __forceinline__ void chain_pass(const ChainTask &t, uint32_t np, uint32_t shift, uint32_t mask, const uint32_t *src, const uint16_t *srch,
                                           uint32_t *dst, uint16_t *dsth, ChainSmem &sm) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *base = sm.base[HI ? 1 : 0];
    for (uint32_t s0 = 0; s0 < np; s0 += CH_THREADS) {
        const uint32_t s = s0 + tid; const bool ok = s < np;
        const uint32_t p = ok ? (src ? src[s] : s) : 0;
        const uint32_t h = ok ? (srch ? (uint32_t)srch[s] : hash_at(t.in, p, shift, mask)) : 0;
        const uint32_t d = ok ? (HI ? h >> 8 : h & 255u) : 0xffffffffu;
        { uint32_t *z = (uint32_t *)sm.wcnt; for (uint32_t k = tid; k < CH_WARPS * 128; k += CH_THREADS) z[k] = 0; }
        __syncthreads();
        const uint32_t peers = __match_any_sync(FULL, d), rank = __popc(peers & ((1u << lane) - 1));
        if (ok && rank == 0) sm.wcnt[warp][d] = (uint16_t)__popc(peers);
        __syncthreads();
        uint32_t dest = 0;
        if (ok) { uint32_t off = 0; for (uint32_t w = 0; w < warp; w++) off += sm.wcnt[w][d]; dest = base[d] + off + rank; }
        __syncthreads();
        { uint32_t tot = 0; for (uint32_t w = 0; w < CH_WARPS; w++) tot += sm.wcnt[w][tid]; base[tid] += tot; }
        if (ok) {
            dst[dest] = p;
            if (dsth) dsth[dest] = (uint16_t)h;
            if (FINAL) { t.idx[p] = dest; if (!HI) { uint32_t r = dest - sm.start0[d]; t.cnt[p] = (uint16_t)(r > 65535u ? 65535u : r); } }
        }
        __syncthreads();
    }
}

// One CTA per (plaintext, hash_bits) task: LSD radix sort of the positions by hash (1 or 2 stable 8-bit passes), which
// leaves them ordered by (hash, position).  tmp: per-CTA scratch, np x (u32 position + u16 hash) twice.
__global__ void __launch_bounds__(CH_THREADS) build_chains_kernel(const ChainTask *tasks, uint32_t ntasks, uint32_t *queue, uint32_t *tmp_all, uint64_t tmp_stride) {
    __shared__ ChainSmem sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *tmp = tmp_all + (size_t)blockIdx.x * tmp_stride * 2;                    // positions after pass 1
    uint16_t *tmph = (uint16_t *)(tmp + tmp_stride), *lsth = tmph + tmp_stride;       // their hashes; hashes along the final list
    for (;;) {
        if (tid == 0) sm.task = atomicAdd(queue, 1u);
        __syncthreads();
        const uint32_t ti = sm.task;
        __syncthreads();
        if (ti >= ntasks) break;
        const ChainTask t = tasks[ti];
        const uint32_t np = t.n >= 3 ? t.n - 2 : 0, mask = (1u << t.hbits) - 1, shift = (t.hbits + 2) / 3;
        const bool two = t.hbits > 8;
        sm.base[0][tid] = 0; sm.base[1][tid] = 0;
        __syncthreads();
        for (uint32_t p = tid; p < np; p += CH_THREADS) {
            const uint32_t h = hash_at(t.in, p, shift, mask);
            atomicAdd(&sm.base[0][h & 255u], 1u);
            if (two) atomicAdd(&sm.base[1][h >> 8], 1u);
        }
        __syncthreads();
        { uint32_t v0 = sm.base[0][tid], v1 = sm.base[1][tid];
          uint32_t e0 = cta_excl_scan256(v0, sm), e1 = cta_excl_scan256(v1, sm);
          sm.base[0][tid] = e0; sm.base[1][tid] = e1; sm.start0[tid] = e0; }
        __syncthreads();
        if (!two) chain_pass<false, true>(t, np, shift, mask, nullptr, nullptr, t.list, nullptr, sm);
        else {
            chain_pass<false, false>(t, np, shift, mask, nullptr, nullptr, tmp, tmph, sm);
            chain_pass<true, true>(t, np, shift, mask, tmp, tmph, t.list, lsth, sm);
            // cnt[p] = slot - (first slot of p's bucket): running maximum of the bucket boundaries along the sorted list
            if (tid == 0) { sm.carry = 0; sm.lasth = 0xffffffffu; }
            __syncthreads();
            for (uint32_t s0 = 0; s0 < np; s0 += CH_THREADS) {
                const uint32_t s = s0 + tid; const bool ok = s < np;
                const uint32_t p = ok ? t.list[s] : 0, h = ok ? (uint32_t)lsth[s] : 0xfffffffeu;
                sm.hs[tid] = h;
                __syncthreads();
                const uint32_t hprev = tid ? sm.hs[tid - 1] : sm.lasth;
                uint32_t v = (ok && h != hprev) ? s : 0u;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) { uint32_t y = __shfl_up_sync(FULL, v, dd); if (lane >= (uint32_t)dd && y > v) v = y; }
                if (lane == 31) sm.wtot[warp] = v;
                __syncthreads();
                uint32_t m = sm.carry; for (uint32_t w = 0; w < warp; w++) m = sm.wtot[w] > m ? sm.wtot[w] : m;
                if (v > m) m = v;
                if (ok) { uint32_t r = s - m; t.cnt[p] = (uint16_t)(r > 65535u ? 65535u : r); }
                __syncthreads();
                if (tid == CH_THREADS - 1) { sm.carry = m; sm.lasth = h; }
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_build_chains(const ChainTask *tasks, uint32_t ntasks, uint32_t *queue, uint32_t *tmp_all, uint64_t tmp_stride, int ctas, cudaStream_t s) {
    build_chains_kernel<<<ctas, CH_THREADS, 0, s>>>(tasks, ntasks, queue, tmp_all, tmp_stride);
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
