"""CPU model of "re-entering the hypothesis" for deflate_fast trials (no GPU, no product code).
A trial at level T on a stream made at level O != T leaves the original's parse after some tokens; today it then walks bucket
lists for the rest of the block.  Claim tested here: the row of a position p built under the hypothesis (chain = bucket of p
filtered by the positions the ORIGINAL's tokens insert under level T's rule) still gives the trial's own longest_match at p
unless some position whose inserted state differs between hypothesis and trial shares p's hash bucket.  The model runs zlib's
deflate_fast (restated below, checked against the reference zlib's real token streams) with both chains side by side and
reports how many of the trial's tokens could be taken from rows."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(_R, "tests")); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tools"))
import corpus, zref
from dev_token_sync import token_starts

CFG = {1: (4, 8, 4), 2: (5, 16, 8), 3: (6, 32, 32)}     # max_insert, nice_match, max_chain (Z/deflate.c:131-143)
MINM, MAXM, MIN_LOOK = 3, 258, 262


def hash_of(d, p, bits):
    sh = (bits + 2) // 3; m = (1 << bits) - 1
    return ((((d[p] << sh) ^ d[p + 1]) << sh) ^ d[p + 2]) & m


def longest(d, n, p, chain, maxd, nice, budget):
    """zlib longest_match (Z/deflate.c:1148-1289) for deflate_fast (prev_length = 2) over `chain` = candidate positions, most recent first"""
    look = n - p; best = 2; start = 0
    nice = min(nice, look); limit = p - maxd if p > maxd else 0
    left = budget
    for q in chain:
        if q <= limit: break
        if d[q + best] == d[p + best] and d[q + best - 1] == d[p + best - 1] and d[q] == d[p] and d[q + 1] == d[p + 1]:
            l = 2
            mx = min(MAXM, look)
            while l < mx and d[p + l] == d[q + l]: l += 1
            if l > best:
                best = l; start = q
                if l >= nice: break
        left -= 1
        if left == 0: break
    return (best if best <= look else look), start


def parse_fast(d, level, wbits=15, memlevel=8, hyp=None):
    """deflate_fast (Z/deflate.c:1628-1722) on a plaintext shorter than wsize + MAX_DIST (no window slide).  hyp = (token starts of the
    original as a set, the positions its tokens insert under this level's rule): the model's second chain and the bookkeeping"""
    n = len(d); bits = memlevel + 7; maxd = (1 << wbits) - MIN_LOOK
    max_insert, nice, budget = CFG[level]
    d = bytes(d) + bytes(MAXM + 4)
    buckets = {}; hbuckets = {}                     # hash -> inserted positions, oldest first (actual / hypothesis)
    inserted = set(); dirty = set()
    tokens = []; stats = dict(tokens=0, on_boundary=0, clean=0, row_ok=0, row_wrong_when_clean=0)
    if hyp:
        ostarts, oins = hyp
        for q in sorted(oins):
            if q + 2 < n: hbuckets.setdefault(hash_of(d, q, bits), []).append(q)
    p = 0; mlen = 0
    def insert(q):
        h = hash_of(d, q, bits); buckets.setdefault(h, []).append(q); inserted.add(q)
        if hyp and q not in oins: dirty.add(h)
    def skipped(q):                                  # a position the hypothesis inserts and the trial does not
        if hyp and q in oins and q + 2 < n: dirty.add(hash_of(d, q, bits))
    while p < n:
        look = n - p; head = None; mstart = 0
        if look >= MINM:
            h = hash_of(d, p, bits)
            b = buckets.get(h)
            head = b[-1] if b else None
            if hyp:
                stats["tokens"] += 1
                if p in ostarts:
                    stats["on_boundary"] += 1
                    hb = [q for q in hbuckets.get(h, []) if q < p]
                    rl, rs = (longest(d, n, p, reversed(hb), maxd, nice, budget) if hb and hb[-1] != 0 and p - hb[-1] <= maxd else (mlen if mlen < MINM else 0, 0))
            insert(p)
            if head is not None and head != 0 and p - head <= maxd:
                mlen, mstart = longest(d, n, p, reversed(b[:-1]), maxd, nice, budget)
            if hyp and p in ostarts:
                same = (rl >= MINM) == (mlen >= MINM) and (mlen < MINM or (rl, rs) == (mlen, mstart))
                if h not in dirty:      # (p itself is inserted by both sides: it is a token start of both)
                    stats["clean"] += 1
                    if same: stats["row_ok"] += 1
                    else: stats["row_wrong_when_clean"] += 1
        if mlen >= MINM:
            tokens.append((p, mlen, p - mstart)); look -= mlen
            if mlen <= max_insert and look >= MINM:
                for q in range(p + 1, p + mlen): insert(q)
            else:
                for q in range(p + 1, p + mlen): skipped(q)
            p += mlen; mlen = 0
        else:
            tokens.append((p, 1, 0)); p += 1
    return tokens, stats


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    for name, plain in (("text", corpus.text(n, 5)), ("binaryish", corpus.binaryish(n, 6))):
        real = {}
        for lvl in (1, 2, 3):
            real[lvl] = token_starts(zref.ref_deflate(plain, lvl, 15, 8))[0]
            mine, _ = parse_fast(plain, lvl)
            assert [t[0] for t in mine] == real[lvl], f"the model's deflate_fast differs from zlib at level {lvl}"
        print(f"{name}: the model's deflate_fast reproduces zlib's token starts at levels 1-3 ({len(plain)} bytes)")
        for orig, trial in ((2, 3), (3, 2), (1, 2), (2, 1)):
            otok, _ = parse_fast(plain, orig)
            ostarts = set(t[0] for t in otok)
            max_insert = CFG[trial][0]
            oins = set()
            for (s, l, dist) in otok:
                oins.add(s)
                if l >= MINM and l <= max_insert and len(plain) - (s + l) >= MINM: oins.update(range(s + 1, s + l))
            _, st = parse_fast(plain, trial, hyp=(ostarts, oins))
            print(f"  original level {orig}, trial level {trial}: {st['tokens']} tokens; {100 * st['on_boundary'] / st['tokens']:.1f} % on a boundary of the original (a row exists); "
                  f"{100 * st['clean'] / st['tokens']:.1f} % with a clean bucket as well; of those the hypothesis row gives the trial's match in {st['row_ok']} cases and not in {st['row_wrong_when_clean']}")


if __name__ == "__main__":
    main()
