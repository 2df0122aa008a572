"""How far apart are the parses of neighbouring deflate_fast levels?  (CPU only; input to the design of the deflate_fast walk,
DESIGN.md section 8.)  Decodes zlib streams made by the reference zlib at levels 1-3 of the same plaintext and compares their
token boundaries: which share of a trial's tokens starts on a boundary of the "original", and how often the original's next
boundary after a trial token start is where the trial's next token starts (the prediction a preloading walk would use)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(_R, "tests")); sys.path.insert(0, _R)
import corpus, zref

LBASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEXT = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
DBASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577]
DEXT = [0, 0, 0, 0] + [i // 2 for i in range(2, 28)]


class Bits:
    def __init__(self, data): self.d = data; self.p = 0
    def get(self, n):
        v = 0
        for i in range(n):
            v |= ((self.d[self.p >> 3] >> (self.p & 7)) & 1) << i; self.p += 1
        return v


def build(lens):
    cnt = [0] * 16
    for l in lens: cnt[l] += 1
    cnt[0] = 0; code = 0; nxt = [0] * 16
    for b in range(1, 16): code = (code + cnt[b - 1]) << 1; nxt[b] = code
    table = {}
    for s, l in enumerate(lens):
        if l: table[(l, nxt[l])] = s; nxt[l] += 1
    return table


def sym(b, t):
    code = 0
    for l in range(1, 16):
        code = (code << 1) | b.get(1)
        if (l, code) in t: return t[(l, code)]
    raise ValueError("bad code")


def token_starts(z):
    """positions (in the plaintext) where the tokens of a zlib stream start, and how many of them are literals"""
    b = Bits(z[2:]); pos = 0; starts = []; lits = 0
    while True:
        last = b.get(1); typ = b.get(2)
        if typ == 0:
            b.p = (b.p + 7) & ~7; n = b.get(16); b.get(16); b.p += 8 * n
            starts.extend(range(pos, pos + n)); lits += n; pos += n
        else:
            if typ == 1:
                lt = build([8] * 144 + [9] * 112 + [7] * 24 + [8] * 8); dt = build([5] * 30)
            else:
                hl = b.get(5) + 257; hd = b.get(5) + 1; hc = b.get(4) + 4
                cl = [0] * 19
                for i in range(hc): cl[[16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15][i]] = b.get(3)
                ct = build(cl); lens = []
                while len(lens) < hl + hd:
                    s = sym(b, ct)
                    if s < 16: lens.append(s)
                    elif s == 16: lens.extend([lens[-1]] * (3 + b.get(2)))
                    elif s == 17: lens.extend([0] * (3 + b.get(3)))
                    else: lens.extend([0] * (11 + b.get(7)))
                lt = build(lens[:hl]); dt = build(lens[hl:])
            while True:
                s = sym(b, lt)
                if s < 256: starts.append(pos); lits += 1; pos += 1
                elif s == 256: break
                else:
                    ln = LBASE[s - 257] + b.get(LEXT[s - 257]); d = sym(b, dt); b.get(DEXT[d])
                    starts.append(pos); pos += ln
        if last: return starts, lits, pos


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    for name, plain in (("text", corpus.text(n, 5)), ("binaryish", corpus.binaryish(n, 6))):
        tok = {}
        for lvl in (1, 2, 3, 4, 6):
            st, lits, total = token_starts(zref.ref_deflate(plain, lvl, 15, 8)); assert total == len(plain)
            tok[lvl] = (st, lits)
            print(f"{name} level {lvl}: {len(st)} tokens, {100 * lits / len(st):.0f} % literals, {len(plain) / len(st):.2f} bytes per token")
        for orig, trial in ((2, 3), (3, 2), (1, 2), (2, 1), (3, 1), (6, 3), (4, 3)):
            so = tok[orig][0]; st = tok[trial][0]; sset = set(so)
            on = sum(1 for p in st if p in sset)
            import bisect
            hit = 0
            for i in range(len(st) - 1):
                j = bisect.bisect_right(so, st[i])
                if j < len(so) and so[j] == st[i + 1]: hit += 1
            nxt1 = sum(1 for i in range(len(st) - 1) if st[i + 1] == st[i] + 1)
            first = next((i for i, (a, b2) in enumerate(zip(so, st)) if a != b2), min(len(so), len(st)))
            print(f"  original level {orig}, trial level {trial}: {100 * on / len(st):.1f} % of the trial's tokens start on a boundary of the original; "
                  f"'next boundary of the original' predicts the trial's next token start {100 * hit / (len(st) - 1):.1f} % of the time "
                  f"('next byte': {100 * nxt1 / (len(st) - 1):.1f} %); the parses agree for the first {first} tokens")


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "coverage"):
    main()


def visited(starts, n):
    """positions where a deflate_slow-style parse calls longest_match, as build_rows_kernel derives them from a token map: token starts
    and the position after a match start"""
    v = set(starts)
    for a, b in zip(starts, starts[1:] + [n]):
        if b - a >= 3: v.add(a + 1)
    return v


def coverage():
    """which share of the positions a deflate_slow trial looks at has a row when rows are built only where the ORIGINAL looked"""
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    for name, plain in (("text", corpus.text(n, 5)), ("binaryish", corpus.binaryish(n, 6))):
        tok = {lvl: token_starts(zref.ref_deflate(plain, lvl, 15, 8))[0] for lvl in (1, 2, 3, 4, 5, 6, 7, 9)}
        for orig, trial in ((6, 6), (6, 4), (6, 5), (4, 5), (5, 4), (2, 4), (2, 5), (3, 5), (9, 7), (7, 9), (1, 6)):
            vo = visited(tok[orig], n); vt = visited(tok[trial], n)
            print(f"{name}: original level {orig}, trial level {trial}: {100 * len(vt & vo) / len(vt):.1f} % of the trial's longest_match positions have a row")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "coverage":
    coverage()
