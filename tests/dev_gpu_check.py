"""Developer smoke run on a GPU box: parity of the CUDA deflate/inflate against the reference zlib 1.2.8."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zref
import antiz_b200 as az


def text(n, seed, nwords=3000):
    r = random.Random(seed)
    words = [''.join(r.choice('abcdefghijklmnopqrstuvwxyz') for _ in range(r.randint(2, 9))) for _ in range(nwords)]
    out = []; l = 0
    while l < n:
        w = r.choice(words) + ('\n' if r.random() < 0.08 else ' '); out.append(w); l += len(w)
    return ''.join(out).encode()[:n]


def main():
    R = random.Random(3)
    ctx = az.Context(0)
    print(az.lib().atz_version().decode())
    datas = {'text70k': text(70000, 2), 'text3k': text(3000, 3), 'zeros': bytes(100000), 'rand': R.randbytes(50000),
             'short': b'ab', 'empty': b'', 'one': b'x', 'asd': b'asd' * 4608,
             'mixed': text(20000, 5) + R.randbytes(3000) + bytes(5000) + text(40000, 5) + text(40000, 6),
             'bin': bytes((i * i >> 3) & 0xff for i in range(90000))}
    which = sys.argv[1] if len(sys.argv) > 1 else 'quick'
    bad = 0; n = 0; t0 = time.time()
    for name, d in datas.items():
        items = []
        for lvl in range(10):
            for w in (range(10, 16) if which == 'full' else (10, 15)):
                for m in (range(1, 10) if which == 'full' else (1, 8, 9)):
                    items.append((d, lvl, w, m))
        outs = ctx.deflate_batch(items)
        for (dd, lvl, w, m), o in zip(items, outs):
            exp = zref.ref_deflate(dd, lvl, w, m); n += 1
            if o != exp:
                bad += 1
                if bad < 15:
                    k = next((i for i in range(min(len(o), len(exp))) if o[i] != exp[i]), -1)
                    print('DEFLATE MISMATCH', name, lvl, w, m, 'len', len(o), len(exp), 'first diff', k)
        print(name, 'done', n, 'bad', bad, 'elapsed %.1f' % (time.time() - t0), flush=True)
    st = ctx.stats()
    print('deflate cases', n, 'bad', bad, 'ms_trials', st.ms_trials, 'ms_chains', st.ms_chains)
    # inflate
    ibad = 0
    for name, d in datas.items():
        for lvl in (0, 1, 6, 9):
            z = zref.ref_deflate(d, lvl, 15, 8)
            rc, out, used = ctx.inflate_stream(z + b'trailing', len(d) + 16)
            if rc != 0 or out != d or used != len(z):
                ibad += 1; print('INFLATE MISMATCH', name, lvl, rc, len(out), len(d), used, len(z))
    print('inflate bad', ibad)
    return 1 if (bad or ibad) else 0


if __name__ == '__main__':
    sys.exit(main())
