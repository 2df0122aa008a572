import os, sys, time
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(_R, "tests")); sys.path.insert(0, _R)
import antiz_b200 as az, bench
kind, n = sys.argv[1], int(sys.argv[2])
data = bench.make_container(kind, n, 2)
ctx = az.Context(0)
for it in range(4):
    t0 = time.perf_counter(); ctx.load(data); t1 = time.perf_counter(); ctx.scan(); t2 = time.perf_counter(); ctx.search(az.Options()); t3 = time.perf_counter()
    st = ctx.stats()
    print(f"{kind} it{it}: load {1e3*(t1-t0):.1f} scan wall {1e3*(t2-t1):.1f} (gpu {st.ms_scan+st.ms_inflate_probe+st.ms_inflate:.1f}) search wall {1e3*(t3-t2):.1f} (gpu {st.ms_chains+st.ms_rows+st.ms_trials+st.ms_diff:.1f}) streams {st.n_streams} cands {st.n_candidates} launches {st.kernel_launches}")
