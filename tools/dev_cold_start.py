"""Cold-start cost of the C ABI in a fresh process (what a one-shot `uncomp` run pays): context creation, first load/scan/search
(device allocations included), second pass for comparison."""
import os, sys, time
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(_R, "tests")); sys.path.insert(0, _R)
import antiz_b200 as az
data = open(sys.argv[1], "rb").read()
t = time.perf_counter(); ctx = az.Context(0); t1 = time.perf_counter(); print(f"ctx_create {1e3*(t1-t):.0f} ms")
for it in range(2):
    t0 = time.perf_counter(); ctx.load(data); t1 = time.perf_counter(); ctx.scan(); t2 = time.perf_counter(); ctx.search(az.Options()); t3 = time.perf_counter()
    p = ctx.inflated_recomp(); t4 = time.perf_counter()
    print(f"pass {it}: load {1e3*(t1-t0):.0f} scan {1e3*(t2-t1):.0f} search {1e3*(t3-t2):.0f} payload {1e3*(t4-t3):.0f} ms ({len(p)} B)")
