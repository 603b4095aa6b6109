"""cProfile of a whole mixed_fit on config 3 (host-side share of the fit)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import glmmadaptive_b200 as gb
from glmmadaptive_b200 import synth
d, th = synth.make(sys.argv[1] if len(sys.argv) > 1 else "C3", n=1_000_000)
fam = d.pop("family")
eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, nAGQ=th["nAGQ"])
inits = gb.starting_values_device(eng, False)
gb.mixed_fit(eng, dict(inits), {}, hessian=True)          # warm-up
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
fit = gb.mixed_fit(eng, dict(inits), {}, hessian=True)
pr.disable()
print("fit seconds", time.perf_counter() - t0, fit.get("phases_s"))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
