import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import glmmadaptive_b200 as gb
from glmmadaptive_b200 import synth
for name in ("C1", "C3"):
    d, th = synth.make(name, n=1_000_000)
    fam = d.pop("family")
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, nAGQ=th["nAGQ"])
    eng.find_modes(th["betas"], np.linalg.inv(th["D"]), th.get("phis"), None, warm=False)
    for rep in range(3):
        t0 = time.perf_counter()
        H = eng.observed_information(th["betas"], th["D"], th.get("phis"), None)
        print(name, "observed_information call", rep, "%.3f s" % (time.perf_counter() - t0))
    eng.close()
