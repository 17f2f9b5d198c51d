"""Speed of the oracle port (oracle/ch_oracle.py) against the UNMODIFIED reference (run live through
oracle/ref_shim.py) on one host core: same N=512 default configuration, interleaved repeats.  Build container
only (/root/reference does not exist on the GPU box).  Writes profiles/port_over_reference.json, which
bench.py reports as cpu_baseline.port_over_reference."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from threadpoolctl import threadpool_limits
import ch_oracle as orc
import ref_shim

STEPS, WARM, REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 120, 20, int(sys.argv[2]) if len(sys.argv) > 2 else 4
KAPPA = 2.989112919661156e-4


def ref_rate():
    ch = ref_shim.import_reference()
    p = ch.Parameters()
    p.no_gui, p.full_sim, p.N, p.ntmax, p.kappa_tilde = True, True, 512, 10 ** 6, KAPPA
    s = ch.Solver(p)
    s.prepare()
    s.solve_or_resume(1 + WARM)
    t = time.perf_counter()
    s.solve_or_resume(STEPS)
    return STEPS / (time.perf_counter() - t)


def port_rate():
    k = orc.Consts.from_params(N=512, A0=orc.redlich_kister_A0(923.15), A1=orc.redlich_kister_A1(923.15), kappa_tilde=KAPPA)
    U0, draw = orc.initial_field(512, 0.875, "uniform", 2023)
    s = orc.OracleSolver(k, U0, full_sim=True, create_rand=draw)
    s.prepare()
    s.run(1 + WARM)
    t = time.perf_counter()
    s.run(STEPS)
    return STEPS / (time.perf_counter() - t)


with threadpool_limits(limits=1):
    ref, port = [], []
    for _ in range(REPS):
        ref.append(ref_rate())
        port.append(port_rate())
out = {"config": f"N=512 defaults, full_sim, {STEPS} steps after {WARM} warm-up, 1 thread, {REPS} interleaved repeats",
       "reference_steps_per_s": [round(x, 2) for x in ref], "port_steps_per_s": [round(x, 2) for x in port],
       "reference_median": round(float(np.median(ref)), 2), "port_median": round(float(np.median(port)), 2),
       "port_over_reference": round(float(np.median(port) / np.median(ref)), 4),
       "host": f"{os.cpu_count()} logical cores (build container)", "numpy": np.__version__}
json.dump(out, open(os.path.join(ROOT, "profiles", "port_over_reference.json"), "w"), indent=1)
print(json.dumps(out))
