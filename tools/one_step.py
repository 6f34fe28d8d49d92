"""Whole recalibrations of a synthetic workload for profiling and A/B timing (run on the GPU box).
usage: python tools/one_step.py [n_iters] [genome] [coverage]      default: 2 iterations of the config-2 shape
Prints per-kernel device times per iteration and a digest of the count tables (A/B runs must agree)."""
import hashlib
import sys
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np
from kbbq_b200 import api, synth
from kbbq_b200.pipeline import Recalibrator

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4_600_000
COV = int(sys.argv[3]) if len(sys.argv) > 3 else 50
L = 150
rc = Recalibrator(32, 2, L, G, COV, 2854027472)
sp = api.SynthParams(seed=1000, genome_seed=999, genome_len=G, read_len=L, n_rg=2, err_scale=1.5)
rc.add_resident(rc.ctx.synth_batch(sp, 0, G * COV // L))  # bench.py's reads, made on the device
for it in range(iters):
    rc.reset()
    st = rc.run(fetch_quals=False)
    print(it, {k: round(v, 3) for k, v in st.kernel_ms.items()}, "total %.2f" % sum(st.pass_ms.values()), flush=True)
h = hashlib.sha1()
for a in rc.ctx.covariates():
    h.update(np.ascontiguousarray(a).tobytes())
print("counts sha1", h.hexdigest()[:12], "sampled", st.sampled_inserted, "trusted", st.trusted_inserted)
