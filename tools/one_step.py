"""Two whole recalibrations of the bench workload (config-2 shape) for profiling: ncu skips the first (6 launches) and
captures the second.  usage: python tools/one_step.py [n_iters]"""
import sys
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from kbbq_b200 import api, synth
from kbbq_b200.pipeline import Recalibrator

G, COV, L = 4_600_000, 50, 150
rs = synth.make_reads(G, COV, L, seed=1000, n_rg=2)
pb = api.PackedBatch(rs, pinned=True)
rc = Recalibrator(32, 2, L, G, COV, 2854027472)
rc.add_batch(pb)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    rc.reset()
    st = rc.run(fetch_quals=False)
    print(it, {k: round(v, 3) for k, v in st.pass_ms.items()}, flush=True)
