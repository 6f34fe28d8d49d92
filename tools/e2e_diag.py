import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from kbbq_b200 import api, synth
from kbbq_b200.pipeline import Recalibrator
G, COV, L = 4_600_000, 50, 150
rs = synth.make_reads(G, COV, L, seed=1000, n_rg=2)
t=time.perf_counter(); pb = api.PackedBatch(rs, pinned=True); print("pack+pin", time.perf_counter()-t)
rc = Recalibrator(32, 2, L, G, COV, 2854027472)
qo = torch.zeros(pb.n_words*32, dtype=torch.uint8).pin_memory().numpy()
for it in range(6):
    t0=time.perf_counter(); rc.reset(); rc.ctx.sync(); t1=time.perf_counter()
    rc.add_batch(pb); rc.ctx.sync(); t2=time.perf_counter()
    st = rc.run(fetch_quals=True, qual_bufs=[qo]); t3=time.perf_counter()
    rc.clear_batches(); t4=time.perf_counter()
    print(f"it{it}: reset {1e3*(t1-t0):.1f} upload {1e3*(t2-t1):.1f} run {1e3*(t3-t2):.1f} free {1e3*(t4-t3):.1f} | passes {sum(st.pass_ms.values()):.1f}")
