"""Where the end-to-end step of bench.py spends its wall time (run on the GPU box)."""
import sys, time, numpy as np, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from kbbq_b200 import api, synth
from kbbq_b200.pipeline import Recalibrator
G, COV, L = 4_600_000, 50, 150
rs = synth.make_reads(G, COV, L, seed=1000, n_rg=2)
t = time.perf_counter(); pb = api.PackedBatch(rs, pinned=True); print("pack+pin %.1f ms" % (1e3 * (time.perf_counter() - t)))
rc = Recalibrator(32, 2, L, G, COV, 2854027472)
qo = torch.zeros(pb.n_words * 32, dtype=torch.uint8).pin_memory().numpy()
for it in range(6):
    rc.ctx.sync()
    t0 = time.perf_counter(); rc.reset(); t1 = time.perf_counter()
    rc.add_batch(pb); t2 = time.perf_counter()
    st = rc.run(fetch_quals=True, qual_bufs=[qo]); t3 = time.perf_counter()
    rc.clear_batches(); rc.ctx.sync(); t4 = time.perf_counter()
    print(f"it{it}: total {1e3*(t4-t0):.1f} = reset {1e3*(t1-t0):.1f} + upload call {1e3*(t2-t1):.1f} + run {1e3*(t3-t2):.1f} + free {1e3*(t4-t3):.1f} | "
          f"passes {sum(st.pass_ms.values()):.1f} {({k: round(v, 2) for k, v in st.pass_ms.items()})} host {({k: round(v, 2) for k, v in st.host_ms.items()})}")
# raw copy speeds for reference
a = torch.empty(pb.n_words * 32, dtype=torch.uint8, device="cuda")
h = torch.from_numpy(qo)
for name, fn in (("H2D", lambda: a.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(a, non_blocking=True))):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"{name} {a.numel()/1e6:.0f} MB: {1e3*dt:.2f} ms = {a.numel()/dt/1e9:.1f} GB/s")
