"""Multi-GPU parity check (run under torchrun, one rank per GPU; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank recalibrates its shard on its own GPU; after the three merges every rank must hold the filters, count
tables and model the oracle port computes for the WHOLE input, and its shard's flags/qualities must match."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from kbbq_b200 import api, synth  # noqa: E402
from kbbq_b200.dist import GpuComm, shard_bounds  # noqa: E402
from kbbq_b200.pipeline import Recalibrator  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
    G, cov, L, k, seed = 150000, 30, 150, 32, 777
    rs = synth.make_reads(G, cov, L, seed=41, n_rg=2, ragged=True, min_len=80, n_frac=0.002)
    b = shard_bounds(rs.n, world)
    mine = rs.subset(int(b[rank]), int(b[rank + 1]))
    comm = GpuComm(rank, world, local)
    rc = Recalibrator(k, rs.n_rg, L, G, cov, seed, device=local, comm=comm)
    pb = api.PackedBatch(mine)
    rc.add_batch(pb)
    rc.run(fetch_quals=True, want_errors=True)
    want = common.run_port(rs, k, G, cov, seed)
    ctx = rc.ctx
    ok = True

    def check(name, cond):
        nonlocal ok
        if not cond:
            ok = False
            print(f"[rank {rank}] MISMATCH {name}", flush=True)

    check("sampled words", np.array_equal(ctx.bloom_words(0), want["sampled_words"]))
    check("sampled inserted", ctx.bloom_inserted(0) == want["sampled_inserted"])
    check("trusted words", np.array_equal(ctx.bloom_words(1), want["trusted_words"]))
    check("trusted inserted", ctx.bloom_inserted(1) == want["trusted_inserted"])
    rgc, qc, cyc, dic = ctx.covariates()
    check("rgc", np.array_equal(rgc, want["rgc"]))
    check("qc", np.array_equal(qc, want["qc"]))
    check("cyc", np.array_equal(common.pad_cycle(cyc, L), common.pad_cycle(want["cyc"], L)))
    check("dic", np.array_equal(dic, want["dic"]))
    lo, hi = int(rs.off[b[rank]]), int(rs.off[b[rank + 1]])
    check("pass2 errors", np.array_equal(pb.unpack_bits(rc.last_errors2[0], mine.total_bases), want["pass2_errors"][lo:hi]))
    check("pass3 errors", np.array_equal(pb.unpack_bits(rc.last_errors3[0], mine.total_bases), want["pass3_errors"][lo:hi]))
    check("new quals", np.array_equal(pb.unpack_qual(rc.last_quals[0], mine.total_bases), want["newqual"][lo:hi]))
    t = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU PARITY", "OK" if int(t.item()) == 1 else "FAILED", f"world={world} reads={rs.n} merge={comm.last_merge}", flush=True)
    rc.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
