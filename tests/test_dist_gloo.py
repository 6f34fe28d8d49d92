"""The N>1 path on CPU: two gloo ranks shard the reads, the oracle port plays the compute engine, and the merge logic
of kbbq_b200.dist (slot ordinals, OR of filter bits, sums of counters and count tables) must reproduce the
single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from kbbq_b200 import synth
from kbbq_b200.dist import Comm, shard_bounds

G, COV, L, K, SEED = 40000, 20, 100, 32, 4242


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port as P

    comm = Comm(rank, world)
    rs = synth.make_reads(G, COV, L, seed=21, n_rg=2, ragged=True, min_len=70, n_frac=0.002)
    b = shard_bounds(rs.n, world)
    mine = rs.subset(int(b[rank]), int(b[rank + 1]))
    alpha_ld = np.longdouble(7.0) / np.longdouble(COV)
    approx = int(np.longdouble(G) * COV * alpha_ld)
    pp = P.PortPipeline(mine, K)
    sampled, trusted = P.PortBloom(approx, 0.01), P.PortBloom(approx, 0.0005)

    def merge(f):
        words = torch.from_numpy(f.words().view(np.int64))
        comm.or_reduce_(words)
        ins = torch.tensor([f.inserted], dtype=torch.int64)
        comm.sum_reduce_(ins)
        f.set_words(words.numpy().view(np.uint64), int(ins.item()))

    my_slots = int(np.maximum(np.diff(mine.off).astype(np.int64) - K + 1, 0).sum())
    base = comm.exclusive_prefix(my_slots)
    assert pp.subsample(sampled, float(alpha_ld), SEED, skip=base) == my_slots
    merge(sampled)
    th = pp.thresholds(sampled, 7.0, float(COV))
    e2 = pp.infer_errors(sampled, th)
    pp.find_trusted(trusted, sampled, th)
    merge(trusted)
    e3, seq_after = pp.get_errors(trusted, want_seq=True)
    tables = pp.consume(e3, seq_after, maxlen=L)
    flat = torch.from_numpy(np.concatenate([t.reshape(-1) for t in tables]).view(np.int64))
    comm.sum_reduce_(flat)
    flat = flat.numpy().view(np.uint64)
    sizes = [t.size for t in tables]
    merged = [flat[sum(sizes[:i]): sum(sizes[:i + 1])].reshape(tables[i].shape) for i in range(4)]
    dqs = pp.get_dqs(merged)
    newq = pp.recalibrate(dqs)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), sampled=sampled.words(), sampled_ins=sampled.inserted,
             trusted=trusted.words(), trusted_ins=trusted.inserted, th=th, e2=e2, e3=e3, newq=newq,
             rgc=merged[0], qc=merged[1], cyc=merged[2], dic=merged[3], lo=int(mine.off[0]), base=base)
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharded_run_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rs = synth.make_reads(G, COV, L, seed=21, n_rg=2, ragged=True, min_len=70, n_frac=0.002)
    want = common.run_port(rs, K, G, COV, SEED)
    b = shard_bounds(rs.n, world)
    e2, e3, newq = [], [], []
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert np.array_equal(z["sampled"], want["sampled_words"]) and int(z["sampled_ins"]) == want["sampled_inserted"]
        assert np.array_equal(z["trusted"], want["trusted_words"]) and int(z["trusted_ins"]) == want["trusted_inserted"]
        assert np.array_equal(z["th"], want["thresholds"])
        for key in ("rgc", "qc", "dic"):
            assert np.array_equal(z[key], want[key]), key
        assert np.array_equal(common.pad_cycle(z["cyc"], L), common.pad_cycle(want["cyc"], L))
        e2.append(z["e2"]); e3.append(z["e3"]); newq.append(z["newq"])
    assert np.array_equal(np.concatenate(e2), want["pass2_errors"])
    assert np.array_equal(np.concatenate(e3), want["pass3_errors"])
    assert np.array_equal(np.concatenate(newq), want["newqual"])


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000, 1533333):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0) and len(b) == w + 1
