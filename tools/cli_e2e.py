"""End-to-end CLI timing on a config-2-shaped input (run on the GPU box):
    python tools/cli_e2e.py [fraction of config 2, default 1.0]
Writes the synthetic reads as FASTQ, gzip FASTQ-less BGZF... and BAM, runs kbbq_b200/host/kbbq on each with
KBBQ_HOST_TIMING=1 and, on a small prefix, the reference CLI for scale."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kbbq_b200 import synth  # noqa: E402

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
G = int(4_600_000 * frac)
rs = synth.make_reads(G, 50, 150, seed=1000, n_rg=2, keep_clean=False)
tmp = os.environ.get("TMPDIR", "/tmp")
fq, bam = os.path.join(tmp, "c2.fq"), os.path.join(tmp, "c2.bam")
t = time.time(); synth.to_fastq(rs, fq); print("wrote %s (%d reads, %.0f MB) in %.1f s" % (fq, rs.n, os.path.getsize(fq) / 1e6, time.time() - t), flush=True)
t = time.time(); synth.to_bam(rs, bam, genome_len=G); print("wrote %s (%.0f MB) in %.1f s" % (bam, os.path.getsize(bam) / 1e6, time.time() - t), flush=True)
cli = os.path.join(ROOT, "kbbq_b200", "host", "kbbq")
env = dict(os.environ, KBBQ_HOST_TIMING="1")
for name, args in (("FASTQ", ["-g", str(G), "-c", "50", fq]), ("BAM --set-oq", ["-c", "50", "--set-oq", bam])):
    for rep in range(2):
        t = time.time()
        with open(os.path.join(tmp, "out.bin"), "wb") as fo:
            r = subprocess.run([cli, "--seed", "2854027472", "-k", "32"] + args, stdout=fo, stderr=subprocess.PIPE, env=env)
        dt = time.time() - t
        print("== %s run %d: rc %d, %.2f s wall, %.1f Mbases/s, output %.0f MB" % (name, rep, r.returncode, dt, rs.total_bases / dt / 1e6, os.path.getsize(os.path.join(tmp, "out.bin")) / 1e6), flush=True)
        if rep == 1 or r.returncode:
            print("\n".join(l for l in r.stderr.decode().splitlines() if l.startswith("[") and ("host]" in l or "cli]" in l)), flush=True)
ref = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")
if os.path.exists(ref):
    small = rs.subset(0, rs.n // 16)
    sfq = os.path.join(tmp, "c2_small.fq")
    synth.to_fastq(small, sfq)
    t = time.time()
    with open(os.path.join(tmp, "ref_out.gz"), "wb") as fo:
        r = subprocess.run([ref, "-k", "32", "-g", str(G // 16), "-c", "50", sfq], stdout=fo, stderr=subprocess.DEVNULL)
    dt = time.time() - t
    print("== reference CLI on 1/16 of the reads: rc %d, %.2f s wall, %.2f Mbases/s" % (r.returncode, dt, small.total_bases / dt / 1e6), flush=True)
