"""Regenerates the CLI golden files from the UNMODIFIED reference CLI (oracle/_ref/kbbq_ref: the reference's main()
built against oracle/shim with time()/getpid() pinned, so its seed is always 2854027472).

    python tests/golden/make_cli_golden.py          # needs /root/reference (make -C oracle ref)

FASTQ: cli_in.fq.gz (+ cli_fixed.fq.gz, the same reads before error injection) -> cli_out.fq.gz, cli_fixed_out.fq.gz
BAM:   cli_in.bam (2 read groups, every 3rd record reverse strand, OQ tags present) ->
       cli_bam_out.bam.gz        kbbq -k 31 -c 25 --set-oq            (genome length from the header)
       cli_bam_useoq_out.bam.gz  kbbq -k 31 -c 25 -g 4000 --use-oq
Outputs are stored as plain gzip of the uncompressed stream (the shim writes gzip, the product writes BGZF; the
tests compare the uncompressed bytes).
"""
import gzip
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from kbbq_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")


def ref(args, out):
    tmp = out + ".tmp"
    with open(tmp, "wb") as fo:
        subprocess.run([REF] + args, stdout=fo, stderr=subprocess.DEVNULL, check=True)
    data = gzip.open(tmp).read()
    os.remove(tmp)
    with gzip.GzipFile(out, "wb", mtime=0) as g:
        g.write(data)
    return data


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    fq, fx = "/tmp/cli_in.fq", "/tmp/cli_fixed.fq"
    for src, dst in (("cli_in.fq.gz", fq), ("cli_fixed.fq.gz", fx)):
        open(dst, "wb").write(gzip.open(os.path.join(HERE, src)).read())
    ref(["-k", "31", "-g", "4000", "-c", "25", fq], os.path.join(HERE, "cli_out.fq.gz"))
    ref(["-k", "31", "-f", fx, fq], os.path.join(HERE, "cli_fixed_out.fq.gz"))
    # BAM: the same shape of input as the FASTQ case
    rs = synth.make_reads(4000, 25, 150, seed=11, n_rg=2, n_frac=0.002)
    bam = os.path.join(HERE, "cli_in.bam")
    synth.to_bam(rs, bam, genome_len=4000, rev_every=3, with_oq=True)
    ref(["-k", "31", "-c", "25", "--set-oq", bam], os.path.join(HERE, "cli_bam_out.bam.gz"))
    ref(["-k", "31", "-c", "25", "-g", "4000", "--use-oq", bam], os.path.join(HERE, "cli_bam_useoq_out.bam.gz"))
    # a BAM without OQ tags: --set-oq appends the tag (bam_aux_update_str, htsiter.cc:13-26).  The reference fetches
    # bam_get_qual() before that call and writes through it afterwards, so its output is only defined while htslib does
    # not have to reallocate the record: record 0's long name sizes the (reused) buffer for all later appends.
    bam2 = os.path.join(HERE, "cli_in_nooq.bam")
    synth.to_bam(rs, bam2, genome_len=4000, rev_every=4, with_oq=False, first_name=b"r0_" + b"x" * 40)
    ref(["-k", "31", "-c", "25", "--set-oq", bam2], os.path.join(HERE, "cli_bam_nooq_out.bam.gz"))


if __name__ == "__main__":
    main()
