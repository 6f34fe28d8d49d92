"""Wall time and phase times of the kbbq CLI on a config-2-sized FASTQ (and on a tiny one: start-up + exit alone).
usage: python tools/cli_timing.py [runs]   -- measurement helper, prints to stdout"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kbbq_b200 import api  # noqa: E402


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    cli = os.path.join(ROOT, "kbbq_b200", "host", "kbbq")
    G, cov = 4_600_000, 50
    with tempfile.TemporaryDirectory() as d:
        for label, n in (("tiny", 2000), ("config2", G * cov // bench.READ_LEN)):
            rs = api.synth_host(bench.synth_params(G, 1000), 0, n, n_threads=os.cpu_count() or 1)
            fq, out = os.path.join(d, label + ".fq"), os.path.join(d, "out.gz")
            bench.write_fastq(fq, rs)
            variants = ({}, {"KBBQ_CLEAN_EXIT": "1"}, {"KBBQ_BGZF_LEVEL": "huff"})
            if os.environ.get("CLI_TIMING_EXIT_MODES"):  # what the teardown at exit is made of (KBBQ_EXIT_MODE bits: 1 host, 2 contexts, 4 reset)
                variants = ({},) + tuple({"KBBQ_EXIT_MODE": m} for m in os.environ["CLI_TIMING_EXIT_MODES"].split(","))
            for extra in variants:
                for i in range(runs):
                    env = dict(os.environ, KBBQ_HOST_TIMING="1", **extra)
                    if "KBBQ_BGZF_LEVEL" not in extra:
                        env.pop("KBBQ_BGZF_LEVEL", None)
                    t0 = time.perf_counter()
                    with open(out, "wb") as fo:
                        r = subprocess.run([cli, "--seed", "7", "-k", "32", "-g", str(G), "-c", str(cov), fq], stdout=fo, stderr=subprocess.PIPE, env=env)
                    dt = time.perf_counter() - t0
                    print(f"== {label} {extra} run {i}: rc={r.returncode} wall {dt * 1e3:.1f} ms, out {os.path.getsize(out)} bytes")
                    if i == runs - 1:
                        print("\n".join(ln for ln in r.stderr.decode(errors="replace").splitlines() if ln.startswith("[")))


if __name__ == "__main__":
    main()
