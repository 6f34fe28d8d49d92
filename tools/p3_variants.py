"""Times the passes for each library variant under build/variants (one subprocess per variant: the library path is
read at import) and prints a digest of the results so that variants can be checked for identical output.
usage: python tools/p3_variants.py name[:ENV=VAL,...] ...      (run on the GPU box)"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, json, hashlib, numpy as np
sys.path.insert(0, %r)
from kbbq_b200 import api, synth
from kbbq_b200.pipeline import Recalibrator
import os
G, COV, L = int(os.environ.get("VAR_GENOME", 4_600_000)), int(os.environ.get("VAR_COVERAGE", 50)), 150
rc = Recalibrator(32, 2, L, G, COV, 2854027472)
sp = api.SynthParams(seed=1000, genome_seed=999, genome_len=G, read_len=L, n_rg=2, err_scale=1.5)
rc.add_resident(rc.ctx.synth_batch(sp, 0, G * COV // L))
best = {}
for it in range(int(os.environ.get("VAR_ITERS", 5))):
    rc.reset(); st = rc.run(fetch_quals=False)
    for k, v in list(st.pass_ms.items()) + list(st.kernel_ms.items()):
        best[k] = round(min(best.get(k, 1e9), v), 3)
cnt = rc.ctx.covariates()
h = hashlib.sha1()
for a in cnt: h.update(np.ascontiguousarray(a).tobytes())
print(json.dumps({"best_ms": best, "counts_sha1": h.hexdigest()[:12], "trusted": int(st.trusted_inserted)}))
'''


def main():
    for spec in sys.argv[1:]:
        name, _, envs = spec.partition(":")
        env = dict(os.environ)
        env["KBBQ_B200_LIB"] = os.path.join(ROOT, "build", "variants", name + ".so")
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        r = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED: " + r.stderr[-400:]
        print("%-28s %s" % (spec, line), flush=True)


if __name__ == "__main__":
    main()
