"""Times the global coarse solve of the host driver (host/coarse.cpp) on the host cores and on the GPU (include/msfec_coarse.h):
python profiles/tools/coarse_timing.py <PAIRING> <global refinements> [host|device|both] [dense-LU limit, default 6000].  Element matrices come from the oracle
at 1 local refinement with the rough random field of C5 (cheap; only the coarse system's size and conditioning matter here)."""
import os
import subprocess
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from common import oracle_problem  # noqa: E402
from oracle import msfec_oracle as mo  # noqa: E402

PAIRING = {"Q": 0, "Q_NED": 1, "NED_RT": 2, "RT_DQ": 3}
pairing = sys.argv[1]; g = int(sys.argv[2]); which = sys.argv[3] if len(sys.argv) > 3 else "both"
dense_limit = int(sys.argv[4]) if len(sys.argv) > 4 else 6000
cells = mo.morton_cells(g)
prob = oracle_problem(pairing, 1, random_seed=20261017)


def one(c):
    Mc, rc, *_ = mo.build_basis(prob, cells[c], c)
    return Mc, rc


if __name__ == "__main__":
    t0 = time.time()
    with ProcessPoolExecutor(os.cpu_count()) as ex:
        out = list(ex.map(one, range(len(cells)), chunksize=64))
    print(f"{pairing}, {len(cells)} coarse cells: oracle element matrices in {time.time() - t0:.1f} s", flush=True)
    os.makedirs(os.path.join(ROOT, ".scratch"), exist_ok=True)
    inp = os.path.join(ROOT, ".scratch", "coarse_in.bin")
    with open(inp, "wb") as f:
        f.write(np.array([PAIRING[pairing], g, len(cells), dense_limit], np.int64).tobytes())
        for c in range(len(cells)):
            f.write(np.array([c], np.int64).tobytes()); f.write(out[c][0].tobytes()); f.write(out[c][1].tobytes())
    exe = os.path.join(ROOT, "mpi-msfec_b200", "host", "coarse_test")
    res = {}
    for mode in (("host", "device") if which == "both" else (which,)):
        outp = os.path.join(ROOT, ".scratch", f"coarse_{mode}.bin")
        t0 = time.time()
        r = subprocess.run([exe, inp, outp] + (["0"] if mode == "device" else []), capture_output=True, text=True)
        print(f"  {mode}: {r.stdout.strip()} {r.stderr.strip()} -- wall {time.time() - t0:.2f} s (incl. reading the file and assembly)", flush=True)
        res[mode] = np.fromfile(outp)
    if len(res) == 2:
        print(f"  max |w_device - w_host| / max |w_host| = {np.abs(res['device'] - res['host']).max() / np.abs(res['host']).max():.2e}")
