"""Generates tests/golden/nedelec_projection_digest.json from the reference's own golden
output (run in the build container where /root/reference exists):

    python tests/golden/make_nedelec_digest.py

The reference test test/test_fe_projection_nedelec_mpi.cc L2-projects the constant field
(1,2,3) onto FE_Nedelec<3>(0) on an 8^3 mesh and prints rank 0's coefficients; deal.II's
harness diffs them against test_fe_projection_nedelec_mpi.mpirun=1.output.  DoF numbering
is deal.II-internal, so the digest keeps what is numbering independent: the sorted
multiset of values."""
import collections
import json
import os
import re

SRC = "/root/reference/test/test_fe_projection_nedelec_mpi.mpirun=1.output"
vals = []
for line in open(SRC):
    line = line.strip()
    if re.fullmatch(r"[-+0-9.eE]+", line):
        vals.append(float(line))
cnt = collections.Counter(vals)
out = {"source": "test/test_fe_projection_nedelec_mpi.mpirun=1.output", "mesh": "hyper_cube refined 3x (8^3)",
       "field": [1, 2, 3], "n_values": len(vals), "histogram": {repr(k): v for k, v in sorted(cnt.items())}}
path = os.path.join(os.path.dirname(__file__), "nedelec_projection_digest.json")
json.dump(out, open(path, "w"), indent=1)
print(out)
