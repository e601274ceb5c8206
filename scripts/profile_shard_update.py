"""ncu target: a few REAL rank-1 pivots on one shard of BASELINE config 5 (131072 rows x C stored columns) driven alone,
for the DRAM traffic of the update kernel on the shard shapes of 2 / 4 / 8 GPUs.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_update -s 2 -c 1 \
        python scripts/profile_shard_update.py 16384
"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
from simplex_solver_b200.sharded import row_stride

R = 131072
C = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ld = row_stride(C)
s = native.Solver(0)
T = torch.empty(R * ld, dtype=torch.float64, device="cuda:0")
s.attach(T.data_ptr(), R - 1, 1, C, ld, R - 1, 2 * R - 2, keep=T)
s.generate(4, R - 1, 0)
r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=4, check_every=4, loop_mode=native.LOOP_LAUNCHES))
print(C, r["n_pivots"], r["device_ms"])
