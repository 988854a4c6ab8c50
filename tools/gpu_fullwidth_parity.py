"""GPU box: the parity_full_width leg of bench.py on its own (oracle 1-layer full-width fwd+bwd on the host cores,
then the CUDA path on the same model / batch).  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from oracle import llava_dpo_oracle as O

print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)), 'usable', bench.usable_cores(),
      'picked threads', bench.pick_cpu_threads(), file=sys.stderr)
cfg, p, batch = bench.full_width_case(1)
out = O.dpo_step(p, cfg, batch, beta=0.1)
out["loss"].backward()
if "--cpu-only" in sys.argv:
    print("cpu half ok", out["logp"].tolist(), float(out["loss"]))
    sys.exit(0)
print(json.dumps(bench.gpu_full_width_parity(p, cfg, batch, out)))
