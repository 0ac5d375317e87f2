"""Diagnostic (not a benchmark): run bench.py under torchrun with the data-path collectives turned into no-ops, to separate
the cost of the exchanged bytes from the cost of the extra data-parallel compute. Results are numerically meaningless."""
import os
import runpy
import sys

import torch.distributed as dist

_noop = lambda *a, **k: None  # noqa: E731
dist.all_reduce = _noop
dist.all_gather_into_tensor = _noop
dist.reduce_scatter_tensor = _noop
sys.argv = ["bench.py"] + sys.argv[1:]
runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"), run_name="__main__")
