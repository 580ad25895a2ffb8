"""N > 1 plumbing on CPU: world_size-2 gloo, round-robin sharding + the stats all_gather."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from aircraftoptimalcontrol_b200 import dist as D
from tests.util import ROOT


def test_sharding_partitions_instances():
    for n, w in ((10, 2), (11, 4), (7, 8), (1048576, 8)):
        idx = [D.shard_indices(n, r, w) for r in range(w)]
        assert sorted(np.concatenate(idx).tolist()) == list(range(n)) if n < 100 else sum(map(len, idx)) == n
        assert [len(i) for i in idx] == D.shard_sizes(n, w)
        assert max(map(len, idx)) - min(map(len, idx)) <= 1


WORKER = """
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from aircraftoptimalcontrol_b200 import dist as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
n = 11
r = dist.get_rank()
idx = D.shard_indices(n, r, 2)
stats = dict(iters=(idx + 20).astype(np.int32), status=np.ones(len(idx), dtype=np.int32), J=idx * 1.5, descent=-idx * 1e-7)
g = D.gather_stats(stats, n)
assert np.array_equal(g["iters"], np.arange(n) + 20), g["iters"]
assert np.array_equal(g["J"], np.arange(n) * 1.5) and np.all(g["status"] == 1)
s = D.summarize(g)
assert s["total_iters"] == sum(range(20, 31)) and s["converged"] == n
dist.destroy_process_group()
print("ok", r)
"""


def test_gather_stats_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(WORKER % ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), PORT=str(port)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o
