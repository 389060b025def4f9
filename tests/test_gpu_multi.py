"""Multi-GPU leg (-m gpu; skipped on a one-GPU box): BASELINE config 5's wideband path on 2 ranks -- the ingest rank's
u8 block is broadcast with NCCL, each rank channelizes and demodulates its own channels (c mod 2), and every station
must decode its own PI on the device.  Runs bench.py's wideband workload under torchrun, like the driver does."""
import json
import os
import subprocess
import sys

import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_wideband_broadcast_on_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(H.ROOT, "bench.py"), "--workload", "wideband", "--gpus", "2", "--steps", "8",
           "--warmup", "3", "--stations", "12", "--clock-warmup-ms", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, NCCL_DEBUG="WARN"))
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    assert line["config"]["broadcast_bytes_per_step"] == 2 * 65536 * 20
    assert line["rds_check"] == {"stations_with_own_pi_decoded_on_device": 12, "of": 12}
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
