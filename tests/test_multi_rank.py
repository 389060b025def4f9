"""world_size-2 gloo test of the N>1 path's host logic (no GPU): stream sharding, per-rank RDS decode
through the product's host decoder, and the result gather.  The CUDA demodulator is replaced by a stand-in
driven by the CPU checker (oracle/ is test infrastructure and may be used here)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as dist
    import fm_radio_b200 as fm
    from fm_radio_b200 import Buf, synth
    from fm_radio_b200.batch import StreamBatch, shard_streams, gather_results
    from oracle import bind

    class CheckerDemod:
        '''process_u8/get stand-in for FMDemod: one CPU checker per local stream.'''
        def __init__(self, n, B):
            self.chk = [bind.CpuDemod(B, "port") for _ in range(n)]
        def process_u8(self, iq):
            for c, row in zip(self.chk, iq):
                c.process_u8(row)
        def get(self, buf, stream=0):
            assert buf == Buf.RDS_PRED_SYM
            return self.chk[stream].get("rds_pred_sym")

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N, B, NB = 5, 16384, 150                      # 5 streams over 2 ranks: 3 + 2
    mine = shard_streams(N, rank, world)
    caps = [synth.synth_u8_numpy(B * NB, synth.StreamParams.for_stream(s)) for s in mine]
    batch = StreamBatch(mine, block_size=B, demod=CheckerDemod(len(mine), B))
    for k in range(NB):
        batch.process_u8(np.stack([c[2 * B * k:2 * B * (k + 1)] for c in caps]))
    res = gather_results(batch.results())
    if rank == 0:
        print("RESULT " + json.dumps([[r[0], r[1], r[2].decode("latin1"), r[4]] for r in res]))
    dist.destroy_process_group()
""") % ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_shard_decode_gather(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert [x[0] for x in res] == [0, 1, 2, 3, 4]                 # every stream exactly once, in order
    from fm_radio_b200 import synth
    for sid, pi, ps, n_groups in res:
        p = synth.StreamParams.for_stream(sid)
        assert pi == p.pi_code, (sid, hex(pi))
        assert ps == p.ps, (sid, ps)
        assert n_groups >= 15
