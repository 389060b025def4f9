"""Multi-stream / multi-GPU batch driver (new component; the reference runs exactly one
Broadcast_FM_Demod per process, src/app.cpp:19).

Streams share nothing (all state is per demodulator object, broadcast_fm_demod.h:143-166), so the
batch shards by stream id across ranks with NO data-path collective: rank r owns a contiguous
range of streams, demodulates them with one FMDemod, decodes RDS on its host cores, and only the
decoded results (a few bytes per stream) are gathered with torch.distributed for reporting.
"""
from __future__ import annotations

import numpy as np

from .api import Buf, FMDemod, RDSDecoder


def shard_streams(n_streams: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced partition of stream ids (sizes differ by at most one)."""
    base, extra = divmod(n_streams, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class StreamBatch:
    """One rank's share of a batch of streams: FMDemod + one host RDS decoder per stream."""

    def __init__(self, stream_ids, block_size: int = 65536, device: int = -1, pipeline_depth: int = 0,
                 keep_intermediates: bool = False, demod=None):
        self.stream_ids = list(stream_ids)
        # `demod` is injectable so the rank/shard/gather logic can be exercised without a GPU (the CPU
        # test-suite passes a stand-in with the same process_u8/get interface); the product always
        # builds the CUDA demodulator.
        self.demod = demod if demod is not None else FMDemod(
            block_size, len(self.stream_ids), device=device,
            keep_intermediates=keep_intermediates, pipeline_depth=pipeline_depth)
        self.rds = [RDSDecoder() for _ in self.stream_ids]
        self.audio = None

    def process_u8(self, iq_u8: np.ndarray, collect_audio: bool = False):
        """iq_u8: [n_local_streams, 2*block_size] host array.  Synchronous: demodulate, then feed
        each stream's soft symbols to its RDS decoder (the reference's OnRDSOut observer,
        src/app.cpp:27-34)."""
        self.demod.process_u8(iq_u8)
        self.drain(collect_audio)

    def drain(self, collect_audio: bool = False):
        for i, dec in enumerate(self.rds):
            dec.push_symbols(self.demod.get(Buf.RDS_PRED_SYM, i))
        if collect_audio:
            self.audio = [self.demod.get(Buf.AUDIO_OUT, i) for i in range(len(self.rds))]

    def results(self):
        """[(stream_id, PI, PS, RT, n_groups)] for this rank's streams."""
        out = []
        for sid, dec in zip(self.stream_ids, self.rds):
            db = dec.db()
            out.append((sid, db["pi"], db["ps"], db["rt"], len(dec.groups()[0])))
        return out


def gather_results(local_results, group=None):
    """All ranks' decoded results on every rank (torch.distributed all_gather_object; gloo or nccl).
    This is reporting, not the data path: a few dozen bytes per stream."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sorted(local_results)
    gathered = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, local_results, group=group)
    return sorted(r for part in gathered for r in part)
