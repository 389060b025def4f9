"""Multi-stream / multi-GPU batch driver (new component; the reference runs exactly one
Broadcast_FM_Demod per process, src/app.cpp:19).

Streams share nothing (all state is per demodulator object, broadcast_fm_demod.h:143-166), so the
batch shards by stream id across ranks with NO data-path collective: rank r owns a contiguous
range of streams, demodulates them with one FMDemod, decodes RDS on its host cores, and only the
decoded results (a few bytes per stream) are gathered with torch.distributed for reporting.
"""
from __future__ import annotations

import numpy as np

from .api import Buf, ChanMode, Channelizer, FMDemod, RDSDecoder


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


def shard_channels(n_channels: int, rank: int, world_size: int) -> list[int]:
    """Channels of a wideband capture owned by `rank`: {c : c mod world_size = rank}, so
    neighbouring (equally loaded) channels spread over the ranks."""
    return list(range(rank, n_channels, world_size))


class WidebandReceiver:
    """BASELINE config 4 / 5: one wideband u8 capture -> this rank's channels -> demodulators.

    Every rank needs the SAME wideband block, so the only data-path collective of the whole framework
    is one broadcast of the u8 block per step (`broadcast_and_feed`, NCCL over NVLink on GPUs; gloo in
    the CPU tests); each rank then channelizes and demodulates only the channels it owns -- it
    computes only its own columns of the channelizer's contraction -- and results (PI / PS / RT per
    channel) are gathered on the host with `gather_results`."""

    def __init__(self, fs_in_hz: float, centres_hz, rank: int = 0, world_size: int = 1, block_out: int = 65536,
                 decimation: int = 20, n_taps: int = 192, device: int = -1, mode: ChanMode = ChanMode.AUTO,
                 pipeline_depth: int = 0, chan=None, demod=None):
        centres_hz = np.asarray(centres_hz, np.float64)
        self.channel_ids = shard_channels(len(centres_hz), rank, world_size)
        self.centres_hz = centres_hz[self.channel_ids]
        self.block_in = block_out * decimation
        # chan / demod are injectable so the shard / broadcast / gather logic runs without a GPU in the CPU
        # suite; the product always builds the CUDA objects
        self.demod = demod if demod is not None else FMDemod(block_out, len(self.channel_ids), device=device,
                                                             pipeline_depth=pipeline_depth)
        self.chan = chan if chan is not None else Channelizer(fs_in_hz, self.centres_hz, decimation, n_taps, block_out,
                                                              device=device, mode=mode,
                                                              ring_depth=getattr(self.demod, "depth", 0))

    def feed(self, iq_block, producer_stream=None) -> int:
        """One wideband block (device u8 tensor / pointer of 2 * block_in bytes), asynchronous.  producer_stream:
        handle of the CUDA stream whose queued work produces iq_block -- 0 is a valid handle (the legacy default
        stream); None means "no producer to wait for" (the block is already complete)."""
        if producer_stream is not None:
            self.chan.wait_external_stream(int(producer_stream))
        return self.chan.feed(self.demod, iq_block)

    def broadcast_and_feed(self, iq_block, src: int = 0, group=None) -> int:
        """iq_block: torch.uint8 tensor of 2 * block_in bytes, valid on rank `src`, overwritten elsewhere.  The
        channelizer copies the block into its own staging buffer on its own stream; the caller's current stream is
        made to wait for that copy, so the next broadcast into iq_block cannot overtake it."""
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(iq_block, src=src, group=group)
        if not iq_block.is_cuda:
            return self.feed(iq_block, None)
        cur = torch.cuda.current_stream()
        rc = self.feed(iq_block, cur.cuda_stream)
        cur.wait_stream(torch.cuda.ExternalStream(self.chan.stream))
        return rc

    def results(self):
        """[(channel id, PI, PS, RT, n_groups)] of this rank's channels, decoded on the device (K6)."""
        self.demod.rds_fetch()
        out = []
        for i, c in enumerate(self.channel_ids):
            db = self.demod.rds_db(i)
            out.append((c, db["pi"], db["ps"], db["rt"], self.demod.rds_counts(i)[0]))
        return out

    def close(self):
        for o in (self.chan, self.demod):
            if hasattr(o, "close"):
                o.close()
