"""Time-block and channel sharding of the frame path over ranks.

Frame f depends only on samples [f*hop, f*hop + N) (src/fourier/audio_transform.rs:37-41), so a
recording splits into contiguous frame blocks with an N-hop sample halo and NO exchange between
ranks.  torch.distributed is used only for the barrier and for reducing the timing (max over ranks).
"""
from dataclasses import dataclass


def num_frames(n_samples, window_n, hop):
    return 0 if n_samples < window_n or hop <= 0 else (n_samples - window_n) // hop + 1


@dataclass(frozen=True)
class TimeBlock:
    rank: int
    frame_lo: int          # first frame of this rank (inclusive)
    frame_hi: int          # last frame (exclusive)
    sample_lo: int         # first sample the rank needs
    sample_hi: int         # one past the last sample it needs (includes the N-hop halo)

    @property
    def frames(self):
        return self.frame_hi - self.frame_lo

    @property
    def samples(self):
        return max(0, self.sample_hi - self.sample_lo)


def time_blocks(n_samples, window_n, hop, world, frame_align=1):
    """Rank g of `world` owns frames [F*g/world, F*(g+1)/world) of the F frames of the recording.

    `frame_align`: block boundaries are rounded down to a multiple of it.  Mono plans transform frames in pairs
    (frame 2c and 2c+1 share one complex FFT), so use frame_align=2 there: a shard that started on an odd frame
    would put every frame in the other slot of its pair, which is the same column only up to the last bit of
    the FFT.  With aligned boundaries shard == unsharded byte for byte."""
    total = num_frames(n_samples, window_n, hop)
    blocks = []
    for g in range(world):
        lo = (total * g // world) // frame_align * frame_align
        hi = total if g == world - 1 else (total * (g + 1) // world) // frame_align * frame_align
        s_lo = lo * hop
        s_hi = (hi - 1) * hop + window_n if hi > lo else s_lo
        blocks.append(TimeBlock(g, lo, hi, s_lo, s_hi))
    return blocks


def channel_blocks(n_channels, world):
    """cfg4: whole channels per rank, no halo.  Returns [(channel_lo, channel_hi)] per rank."""
    return [(n_channels * g // world, n_channels * (g + 1) // world) for g in range(world)]


def max_over_ranks(value, dist=None, device=None):
    """Max of a Python float over all ranks (the device time of a multi-GPU step)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_block_native(n_samples, window_n, hop, channels, rank, world):
    """The same block from the C ABI (sbx_shard_time_block: host arithmetic, no GPU needed); channels == 1 aligns the
    boundaries to even frames like time_blocks(frame_align=2)."""
    import ctypes as C
    from . import binding
    v = [C.c_uint64() for _ in range(4)]
    st = binding.load().sbx_shard_time_block(n_samples, window_n, hop, channels, rank, world, *[C.byref(x) for x in v])
    if st != binding.SBX_OK:
        raise ValueError(f"sbx_shard_time_block: status {st}")
    lo, hi, s_lo, s_hi = (int(x.value) for x in v)
    return TimeBlock(rank, lo, hi, s_lo, s_hi)
