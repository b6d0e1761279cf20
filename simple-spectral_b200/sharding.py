"""Multi-GPU decomposition of a frame.  Samples are independent and seeded per sample index,
so any partition of (pixels x samples) gives the same per-pixel sums up to f64 summation order.  The path has
exactly one exchange step: the sum of the ranks' f64 XYZA accumulators (one NCCL reduce to rank 0)."""


def sample_shard(rank, world, spp_total):
    """Contiguous sample-index range [begin, end) of `rank` (option B of SURVEY.md §8e: perfectly balanced)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(spp_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def tile_shard(rank, world, height):
    """Row band [y0, y1) of `rank` (option A: matches the reference's Framebuffer::Tile decomposition)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(height, world)
    y0 = rank * base + min(rank, rem)
    return y0, y0 + base + (1 if rank < rem else 0)


def shard_options(opt_factory, rank, world, spp_total, mode="samples", height=None, band_height=8):
    """opt_factory(**fields) -> ssb_options for this rank's part of the job."""
    if mode == "samples":
        b, e = sample_shard(rank, world, spp_total)
        return opt_factory(spp=spp_total, sample_begin=b, sample_end=e)
    if mode == "bands":  # interleaved bands of band_height rows (ssb_options.band_*): what the library and bench.py use
        return opt_factory(spp=spp_total, band_height=band_height, band_count=world, band_index=rank)
    y0, y1 = tile_shard(rank, world, height)
    return opt_factory(spp=spp_total, y0=y0, y1=y1)


def reduce_accumulators(accum, dist=None, dst=0):
    """The single exchange step.  `accum`: torch tensor (f64, width*height*4) holding this rank's raw sums."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
