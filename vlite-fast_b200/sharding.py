"""Antenna -> rank assignment of the multi-GPU path.

The reference runs one writer + process_baseband (+ heimdall) per antenna per GPU
(config/hosts:4-19, scripts/start_single:29-45) and co-adds filterbanks with an
external MPI program (scripts/start_coadd:16-59).  Here every rank (one per
GPU) channelises its own antennas with no data-path collective; the co-add is
one reduce of the f32 pre-digitisation tiles per segment, scaled by
1/sqrt(total antennas) and digitised on the root (vf_coadd_segment)."""


def antennas_of_rank(total_antennas, world_size, rank):
    """antenna a lives on rank a % world_size (16 antennas -> 8/4/2 per GPU at 2/4/8 GPUs)"""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside 0..%d" % (rank, world_size - 1))
    return [a for a in range(total_antennas) if a % world_size == rank]


def coadd_scale(total_antennas):
    import math
    return 1.0 / math.sqrt(float(total_antennas))
