"""vlite-fast_b200: B200-native baseband -> filterbank chain of VLITE-Fast.

The product is the C-ABI library ``libvlitefast.so`` (include/vlitefast.h,
sources in csrc/) and the C host programs in host/.  This Python package is only
the ctypes view of that ABI which the tests and bench.py use; it has no compute
of its own and no fallback: importing ``Pipeline`` without the built library
raises.

The directory name carries a hyphen (it mirrors the reference's name), so it
is loaded through ``__graft_entry__.load_package()`` under the module name
``vlite_fast_b200``.
"""
from .sharding import antennas_of_rank, coadd_scale
from .binding import (VfConfig, Pipeline, bind_thread_to_gpu, VfError, lib, hostlib, GenParams, gen_samples,
                      gen_vdif_second, NFFT, NCHANOUT, NSCRUNCH, NSUB, VD_FRM, VD_DAT,
                      FRAMES_PER_SEC, VfgConfig, GpuGenerator, genlib)

__all__ = ["VfConfig", "Pipeline", "bind_thread_to_gpu", "VfError", "lib", "hostlib", "GenParams", "gen_samples",
           "gen_vdif_second", "NFFT", "NCHANOUT", "NSCRUNCH", "NSUB", "VD_FRM", "VD_DAT",
           "FRAMES_PER_SEC", "VfgConfig", "GpuGenerator", "genlib"]
