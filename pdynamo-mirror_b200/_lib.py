"""ctypes binding of libnbabfs_b200.so (include/nbabfs_b200.h).  There is no CPU fallback: if the CUDA
library is missing or no device is visible, loading / SetUp fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnbabfs_b200.so")
STATUS_CONTINUE = 16
_LIB = None

dp, ip, vp, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_long)

SIGNATURES = {
    "nbb200_device_count": (C.c_int, []),
    "nbb200_last_error": (C.c_char_p, []),
    "nbb200_version": (C.c_char_p, []),
    "NBModelABFSState_B200_SetUp": (vp, [C.c_int, C.c_int, dp, ip, C.c_int, ip, dp, dp, C.c_int, ip, dp, dp,
                                        C.c_int, ip, C.c_int, ip, C.c_int, dp, dp, ip]),
    "NBModelABFSState_B200_Deallocate": (None, [C.POINTER(vp)]),
    "NBModelABFSState_B200_SetFixedAtoms": (None, [vp, C.c_int, ip, ip]),
    "NBModelABFSState_B200_SetQCAtoms": (None, [vp, C.c_int, ip, ip]),
    "NBModelABFS_B200_QCMMEnergyLJ": (None, [vp, dp, dp, dp, ip]),
    "NBModelABFS_B200_QCMMPotentials": (None, [vp, dp, dp, ip]),
    "NBModelABFS_B200_QCMMGradients": (None, [vp, dp, dp, dp, ip]),
    "NBModelABFSState_B200_SetUpCentering": (None, [vp, C.c_int, ip]),
    "NBModelABFS_B200_SetOptions": (None, [vp] + [C.c_double] * 6 + [C.c_int, C.c_int]),
    "NBModelABFS_B200_Update": (C.c_int, [vp, dp, dp, C.c_int, ip]),
    "NBModelABFS_B200_UpdateDevice": (C.c_int, [vp, vp, dp, C.c_int, ip]),
    "NBModelABFS_B200_MMMMEnergy": (None, [vp, dp, dp, dp, ip]),
    "NBModelABFS_B200_MMMMEnergyDevice": (None, [vp, dp, vp, dp, ip]),
    "NBModelABFS_B200_MMMMEnergyDeviceDeferred": (None, [vp, dp, vp, dp, ip]),
    "nbb200_flush": (None, [vp, ip]),
    "nbb200_copy_to_host_async": (None, [vp, vp, vp, C.c_size_t]),
    "NBModelABFSState_B200_NumberOfPairs": (C.c_long, [vp, C.c_int]),
    "NBModelABFSState_B200_NumberOfImages": (C.c_int, [vp]),
    "NBModelABFSState_B200_NumberOfImagePairs": (C.c_long, [vp]),
    "NBModelABFSState_B200_NumberOf14Pairs": (C.c_long, [vp]),
    "NBModelABFSState_B200_GetImageInfo": (None, [vp, C.c_int, ip, dp]),
    "NBModelABFSState_B200_GetPairs": (C.c_long, [vp, C.c_int, ip, ip]),
    "PairListGenerator_B200_SelfPairListFromCoordinates3": (C.c_long, [C.c_int, C.c_int, dp, C.c_double, C.c_int, ip, C.POINTER(ip), ip]),
    "PairListGenerator_B200_CrossPairListFromDoubleCoordinates3": (C.c_long, [C.c_int, C.c_int, dp, C.c_int, dp, C.c_double, C.POINTER(ip), ip]),
    "nbb200_free": (None, [vp]),
    "nbb200_host_alloc": (vp, [C.c_size_t]),
    "nbb200_host_free": (None, [vp]),
    "PairwiseInteractionABFS_B200_MakeFactors": (None, [C.c_double] * 3 + [dp]),
    "PairwiseInteractionABFS_B200_SetInteractionForm": (None, [vp, C.c_int, C.c_int, ip]),
    "PairwiseInteractionABFS_B200_MakeSpline": (C.c_int, [C.c_int] + [C.c_double] * 3 + [C.c_int, dp, dp, dp]),
    "nbb200_set_stream": (None, [vp, vp]),
    "nbb200_measure_fp32_peak": (C.c_double, [C.c_int, ip]),
    "nbb200_enable_timing": (None, [vp, C.c_int]),
    "nbb200_get_timings": (None, [vp, dp]),
    "nbb200_get_counters": (None, [vp, lp]),
    "nbb200_set_partition": (None, [vp, C.c_int, C.c_int]),
    "nbb200_set_restricted_sort": (None, [vp, C.c_int]),
    "nbb200_own_slab_to_host": (None, [vp, vp, vp]),
    "nbb200_host_gather_rows": (None, [vp, vp, C.c_long, vp]),
    "nbb200_host_copy": (None, [vp, vp, C.c_long]),
    "nbb200_host_add": (None, [vp, vp, C.c_long]),
    "nbb200_peer_export_chunks": (C.c_int, [vp, C.c_char_p]),
    "nbb200_peer_import_chunks": (C.c_int, [vp, C.c_int, C.c_char_p]),
    "nbb200_peer_attach_local_chunks": (C.c_int, [vp, C.c_int, vp]),
    "nbb200_chunk_upload": (None, [vp, vp, C.c_long, C.c_long]),
    "nbb200_chunk_signal": (None, [vp, C.c_long, C.c_int]),
    "nbb200_chunk_wait": (None, [vp, C.c_long, C.c_int]),
    "nbb200_chunk_gather_owned": (None, [vp, vp]),
    "nbb200_chunk_scatter_gradients": (None, [vp]),
    "nbb200_chunk_download_add": (C.c_int, [vp, vp, C.c_long, C.c_long]),
    "nbb200_chunk_download": (C.c_int, [vp, vp, C.c_long, C.c_long, C.c_int]),
    "nbb200_host_scatter_add_rows": (None, [vp, vp, C.c_long, vp]),
    "nbb200_set_gradient_overwrite": (None, [vp, C.c_int]),
    "nbb200_set_optimistic_updates": (None, [vp, C.c_int]),
    "nbb200_set_list_reuse_hint": (None, [vp, C.c_int]),
    "NBModelABFSState_B200_GetStatistics": (None, [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]),
    "nbb200_vv_first_half": (None, [vp, vp, vp, vp, C.c_double]),
    "nbb200_vv_second_half": (None, [vp, vp, vp, vp, vp, C.c_double, vp]),
    "nbb200_md_run": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, dp, C.c_double, dp, C.c_double, C.c_ulonglong, C.c_ulonglong, vp, dp, dp, dp, dp, ip]),
    "nbb200_langevin_first_half": (None, [vp, vp, vp, vp, vp, dp, C.c_ulonglong, C.c_ulonglong]),
    "nbb200_set_langevin_constraints": (None, [vp, C.c_int, C.c_double]),
    "MMTerms_B200_Allocate": (vp, [C.c_int, C.c_int, ip]),
    "MMTerms_B200_Deallocate": (None, [C.POINTER(vp)]),
    "MMTerms_B200_SetStream": (None, [vp, vp]),
    "HarmonicBondContainer_B200_Define": (None, [vp, C.c_int, C.c_int, ip, ip, C.c_char_p, C.c_int, dp, dp, ip]),
    "HarmonicAngleContainer_B200_Define": (None, [vp, C.c_int, ip, ip, C.c_char_p, C.c_int, dp, dp, ip]),
    "FourierDihedralContainer_B200_Define": (None, [vp, C.c_int, ip, ip, C.c_char_p, C.c_int, dp, ip, dp, ip]),
    "HarmonicImproperContainer_B200_Define": (None, [vp, C.c_int, ip, ip, C.c_char_p, C.c_int, dp, dp, ip]),
    "MMTerms_B200_Energy": (None, [vp, dp, dp, dp, ip]),
    "MMTerms_B200_EnergyDevice": (None, [vp, vp, dp, vp, ip]),
    "MMTerms_B200_EnergyDeviceEnqueue": (None, [vp, vp, vp, ip]),
    "MMTerms_B200_EnergyDeviceCollect": (None, [vp, dp, ip]),
    "MMTerms_B200_LastEnergies": (None, [vp, dp]),
    "MMTerms_B200_NumberOfTerms": (C.c_long, [vp, C.c_int]),
    "nbb200_get_slab": (None, [vp, lp]),
    "nbb200_touched_ranges": (C.c_int, [vp, lp]),
    "nbb200_touched_ranges_device": (C.c_int, [vp, vp]),
    "nbb200_set_sorted_gradient_buffer": (None, [vp, vp]),
    "nbb200_max_displacement": (C.c_double, [vp, vp, ip]),
    "NBModelABFS_B200_UpdateDeviceDecided": (C.c_int, [vp, vp, dp, C.c_int, ip]),
    "NBModelABFS_B200_MMMMEnergySorted": (None, [vp, dp, dp, ip]),
    "NBModelABFS_B200_MMMMEnergySortedEnqueue": (None, [vp, ip]),
    "NBModelABFS_B200_MMMMEnergySortedFinish": (None, [vp, dp, dp, ip]),
    "nbb200_gather_sorted": (None, [vp, vp, C.c_long, C.c_long, vp]),
    "nbb200_scatter_sorted": (None, [vp, vp, C.c_long, C.c_long, vp]),
    "nbb200_unsort_add": (None, [vp, C.c_long, C.c_long, vp]),
    "nbb200_peer_export": (C.c_int, [vp, C.c_char_p]),
    "nbb200_peer_import": (C.c_int, [vp, C.c_int, C.c_char_p]),
    "nbb200_peer_attach_local": (C.c_int, [vp, C.c_int, vp]),
    "nbb200_peer_begin": (None, [vp, vp, C.c_long, C.c_long]),
    "nbb200_peer_pull_positions": (None, [vp, vp, lp, C.c_int, vp]),
    "nbb200_peer_push_gradients": (None, [vp, vp]),
    "nbb200_peer_signal_begin": (None, [vp, C.c_long, vp, C.c_int]),
    "nbb200_peer_wait_begin": (C.c_double, [vp, C.c_long, C.c_int, ip]),
    "nbb200_peer_signal_end": (None, [vp, C.c_long, dp]),
    "nbb200_peer_signal_end_device": (None, [vp, C.c_long, ip]),
    "nbb200_peer_wait_end": (None, [vp, C.c_long]),
    "nbb200_peer_read_sums": (None, [vp, dp, ip]),
}


class CLibraryError(RuntimeError):
    """Raised where the reference raises pCore.CLibraryError (pCore-1.9.0/pCore/CoreObjects.py:12)."""


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise CLibraryError("libnbabfs_b200.so is not built (run python pdynamo-mirror_b200/build.py); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


def last_error():
    return lib().nbb200_last_error().decode()


def d_(a):
    return None if a is None else a.ctypes.data_as(dp)


def i_(a):
    return None if a is None else a.ctypes.data_as(ip)


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            lib().nbb200_host_free(self.ptr)
        except Exception:
            pass


def pinned_array(shape, dtype="float64"):
    """numpy array in page-locked host memory (nbb200_host_alloc): transferred by DMA without a staging copy.
    Falls back to an ordinary array when no CUDA device is available (there is nothing to transfer to then)."""
    import numpy as np
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib().nbb200_host_alloc(n) if lib().nbb200_device_count() > 0 else None
    if not ptr:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_char * n).from_address(ptr)
    buf._owner = _PinnedOwner(ptr)                      # the ctypes buffer is the numpy base: freed with the last view
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    arr[...] = 0
    return arr
