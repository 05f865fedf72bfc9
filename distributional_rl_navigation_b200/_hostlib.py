"""ctypes binding of libmnv_host.so (include/mnv_host.h): the native multi-threaded host-side expander of the compact
observation packet (VecMarineNavEnv.step_host).  Like the CUDA library it is built in-tree by
`python -m distributional_rl_navigation_b200.build`; there is no Python fallback."""
import ctypes as C
import os

from ._lib import MarinenavError

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_PKG, "lib", "libmnv_host.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(SO_PATH):
            raise MarinenavError(f"{SO_PATH} is missing: build it with `python -m distributional_rl_navigation_b200.build`")
        L = C.CDLL(SO_PATH)
        L.mnvh_create.restype, L.mnvh_create.argtypes = C.c_void_p, [C.c_int, C.c_int64, C.c_int, C.c_int]
        L.mnvh_destroy.restype, L.mnvh_destroy.argtypes = None, [C.c_void_p]
        L.mnvh_threads.restype, L.mnvh_threads.argtypes = C.c_int, [C.c_void_p]
        L.mnvh_rescan.restype, L.mnvh_rescan.argtypes = None, [C.c_void_p, C.c_void_p]
        L.mnvh_expand.restype, L.mnvh_expand.argtypes = None, [C.c_void_p] * 7
        L.mnvh_expand_early.restype, L.mnvh_expand_early.argtypes = None, [C.c_void_p] * 7
        L.mnvh_rescan_skipped.restype, L.mnvh_rescan_skipped.argtypes = None, [C.c_void_p] * 3
        _lib = L
    return _lib


def default_threads_and_first_cpu():
    """Workers for this process and the first CPU to pin them to: the CPUs this process may run on are split evenly among
    the ranks of the node (LOCAL_RANK / LOCAL_WORLD_SIZE from torchrun), so that the expanders of different ranks never
    share a core; at most 8 workers per rank.  MNV_HOST_THREADS overrides the count; MNV_HOST_PIN=1 pins worker t to the
    t-th CPU of this rank's share (only when the CPU set is a contiguous range)."""
    cpus = sorted(os.sched_getaffinity(0))
    lw, lr = int(os.environ.get("LOCAL_WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    share = max(1, len(cpus) // max(1, lw))
    n = int(os.environ.get("MNV_HOST_THREADS", "0")) or max(1, min(8, share))
    contiguous = cpus == list(range(cpus[0], cpus[0] + len(cpus)))
    pin = os.environ.get("MNV_HOST_PIN", "0") == "1"           # off by default: measured slower on a shared VM (profiles/README.md)
    first = cpus[0] + lr * share if (pin and contiguous and n <= share) else -1
    return n, first


class Expander:
    """One pool per VecMarineNavEnv: expand(obs, head, skip | None, mask, dir, vals) (raw addresses) updates obs in place."""

    def __init__(self, E, obs_dim, n_threads=None, cpu_first=None):
        n, first = default_threads_and_first_cpu()
        n = n if n_threads is None else n_threads
        first = first if cpu_first is None else cpu_first
        self._L = load()
        self._p = self._L.mnvh_create(int(n), int(E), int(obs_dim), int(first))
        if not self._p:
            raise MarinenavError(f"mnvh_create({n}, {E}, {obs_dim}) failed")
        self.E, self.obs_dim, self.n_threads = int(E), int(obs_dim), self._L.mnvh_threads(self._p)

    def rescan(self, obs_ptr):
        self._L.mnvh_rescan(self._p, obs_ptr)

    def expand(self, obs_ptr, head_ptr, skip_ptr, mask_ptr, dir_ptr, vals_ptr):
        self._L.mnvh_expand(self._p, obs_ptr, head_ptr, skip_ptr, mask_ptr, dir_ptr, vals_ptr)

    def expand_early(self, obs_ptr, head_ptr, skip_ptr, mask_ptr, dir_ptr, vals_ptr):
        """expand() while the rows flagged in skip are still on their way: they are left for rescan_skipped()."""
        self._L.mnvh_expand_early(self._p, obs_ptr, head_ptr, skip_ptr, mask_ptr, dir_ptr, vals_ptr)

    def rescan_skipped(self, obs_ptr, skip_ptr):
        self._L.mnvh_rescan_skipped(self._p, obs_ptr, skip_ptr)

    def close(self):
        if self._p:
            self._L.mnvh_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
