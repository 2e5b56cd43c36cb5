"""ctypes mirror of include/PFAC.h + include/PFAC_ext.h.

Method names follow the C entry points (reference PFAC/include/PFAC.h:87-215):
PFAC.readPatternFromFile, matchFromHost, matchFromDevice, matchFromDeviceReduce, ...
Device buffers are anything with .data_ptr() (torch CUDA tensors) or raw integer addresses.
Every call raises PFACError(status) on a non-success PFAC_status_t, carrying the library's
own PFAC_getErrorString text.  There is no fallback: if libpfac.so is missing or no B200 is
present, construction fails.
"""
import ctypes
import enum
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_PKG, "lib", "libpfac.so")
_lib = None


class Status(enum.IntEnum):  # include/PFAC.h (reference PFAC.h:57-70)
    SUCCESS = 0
    BASE = 10000
    ALLOC_FAILED = 10001
    CUDA_ALLOC_FAILED = 10002
    INVALID_HANDLE = 10003
    INVALID_PARAMETER = 10004
    PATTERNS_NOT_READY = 10005
    FILE_OPEN_ERROR = 10006
    LIB_NOT_EXIST = 10007
    ARCH_MISMATCH = 10008
    MUTEX_ERROR = 10009
    INTERNAL_ERROR = 10010


class Platform(enum.IntEnum):
    GPU = 0
    CPU = 1
    CPU_OMP = 2


class TextureMode(enum.IntEnum):
    AUTOMATIC = 0
    TEXTURE_ON = 1
    TEXTURE_OFF = 2


class PerfMode(enum.IntEnum):
    TIME_DRIVEN = 0
    SPACE_DRIVEN = 1


class TableInfo(ctypes.Structure):
    _fields_ = [("num_patterns", ctypes.c_int), ("num_states", ctypes.c_int),
                ("initial_state", ctypes.c_int), ("max_pattern_len", ctypes.c_int),
                ("num_leaves", ctypes.c_int), ("num_edges", ctypes.c_int),
                ("hash_edges", ctypes.c_int), ("num_chains", ctypes.c_int),
                ("tail_bytes", ctypes.c_int), ("chains_hot", ctypes.c_int),
                ("next2_hot", ctypes.c_int), ("code_bits", ctypes.c_int),
                ("gram_len", ctypes.c_int), ("has_best2", ctypes.c_int), ("has_chk2", ctypes.c_int),
                ("max_depth", ctypes.c_int), ("hot_depth", ctypes.c_int),
                ("hot_buckets", ctypes.c_uint), ("cold_buckets", ctypes.c_uint),
                ("hash_mul", ctypes.c_uint), ("hot_max_probe", ctypes.c_int),
                ("cold_max_probe", ctypes.c_int), ("pre2_bits_set", ctypes.c_int),
                ("root_fanout", ctypes.c_int), ("hashed_filter", ctypes.c_int),
                ("hfilt_bits_set", ctypes.c_int), ("code_shift", ctypes.c_int),
                ("device_bytes", ctypes.c_size_t), ("hfilt_words", ctypes.c_uint)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class PFACError(RuntimeError):
    def __init__(self, status, where=""):
        self.status = int(status)
        try:
            text = load_library().PFAC_getErrorString(self.status).decode()
        except Exception:  # pragma: no cover
            text = "status %d" % self.status
        super().__init__("%s%s" % (where + ": " if where else "", text))


def library_path():
    return _LIB_PATH


def load_library():
    """dlopen pfac_b200/lib/libpfac.so (build it with `python -m pfac_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise FileNotFoundError(
            _LIB_PATH + " is missing: run `python -m pfac_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    L = ctypes.CDLL(_LIB_PATH)
    vp, cp, sz, ip = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)
    ull = ctypes.c_ulonglong
    sig = {
        "PFAC_create": [ctypes.POINTER(vp)],
        "PFAC_destroy": [vp],
        "PFAC_setPlatform": [vp, ctypes.c_int],
        "PFAC_setTextureMode": [vp, ctypes.c_int],
        "PFAC_setPerfMode": [vp, ctypes.c_int],
        "PFAC_dumpTransitionTable": [vp, vp],
        "PFAC_dumpTransitionTableToFile": [vp, cp],
        "PFAC_readPatternFromFile": [vp, cp],
        "PFAC_readPatternFromMemory": [vp, cp, sz],
        "PFAC_readPatternFromArrays": [vp, vp, vp, sz],
        "PFAC_tableCompileArrays": [vp, vp, sz, sz, ctypes.POINTER(vp)],
        "PFAC_matchFromDevice": [vp, vp, sz, vp],
        "PFAC_matchFromHost": [vp, vp, sz, vp],
        "PFAC_matchFromDeviceReduce": [vp, vp, sz, vp, vp, ip],
        "PFAC_matchFromHostReduce": [vp, vp, sz, vp, vp, ip],
        "PFAC_reduceOnDevice": [vp, vp, sz, vp, vp, ip],
        "PFAC_reduceInplaceOnDevice": [vp, vp, sz, vp, vp, ip],
        "PFAC_setStream": [vp, vp],
        "PFAC_matchShardFromDevice": [vp, vp, sz, sz, vp],
        "PFAC_matchFromDeviceReduce64": [vp, vp, sz, vp, vp, ctypes.POINTER(ull)],
        "PFAC_matchShardFromDeviceReduce64": [vp, vp, sz, sz, ctypes.c_longlong, vp, vp,
                                              ctypes.POINTER(ull)],
        "PFAC_tableCompile": [cp, sz, sz, ctypes.POINTER(vp)],
        "PFAC_tableCompileFile": [cp, sz, ctypes.POINTER(vp)],
        "PFAC_tableDestroy": [vp],
        "PFAC_tableDump": [vp, vp],
        "PFAC_tableDumpToFile": [vp, cp],
        "PFAC_tableGetInfo": [vp, ctypes.POINTER(TableInfo)],
        "PFAC_tableGetLayout": [vp] + [ctypes.POINTER(vp)] * 8,
        "PFAC_tableGetLayout2": [vp] + [ctypes.POINTER(vp)] * 3,
        "PFAC_tableGetFilter": [vp, ctypes.POINTER(vp)],
        "PFAC_tableSave": [vp, cp],
        "PFAC_tableLoad": [cp, ctypes.POINTER(vp)],
        "PFAC_saveCompiledPatterns": [vp, cp],
        "PFAC_loadCompiledPatterns": [vp, cp],
        "PFAC_getTableInfo": [vp, ctypes.POINTER(TableInfo)],
        "PFAC_getTableInfoReduce": [vp, ctypes.POINTER(TableInfo)],
        "PFAC_memoryUsage": [vp],
        "PFAC_hostCopy": [vp, vp, sz],
        "PFAC_hostZero": [vp, sz],
        "PFAC_lastHostTransfer": [vp, ctypes.POINTER(sz), ctypes.POINTER(sz)],
        "PFAC_releaseHostBuffers": [vp],
        "PFAC_commCreate": [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, sz, vp],
        "PFAC_commConnect": [vp, vp],
        "PFAC_commCreateLocal": [ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int), ctypes.c_int, sz],
        "PFAC_commDestroy": [vp],
        "PFAC_commGlobalList": [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(sz)],
        "PFAC_commReadGlobalList": [vp, sz, sz, vp, vp],
        "PFAC_matchShardFromDeviceReduce64Global": [vp, vp, vp, sz, sz, ctypes.c_longlong, vp, vp, sz, vp, vp],
        "PFAC_matchShardFromDeviceReduce64Cap": [vp, vp, sz, sz, ctypes.c_longlong, vp, vp, sz, ctypes.POINTER(ull)],
        "PFAC_commGatherRuns": [vp, vp, ctypes.c_int, vp, vp, vp, ctypes.c_int],
        "PFAC_mgpuCreate": [ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int), ctypes.c_int],
        "PFAC_mgpuDestroy": [vp],
        "PFAC_mgpuReadPatternFromFile": [vp, cp],
        "PFAC_mgpuMatchFromHost": [vp, vp, sz, vp],
        "PFAC_mgpuMatchFromHostReduce64": [vp, vp, sz, vp, vp, ctypes.POINTER(ull)],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = ctypes.c_int
    L.PFAC_getErrorString.argtypes = [ctypes.c_int]
    L.PFAC_getErrorString.restype = ctypes.c_char_p
    L.PFAC_versionString.restype = ctypes.c_char_p
    L.PFAC_kernelLaunchCount.restype = ull
    _lib = L
    return L


def host_copy(dst, src):
    """PFAC_hostCopy: the staging pool's multi-threaded memcpy (host only).  dst, src: ndarrays."""
    if dst.nbytes != src.nbytes:
        raise ValueError("size mismatch")
    _check(load_library().PFAC_hostCopy(dst.ctypes.data, src.ctypes.data, src.nbytes), "PFAC_hostCopy")
    return dst


def kernel_launch_count():
    return int(load_library().PFAC_kernelLaunchCount())


def _ptr(x):
    """Address of a device tensor / numpy array / raw int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    raise TypeError("expected a tensor, ndarray or address, got %r" % type(x))


def _check(status, where):
    if status != 0:
        raise PFACError(status, where)


class PFAC:
    """One PFAC_handle_t.  Binds to the current CUDA device (reference PFAC.cpp:103-132)."""

    def __init__(self):
        self._L = load_library()
        self._h = ctypes.c_void_p()
        _check(self._L.PFAC_create(ctypes.byref(self._h)), "PFAC_create")

    # -- lifecycle -----------------------------------------------------------------------
    def destroy(self):
        if self._h:
            st = self._L.PFAC_destroy(self._h)
            self._h = ctypes.c_void_p()
            _check(st, "PFAC_destroy")

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    @property
    def handle(self):
        return self._h

    # -- configuration ---------------------------------------------------------------------
    def setPlatform(self, platform):
        _check(self._L.PFAC_setPlatform(self._h, int(platform)), "PFAC_setPlatform")

    def setTextureMode(self, mode):
        _check(self._L.PFAC_setTextureMode(self._h, int(mode)), "PFAC_setTextureMode")

    def setPerfMode(self, mode):
        _check(self._L.PFAC_setPerfMode(self._h, int(mode)), "PFAC_setPerfMode")

    def setStream(self, stream):
        """stream: a torch.cuda.Stream, a raw cudaStream_t address, or None (default stream)."""
        addr = getattr(stream, "cuda_stream", stream) or None
        _check(self._L.PFAC_setStream(self._h, addr), "PFAC_setStream")

    def readPatternFromFile(self, filename):
        _check(self._L.PFAC_readPatternFromFile(self._h, os.fsencode(filename)),
               "PFAC_readPatternFromFile")

    def readPatternFromMemory(self, image):
        image = bytes(image)
        _check(self._L.PFAC_readPatternFromMemory(self._h, image, len(image)),
               "PFAC_readPatternFromMemory")

    def readPatternFromArrays(self, patterns):
        """patterns: list of bytes; any byte value allowed, ID = index + 1."""
        ptrs, lens, keep = _pattern_arrays(patterns)
        _check(self._L.PFAC_readPatternFromArrays(self._h, ptrs, lens, len(patterns)),
               "PFAC_readPatternFromArrays")
        del keep

    def saveCompiledPatterns(self, filename):
        _check(self._L.PFAC_saveCompiledPatterns(self._h, os.fsencode(filename)), "PFAC_saveCompiledPatterns")

    def loadCompiledPatterns(self, filename):
        _check(self._L.PFAC_loadCompiledPatterns(self._h, os.fsencode(filename)), "PFAC_loadCompiledPatterns")

    def dumpTransitionTable(self, filename):
        _check(self._L.PFAC_dumpTransitionTableToFile(self._h, os.fsencode(filename)),
               "PFAC_dumpTransitionTable")

    def tableInfo(self, reduce=False):
        """Compiled-table facts of the dense kernel's layout (reduce=True: the reduce kernel's, which has its own
        shared-memory budget and first stage)."""
        info = TableInfo()
        if reduce:
            _check(self._L.PFAC_getTableInfoReduce(self._h, ctypes.byref(info)), "PFAC_getTableInfoReduce")
        else:
            _check(self._L.PFAC_getTableInfo(self._h, ctypes.byref(info)), "PFAC_getTableInfo")
        return info.as_dict()

    # -- matching: device buffers ----------------------------------------------------------------
    def matchFromDevice(self, d_input, size, d_result):
        """d_result[i] (int32) for i<size.  Asynchronous on the handle's stream."""
        _check(self._L.PFAC_matchFromDevice(self._h, _ptr(d_input), size, _ptr(d_result)),
               "PFAC_matchFromDevice")

    def matchShardFromDevice(self, d_input, n_owned, n_total, d_result):
        _check(self._L.PFAC_matchShardFromDevice(self._h, _ptr(d_input), n_owned, n_total,
                                                 _ptr(d_result)), "PFAC_matchShardFromDevice")

    def matchFromDeviceReduce(self, d_input, size, d_result, d_pos, alias=None):
        """Returns M; d_result[0..M) ids and d_pos[0..M) int32 positions, ascending position."""
        n = ctypes.c_int(0)
        fn = {None: self._L.PFAC_matchFromDeviceReduce, "reduceOnDevice": self._L.PFAC_reduceOnDevice,
              "reduceInplaceOnDevice": self._L.PFAC_reduceInplaceOnDevice}[alias]
        _check(fn(self._h, _ptr(d_input), size, _ptr(d_result), _ptr(d_pos), ctypes.byref(n)),
               "PFAC_matchFromDeviceReduce")
        return n.value

    def matchFromDeviceReduce64(self, d_input, size, d_result, d_pos64):
        n = ctypes.c_ulonglong(0)
        _check(self._L.PFAC_matchFromDeviceReduce64(self._h, _ptr(d_input), size, _ptr(d_result),
                                                    _ptr(d_pos64), ctypes.byref(n)),
               "PFAC_matchFromDeviceReduce64")
        return n.value

    def matchShardFromDeviceReduce64(self, d_input, n_owned, n_total, pos_base, d_result, d_pos64):
        n = ctypes.c_ulonglong(0)
        _check(self._L.PFAC_matchShardFromDeviceReduce64(self._h, _ptr(d_input), n_owned, n_total,
                                                         pos_base, _ptr(d_result), _ptr(d_pos64),
                                                         ctypes.byref(n)),
               "PFAC_matchShardFromDeviceReduce64")
        return n.value

    def matchShardFromDeviceReduce64Cap(self, d_input, n_owned, n_total, pos_base, d_result, d_pos64):
        """Shard reduce with the buffers' capacity stated (entries of d_result): nothing is stored past it."""
        n = ctypes.c_ulonglong(0)
        cap = min(int(d_result.numel()), int(d_pos64.numel()))
        _check(self._L.PFAC_matchShardFromDeviceReduce64Cap(self._h, _ptr(d_input), n_owned, n_total, pos_base,
                                                            _ptr(d_result), _ptr(d_pos64), cap, ctypes.byref(n)),
               "PFAC_matchShardFromDeviceReduce64Cap")
        return n.value

    def matchShardFromDeviceReduce64Global(self, comm, d_input, n_owned, n_total, pos_base, d_result, d_pos64,
                                           d_scan=None, sync=True):
        """Shard reduce + the cross-GPU exclusive scan of the counts in one kernel (collective over
        `comm`).  sync=True returns (offset, total, count); sync=False leaves them in d_scan (device).
        The capacity passed to the library is the size of d_result / d_pos64."""
        h = (ctypes.c_ulonglong * 3)()
        cap = min(int(d_result.numel()), int(d_pos64.numel())) if hasattr(d_result, "numel") else (1 << 62)
        _check(self._L.PFAC_matchShardFromDeviceReduce64Global(
            self._h, comm._c, _ptr(d_input), n_owned, n_total, pos_base, _ptr(d_result), _ptr(d_pos64), cap,
            _ptr(d_scan), ctypes.cast(h, ctypes.c_void_p) if sync else None),
            "PFAC_matchShardFromDeviceReduce64Global")
        return (int(h[0]), int(h[1]), int(h[2])) if sync else None

    def gatherRuns(self, comm, dst_rank, d_result, d_pos64, d_scan=None, sync=True):
        """Collective: every rank's run goes to rank dst_rank's global list at its scanned offset."""
        _check(self._L.PFAC_commGatherRuns(self._h, comm._c, dst_rank, _ptr(d_result), _ptr(d_pos64), _ptr(d_scan),
                                           1 if sync else 0), "PFAC_commGatherRuns")

    # -- matching: host buffers --------------------------------------------------------------------
    def matchFromHost(self, h_input, h_result=None, size=None):
        """h_input: bytes / uint8 ndarray / pinned CPU tensor.  Returns the int32 result array."""
        src = _host_u8(h_input)
        n = _host_len(src) if size is None else size
        if h_result is None:
            h_result = np.zeros(n, dtype=np.int32)
        _check(self._L.PFAC_matchFromHost(self._h, _ptr(src), n, _ptr(h_result)), "PFAC_matchFromHost")
        return h_result

    def releaseHostBuffers(self):
        _check(self._L.PFAC_releaseHostBuffers(self._h), "PFAC_releaseHostBuffers")

    def lastHostTransfer(self):
        """(h2d_bytes, d2h_bytes) the last matchFromHost* call of this handle moved over PCIe."""
        a, b = ctypes.c_size_t(0), ctypes.c_size_t(0)
        _check(self._L.PFAC_lastHostTransfer(self._h, ctypes.byref(a), ctypes.byref(b)), "PFAC_lastHostTransfer")
        return a.value, b.value

    def matchFromHostReduce(self, h_input, h_result=None, h_pos=None, size=None):
        """Returns (ids[:M], pos[:M]) as int32 arrays (views of the supplied buffers)."""
        src = _host_u8(h_input)
        n = _host_len(src) if size is None else size
        if h_result is None:
            h_result = np.zeros(n, dtype=np.int32)
        if h_pos is None:
            h_pos = np.zeros(n, dtype=np.int32)
        m = ctypes.c_int(0)
        _check(self._L.PFAC_matchFromHostReduce(self._h, _ptr(src), n, _ptr(h_result), _ptr(h_pos),
                                                ctypes.byref(m)), "PFAC_matchFromHostReduce")
        return h_result[:m.value], h_pos[:m.value]


class PFACComm:
    """PFAC_comm: the per-rank block (mailbox + global-list region) behind the in-kernel count scan.

    One process per GPU:  c = PFACComm(rank, world, list_capacity); handles = all-gather of c.handle
    (64 bytes per rank, any transport); c.connect(handles).  `from_torch` does the exchange with
    torch.distributed.  One process, several GPUs: PFACComm.local(devices, list_capacity)."""

    HANDLE_BYTES = 64

    def __init__(self, rank, world, list_capacity=0, _c=None):
        self._L = load_library()
        self.rank, self.world = rank, world
        self.handle = (ctypes.c_ubyte * self.HANDLE_BYTES)()
        if _c is not None:
            self._c = _c
            return
        c = ctypes.c_void_p()
        _check(self._L.PFAC_commCreate(ctypes.byref(c), rank, world, list_capacity,
                                       ctypes.cast(self.handle, ctypes.c_void_p)), "PFAC_commCreate")
        self._c = c
        if world == 1:
            self.connect(bytes(self.handle))

    def connect(self, all_handles):
        buf = bytes(all_handles)
        if len(buf) != self.world * self.HANDLE_BYTES:
            raise ValueError("expected %d handle bytes" % (self.world * self.HANDLE_BYTES))
        _check(self._L.PFAC_commConnect(self._c, buf), "PFAC_commConnect")

    @classmethod
    def from_torch(cls, list_capacity=0, device=None, group=None):
        """Create + exchange the IPC handles with torch.distributed (NCCL all-gather of 64 bytes per rank)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        c = cls(rank, world, list_capacity)
        if world > 1:
            mine = torch.frombuffer(bytearray(bytes(c.handle)), dtype=torch.uint8).to(device)
            allh = torch.empty(world * cls.HANDLE_BYTES, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            c.connect(allh.cpu().numpy().tobytes())
        return c

    @classmethod
    def local(cls, devices, list_capacity=0):
        L = load_library()
        n = len(devices)
        arr = (ctypes.c_void_p * n)()
        devs = (ctypes.c_int * n)(*devices)
        _check(L.PFAC_commCreateLocal(arr, devs, n, list_capacity), "PFAC_commCreateLocal")
        return [cls(i, n, list_capacity, _c=ctypes.c_void_p(arr[i])) for i in range(n)]

    def global_list(self):
        """(ids address, positions address, capacity) of this rank's list region (device memory)."""
        a, b, cap = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._L.PFAC_commGlobalList(self._c, ctypes.byref(a), ctypes.byref(b), ctypes.byref(cap)),
               "PFAC_commGlobalList")
        return a.value, b.value, cap.value

    def read_global_list(self, n, first=0):
        """Host copy of entries [first, first+n) of this rank's list region: (ids int32, positions int64)."""
        ids = np.empty(n, dtype=np.int32)
        pos = np.empty(n, dtype=np.int64)
        _check(self._L.PFAC_commReadGlobalList(self._c, first, n, ids.ctypes.data, pos.ctypes.data),
               "PFAC_commReadGlobalList")
        return ids, pos

    def destroy(self):
        if getattr(self, "_c", None):
            self._L.PFAC_commDestroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class PFACMultiGPU:
    """PFAC_mgpu_t: one process, one handle + host thread per listed device (include/PFAC_ext.h)."""

    def __init__(self, devices):
        self._L = load_library()
        self._h = ctypes.c_void_p()
        arr = (ctypes.c_int * len(devices))(*devices)
        _check(self._L.PFAC_mgpuCreate(ctypes.byref(self._h), arr, len(devices)), "PFAC_mgpuCreate")

    def destroy(self):
        if self._h:
            st = self._L.PFAC_mgpuDestroy(self._h)
            self._h = ctypes.c_void_p()
            _check(st, "PFAC_mgpuDestroy")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def readPatternFromFile(self, filename):
        _check(self._L.PFAC_mgpuReadPatternFromFile(self._h, os.fsencode(filename)),
               "PFAC_mgpuReadPatternFromFile")

    def matchFromHost(self, h_input):
        src = _host_u8(h_input)
        n = _host_len(src)
        out = np.zeros(n, dtype=np.int32)
        _check(self._L.PFAC_mgpuMatchFromHost(self._h, _ptr(src), n, _ptr(out)), "PFAC_mgpuMatchFromHost")
        return out

    def matchFromHostReduce64(self, h_input):
        src = _host_u8(h_input)
        n = _host_len(src)
        ids = np.zeros(max(n, 1), dtype=np.int32)
        pos = np.zeros(max(n, 1), dtype=np.int64)
        m = ctypes.c_ulonglong(0)
        _check(self._L.PFAC_mgpuMatchFromHostReduce64(self._h, _ptr(src), n, _ptr(ids), _ptr(pos),
                                                      ctypes.byref(m)), "PFAC_mgpuMatchFromHostReduce64")
        return ids[:m.value], pos[:m.value]


def _pattern_arrays(patterns):
    keep = [ctypes.create_string_buffer(bytes(p), len(p)) for p in patterns]
    ptrs = (ctypes.c_void_p * len(keep))(*[ctypes.addressof(b) for b in keep])
    lens = (ctypes.c_size_t * len(keep))(*[len(p) for p in patterns])
    return ptrs, lens, keep


def _host_u8(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x.view(np.uint8) if x.dtype != np.uint8 else x)
    return x  # CPU tensor


def _host_len(x):
    if isinstance(x, np.ndarray):
        return x.size
    return x.numel()


class TableCompiler:
    """Host-only table compiler (no GPU needed): PFAC_tableCompile* in include/PFAC_ext.h."""

    def __init__(self, pattern_file=None, image=None, hot_budget_bytes=24 * 1024, patterns=None,
                 compiled_file=None):
        self._L = load_library()
        self._t = ctypes.c_void_p()
        if compiled_file is not None:
            st = self._L.PFAC_tableLoad(os.fsencode(compiled_file), ctypes.byref(self._t))
        elif patterns is not None:
            ptrs, lens, keep = _pattern_arrays(patterns)
            st = self._L.PFAC_tableCompileArrays(ptrs, lens, len(patterns), hot_budget_bytes,
                                                 ctypes.byref(self._t))
            del keep
        elif pattern_file is not None:
            st = self._L.PFAC_tableCompileFile(os.fsencode(pattern_file), hot_budget_bytes,
                                               ctypes.byref(self._t))
        else:
            image = bytes(image)
            st = self._L.PFAC_tableCompile(image, len(image), hot_budget_bytes, ctypes.byref(self._t))
        _check(st, "PFAC_tableCompile")

    def __del__(self):
        if getattr(self, "_t", None):
            self._L.PFAC_tableDestroy(self._t)
            self._t = None

    def info(self):
        info = TableInfo()
        _check(self._L.PFAC_tableGetInfo(self._t, ctypes.byref(info)), "PFAC_tableGetInfo")
        return info.as_dict()

    def dump(self, filename):
        _check(self._L.PFAC_tableDumpToFile(self._t, os.fsencode(filename)), "PFAC_tableDumpToFile")

    def save(self, filename):
        """Write the compiled-table file (PFAC_tableSave); TableCompiler(compiled_file=...) reads it back."""
        _check(self._L.PFAC_tableSave(self._t, os.fsencode(filename)), "PFAC_tableSave")

    def layout(self):
        """Copies of the device layout arrays as a dict: root[256] i32, pre2[2048] u32 (bit-reversed words), rank2[2048] u16, next2 u32, hot/cold
        [n,4] u32 buckets, chains [n,4] u32 records, tails u8, plus hot_depth and mul."""
        ptrs = [ctypes.c_void_p() for _ in range(8)]
        _check(self._L.PFAC_tableGetLayout(self._t, *[ctypes.byref(p) for p in ptrs]),
               "PFAC_tableGetLayout")
        p2 = [ctypes.c_void_p() for _ in range(3)]
        _check(self._L.PFAC_tableGetLayout2(self._t, *[ctypes.byref(p) for p in p2]), "PFAC_tableGetLayout2")
        pf = ctypes.c_void_p()
        _check(self._L.PFAC_tableGetFilter(self._t, ctypes.byref(pf)), "PFAC_tableGetFilter")
        info = self.info()

        def arr(p, nbytes, dt):
            if nbytes == 0 or not p.value:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)),
                                         shape=(nbytes,)).copy().view(dt)
        return {
            "root": arr(ptrs[0], 1024, np.int32),
            "pre2": arr(ptrs[1], 8192, np.uint32),
            "rank2": arr(ptrs[2], 4096, np.uint16),
            "next2": arr(ptrs[3], max(info["pre2_bits_set"], 1) * 4, np.uint32),
            "hot": arr(ptrs[4], info["hot_buckets"] * 16, np.uint32).reshape(-1, 4),
            "cold": arr(ptrs[5], info["cold_buckets"] * 16, np.uint32).reshape(-1, 4),
            "chains": arr(ptrs[6], max(info["num_chains"], 1) * 16, np.uint32).reshape(-1, 4),
            "tails": arr(ptrs[7], info["tail_bytes"], np.uint8),
            "lut": arr(p2[0], 256, np.uint8),
            "best2": arr(p2[1], max(info["pre2_bits_set"], 1) * 4 if info["has_best2"] else 0, np.uint32),
            "chk2": arr(p2[2], max(info["pre2_bits_set"], 1) * 2 if info["has_chk2"] else 0, np.uint16),
            "hfilt": arr(pf, 4 * info["hfilt_words"], np.uint32),
            "hfilt_k": info["hashed_filter"], "code_shift": info["code_shift"],
            "code_bits": info["code_bits"], "gram_len": info["gram_len"],
            "hot_depth": info["hot_depth"], "mul": info["hash_mul"],
        }
