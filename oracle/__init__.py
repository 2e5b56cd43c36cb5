"""ctypes loaders for the two CPU checkers in this directory.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under pfac_b200/ imports this package.

  Oracle     -> oracle/libpfac_oracle.so   (plain-C restatement, oracle/pfac_oracle.c)
  RefOracle  -> oracle/_ref/libpfac_ref.so (the reference's own CPU path, oracle/ref_driver.cpp)

Both expose: num_patterns, num_states, initial_state, max_pattern_len, dense_table(),
match(text, omp=...), reduce(dense), dump(path).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libpfac_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpfac_ref.so")


def build(verbose=False):
    """Compile the C restatement, and oracle/_ref when /root/reference is present."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def ref_available():
    return os.path.exists(REF_SO)


def _as_u8(text):
    if isinstance(text, (bytes, bytearray)):
        text = np.frombuffer(bytes(text), dtype=np.uint8)
    text = np.ascontiguousarray(text, dtype=np.uint8)
    return text


def reduce_dense(dense):
    """The reduce oracle (reference src/PFAC.cpp:1058-1068): non-zero (id, pos), ascending pos."""
    dense = np.asarray(dense)
    pos = np.flatnonzero(dense > 0)
    return dense[pos].astype(np.int32), pos.astype(np.int64)


class Oracle:
    """Plain-C restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build()
            L = ctypes.CDLL(ORACLE_SO)
            L.orc_build_from_file.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
            L.orc_build_from_memory.argtypes = [ctypes.c_char_p, ctypes.c_long,
                                                ctypes.POINTER(ctypes.c_void_p)]
            L.orc_free.argtypes = [ctypes.c_void_p]
            L.orc_free.restype = None
            for f in ("orc_num_patterns", "orc_num_states", "orc_initial_state",
                      "orc_max_pattern_len"):
                getattr(L, f).argtypes = [ctypes.c_void_p]
                getattr(L, f).restype = ctypes.c_int
            L.orc_dense_table.argtypes = [ctypes.c_void_p]
            L.orc_dense_table.restype = ctypes.POINTER(ctypes.c_int)
            for f in ("orc_match", "orc_match_omp"):
                getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_void_p]
                getattr(L, f).restype = None
            L.orc_match_shard.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_void_p]
            L.orc_match_shard.restype = None
            L.orc_reduce.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                     ctypes.c_void_p]
            L.orc_reduce.restype = ctypes.c_int64
            L.orc_dump.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
            L.orc_omp_threads.restype = ctypes.c_int
            L.orc_set_threads.argtypes = [ctypes.c_int]
            L.orc_set_threads.restype = None
            cls._lib = L
        return cls._lib

    def __init__(self, pattern_file=None, image=None):
        L = self.lib()
        self._h = ctypes.c_void_p()
        if pattern_file is not None:
            rc = L.orc_build_from_file(os.fsencode(pattern_file), ctypes.byref(self._h))
        else:
            rc = L.orc_build_from_memory(image, len(image), ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise ValueError("oracle build failed with PFAC status %d" % rc)

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().orc_free(self._h)
            self._h = None

    num_patterns = property(lambda s: s.lib().orc_num_patterns(s._h))
    num_states = property(lambda s: s.lib().orc_num_states(s._h))
    initial_state = property(lambda s: s.lib().orc_initial_state(s._h))
    max_pattern_len = property(lambda s: s.lib().orc_max_pattern_len(s._h))

    @staticmethod
    def threads():
        return Oracle.lib().orc_omp_threads()

    @staticmethod
    def set_threads(n):
        Oracle.lib().orc_set_threads(int(n))

    def dense_table(self):
        p = self.lib().orc_dense_table(self._h)
        return np.ctypeslib.as_array(p, shape=(self.num_states, 256)).copy()

    def match(self, text, omp=True):
        text = _as_u8(text)
        out = np.empty(text.size, dtype=np.int32)
        if text.size:
            f = self.lib().orc_match_omp if omp else self.lib().orc_match
            f(self._h, text.ctypes.data, text.size, out.ctypes.data)
        return out

    def match_shard(self, text, n_owned):
        """Report positions [0,n_owned); the walk may read all of `text` (owned + halo)."""
        text = _as_u8(text)
        out = np.empty(n_owned, dtype=np.int32)
        if n_owned:
            self.lib().orc_match_shard(self._h, text.ctypes.data, n_owned, text.size,
                                       out.ctypes.data)
        return out

    def reduce(self, dense):
        dense = np.ascontiguousarray(dense, dtype=np.int32)
        ids = np.empty(dense.size, dtype=np.int32)
        pos = np.empty(dense.size, dtype=np.int64)
        m = self.lib().orc_reduce(dense.ctypes.data, dense.size, ids.ctypes.data, pos.ctypes.data)
        return ids[:m].copy(), pos[:m].copy()

    def dump(self, path):
        rc = self.lib().orc_dump(self._h, os.fsencode(path))
        if rc != 0:
            raise IOError("orc_dump failed: %d" % rc)


class RefOracle:
    """The reference's own CPU matcher (PFAC_CPU / PFAC_CPU_OMP), built by oracle/Makefile."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (build it where /root/reference exists: make -C oracle)")
            L = ctypes.CDLL(REF_SO)
            L.ref_create.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
            L.ref_destroy.argtypes = [ctypes.c_void_p]
            L.ref_destroy.restype = None
            L.ref_match.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_int]
            L.ref_dump.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
            for f in ("ref_num_patterns", "ref_num_states", "ref_initial_state",
                      "ref_max_pattern_len"):
                getattr(L, f).argtypes = [ctypes.c_void_p]
                getattr(L, f).restype = ctypes.c_int
            L.ref_dense_table.argtypes = [ctypes.c_void_p]
            L.ref_dense_table.restype = ctypes.POINTER(ctypes.c_int)
            L.ref_set_threads.argtypes = [ctypes.c_int]
            L.ref_set_threads.restype = None
            L.ref_get_threads.restype = ctypes.c_int
            cls._lib = L
        return cls._lib

    @staticmethod
    def threads():
        return RefOracle.lib().ref_get_threads()

    @staticmethod
    def set_threads(n):
        RefOracle.lib().ref_set_threads(int(n))

    def __init__(self, pattern_file):
        L = self.lib()
        self._h = ctypes.c_void_p()
        rc = L.ref_create(os.fsencode(pattern_file), ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise ValueError("reference build failed with PFAC status %d" % rc)

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().ref_destroy(self._h)
            self._h = None

    num_patterns = property(lambda s: s.lib().ref_num_patterns(s._h))
    num_states = property(lambda s: s.lib().ref_num_states(s._h))
    initial_state = property(lambda s: s.lib().ref_initial_state(s._h))
    max_pattern_len = property(lambda s: s.lib().ref_max_pattern_len(s._h))

    def dense_table(self):
        p = self.lib().ref_dense_table(self._h)
        return np.ctypeslib.as_array(p, shape=(self.num_states, 256)).copy()

    def match(self, text, omp=True):
        """omp=True -> PFAC_CPU_OMP (threads = OMP_NUM_THREADS or all cores).  n < 2**31."""
        text = _as_u8(text)
        assert text.size < 2 ** 31, "reference matcher takes int input_size"
        out = np.empty(text.size, dtype=np.int32)
        if text.size:
            rc = self.lib().ref_match(self._h, text.ctypes.data, text.size, out.ctypes.data,
                                      1 if omp else 0)
            if rc != 0:
                raise RuntimeError("reference matcher returned %d" % rc)
        return out

    def reduce(self, dense):
        return reduce_dense(dense)

    def dump(self, path):
        rc = self.lib().ref_dump(self._h, os.fsencode(path))
        if rc != 0:
            raise IOError("ref_dump failed: %d" % rc)
