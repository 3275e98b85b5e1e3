"""oracle/sv_oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end to the CPU restatement (oracle/sv_oracle.c) and, when it has been built
(`make -C oracle ref`, needs /root/reference), to the reference's own sources compiled under
the host HIP shim (oracle/_ref/libhipStateVec_ref_{c64,c128}.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package rocquantum_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DT = {"c64": np.complex64, "c128": np.complex128}


def build(ref: bool = True) -> None:
    """Compile the C restatement, and oracle/_ref when /root/reference is present."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/rocquantum/src/hipStateVec"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libsv_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _LIB = C.CDLL(path)
        _LIB.orc_uniform53.restype = C.c_uint64
        _LIB.orc_uniform53.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        _LIB.orc_gate_matrix.restype = C.c_int
        for sfx in ("c64", "c128"):
            getattr(_LIB, f"orc_norm2_{sfx}").restype = C.c_double
            getattr(_LIB, f"orc_expect_pauli_{sfx}").restype = C.c_double
    return _LIB


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def uniform53(seed: int, call: int, shot: int) -> int:
    return int(lib().orc_uniform53(seed, call, shot))


def fix88(p: float) -> int:
    hi, lo = C.c_uint64(), C.c_uint64()
    lib().orc_fix88_parts(C.c_double(p), C.byref(hi), C.byref(lo))
    return (hi.value << 64) | lo.value


def gate_matrix(name: str, theta: float = 0.0) -> np.ndarray:
    """2x2 complex128, row-major [[m00,m01],[m10,m11]], as hipStateVec.cpp builds it."""
    m = (C.c_double * 8)()
    if lib().orc_gate_matrix(name.lower().encode(), C.c_double(theta), m) != 0:
        raise ValueError(f"unknown gate {name}")
    v = np.array(list(m), dtype=np.float64)
    return (v[0::2] + 1j * v[1::2]).reshape(2, 2)


def _uarr(xs):
    return (C.c_uint * max(1, len(xs)))(*xs)


class Oracle:
    """State vector on the host, driven with the reference's operation names."""

    def __init__(self, n: int, prec: str = "c64", batch: int = 1, seed: int = 0):
        self.n, self.prec, self.batch, self.seed = n, prec, batch, seed
        self.call = 0
        self.dtype = DT[prec]
        self.state = np.zeros(batch << n, dtype=self.dtype)
        self._f = lambda name: getattr(lib(), f"{name}_{prec}")
        self.init()

    def _p(self):
        return self.state.ctypes.data_as(C.c_void_p)

    def init(self):
        self._f("orc_init_state")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch))

    def set_state(self, v):
        self.state[:] = np.asarray(v, dtype=self.dtype)

    # --- gates ---------------------------------------------------------------------------
    def matrix1(self, t, m2x2):
        m = np.asarray(m2x2, dtype=np.complex128).reshape(4)
        flat = np.empty(8)
        flat[0::2], flat[1::2] = m.real, m.imag
        self._f("orc_apply_matrix1")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch), C.c_uint(t),
                                     flat.ctypes.data_as(C.POINTER(C.c_double)))

    def cmatrix1(self, c, t, m2x2):
        m = np.asarray(m2x2, dtype=np.complex128).reshape(4)
        flat = np.empty(8)
        flat[0::2], flat[1::2] = m.real, m.imag
        self._f("orc_apply_cmatrix1")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch), C.c_uint(c), C.c_uint(t),
                                      flat.ctypes.data_as(C.POINTER(C.c_double)))

    def gate(self, name, *args):
        """name in h,x,y,z,s,sdg,t (t) | rx,ry,rz (t,theta) | cnot,cz,swap (a,b) |
        crx,cry,crz (c,t,theta) | mcx ([c..],t) | cswap (c,a,b)"""
        name = name.lower()
        if name in ("h", "x", "y", "z", "s", "sdg", "t"):
            self.matrix1(args[0], gate_matrix(name))
        elif name in ("rx", "ry", "rz"):
            self.matrix1(args[0], gate_matrix(name, args[1]))
        elif name in ("crx", "cry", "crz"):
            self.cmatrix1(args[0], args[1], gate_matrix(name[1:], args[2]))
        elif name == "cnot":
            self.mcx([args[0]], args[1])
        elif name == "cz":
            self._f("orc_cz")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch), C.c_uint(args[0]), C.c_uint(args[1]))
        elif name == "swap":
            self._f("orc_cswap")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch), C.c_ulonglong(0),
                                 C.c_uint(args[0]), C.c_uint(args[1]))
        elif name == "mcx":
            self.mcx(args[0], args[1])
        elif name == "cswap":
            self._f("orc_cswap")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch), C.c_ulonglong(1 << args[0]),
                                 C.c_uint(args[1]), C.c_uint(args[2]))
        else:
            raise ValueError(name)

    def mcx(self, controls, t):
        mask = 0
        for c in controls:
            mask |= 1 << c
        self._f("orc_mcx")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch), C.c_ulonglong(mask), C.c_uint(t))

    def apply_matrix(self, targets, M, controls=()):
        """M: (2^k,2^k) array, M[i,j] is row i col j (sent column-major like the C API)."""
        k = len(targets)
        Mc = np.asfortranarray(np.asarray(M, dtype=self.dtype).reshape(1 << k, 1 << k))
        buf = np.ascontiguousarray(Mc.T).reshape(-1)  # column-major linearisation
        self._f("orc_apply_matrix")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch), _uarr(targets), C.c_uint(k),
                                    _uarr(controls), C.c_uint(len(controls)), buf.ctypes.data_as(C.c_void_p))

    def swap_index_bits(self, a, b):
        self._f("orc_swap_index_bits")(self._p(), C.c_uint(self.n), C.c_size_t(self.batch), C.c_uint(a), C.c_uint(b))

    # --- reductions -------------------------------------------------------------------------
    def norm2(self):
        return float(self._f("orc_norm2")(self._p(), C.c_uint(self.n)))

    def expect_pauli(self, paulis: str, qubits):
        return float(self._f("orc_expect_pauli")(self._p(), C.c_uint(self.n), paulis.encode(), _uarr(qubits),
                                                 C.c_uint(len(qubits))))

    def measure(self, q):
        out, prob = C.c_int(), C.c_double()
        self._f("orc_measure")(self._p(), C.c_uint(self.n), C.c_uint(q), C.c_uint64(self.seed), C.c_uint64(self.call),
                               C.byref(out), C.byref(prob))
        self.call += 1
        return out.value, prob.value

    def sample(self, qubits, shots):
        res = np.zeros(max(1, shots), dtype=np.uint64)
        self._f("orc_sample")(self._p(), C.c_uint(self.n), _uarr(qubits), C.c_uint(len(qubits)), C.c_uint(shots),
                              C.c_uint64(self.seed), C.c_uint64(self.call), res.ctypes.data_as(C.c_void_p))
        self.call += 1
        return res[:shots]


# ------------------------------------------------------------------------------------------------
# the reference's own code, compiled under the host shim (25 defined entry points)
# ------------------------------------------------------------------------------------------------
class RefLib:
    """rocsv* API of oracle/_ref (the reference's hipStateVec.cpp + kernels on host threads)."""

    def __init__(self, prec: str = "c64"):
        path = os.path.join(_HERE, "_ref", f"libhipStateVec_ref_{prec}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.prec, self.dtype = prec, DT[prec]
        self.lib = C.CDLL(path)
        self.h = C.c_void_p()
        assert self.lib.rocsvCreate(C.byref(self.h)) == 0
        self.d = C.c_void_p()
        self.n = 0
        self.batch = 1

    def allocate(self, n, batch=1):
        self.n, self.batch = n, batch
        assert self.lib.rocsvAllocateState(self.h, C.c_uint(n), C.byref(self.d), C.c_size_t(batch)) == 0
        assert self.lib.rocsvInitializeState(self.h, self.d, C.c_uint(n)) == 0

    def set_state(self, v):
        v = np.ascontiguousarray(v, dtype=self.dtype)
        C.memmove(self.d, v.ctypes.data, v.nbytes)  # "device" memory is host heap under the shim

    def call(self, fn, *args):
        return getattr(self.lib, fn)(self.h, self.d, C.c_uint(self.n), *args)

    def gate(self, name, *a):
        u, dbl = C.c_uint, C.c_double
        name = name.lower()
        table1 = {"h": "H", "x": "X", "y": "Y", "z": "Z", "s": "S", "sdg": "Sdg", "t": "T"}
        if name in table1:
            st = self.call("rocsvApply" + table1[name], u(a[0]))
        elif name in ("rx", "ry", "rz"):
            st = self.call("rocsvApplyR" + name[1], u(a[0]), dbl(a[1]))
        elif name in ("cnot", "cz", "swap"):
            st = self.call("rocsvApply" + name.upper(), u(a[0]), u(a[1]))
        elif name in ("crx", "cry", "crz"):
            st = self.call("rocsvApply" + name.upper(), u(a[0]), u(a[1]), dbl(a[2]))
        elif name == "mcx":
            st = self.call("rocsvApplyMultiControlledX", _uarr(a[0]), u(len(a[0])), u(a[1]))
        elif name == "cswap":
            st = self.call("rocsvApplyCSWAP", u(a[0]), u(a[1]), u(a[2]))
        else:
            raise ValueError(name)
        return st

    def spec_apply_matrix(self, targets, M):
        """The reference's own (never launched) ApplyMatrix kernel, multi_qubit_kernels.hip:37-115, driven by
        oracle/hip_shim/spec_driver.cpp: 1 <= len(targets) <= 4, batch 1.  M[i, j] = row i, column j; it is handed over
        column-major as hipStateVec.h:145-148 specifies."""
        assert self.batch == 1
        k = len(targets)
        buf = np.ascontiguousarray(np.asarray(M, dtype=self.dtype).reshape(1 << k, 1 << k).T).reshape(-1)
        st = self.lib.refspec_apply_matrix(self.d, C.c_uint(self.n), _uarr(targets), C.c_uint(k), buf.ctypes.data_as(C.c_void_p))
        assert st == 0, st

    def spec_measure_with_outcome(self, q, outcome):
        """prob0 -> collapse -> sum of squares -> renormalise by the reference's own measurement kernels
        (measurement_kernels.hip:12-99) for a given outcome; returns prob0 as the kernel summed it."""
        assert self.batch == 1
        p0 = C.c_double()
        st = self.lib.refspec_measure_with_outcome(self.d, C.c_uint(self.n), C.c_uint(q), C.c_int(outcome), C.byref(p0))
        assert st == 0, st
        return p0.value

    def spec_multi_z_probabilities(self, qubits):
        """Joint Z-basis outcome probabilities of up to 8 qubits (bin bit j <-> qubits[j]) by
        calculate_multi_z_probabilities_kernel + its reduction (measurement_kernels.hip:283-387)."""
        assert self.batch == 1
        probs = np.zeros(1 << len(qubits), dtype=np.float64)
        st = self.lib.refspec_multi_z_probabilities(self.d, C.c_uint(self.n), _uarr(qubits), C.c_uint(len(qubits)),
                                                    probs.ctypes.data_as(C.c_void_p))
        assert st == 0, st
        return probs

    def spec_local_bit_swap(self, q1, q2):
        """local_bit_swap_permutation_kernel (swap_kernels.hip:95-114) run as one real block of 2^n threads (n <= 8)."""
        assert self.batch == 1
        st = self.lib.refspec_local_bit_swap(self.d, C.c_uint(self.n), C.c_uint(q1), C.c_uint(q2))
        assert st == 0, st

    def state(self):
        out = np.empty(self.batch << self.n, dtype=self.dtype)
        assert self.lib.rocsvGetStateVectorFull(self.h, self.d, out.ctypes.data_as(C.c_void_p)) == 0
        return out

    def close(self):
        if self.h:
            self.lib.rocsvDestroy(self.h)
            self.h = C.c_void_p()


def ref_available(prec="c64") -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libhipStateVec_ref_{prec}.so"))
