"""A small object wrapper over the C ABI: one handle + one owned state, gates by the reference's names.

Used by the parity tests, bench.py and smoke().  Every method is one rocsv*/rocsvx* call; nothing is
computed in Python."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

DT = {"c64": np.complex64, "c128": np.complex128}


class RocsvError(RuntimeError):
    def __init__(self, fn, status):
        super().__init__(f"{fn} failed: {capi.STATUS_NAMES[status] if 0 <= status < 7 else status}")
        self.status = status


class StateVector:
    def __init__(self, n: int, prec: str = "c64", batch: int = 1, fusion: bool = False, seed: int = 0, ranks: int | None = None):
        """ranks=None: a plain state (rocsvAllocateState).  ranks=P: the reference's single-process multi-GPU flow --
        rocsvAllocateDistributedState on ONE handle, P slices (0 = one per visible device; more slices than devices are
        placed round-robin, which is how the 1-GPU test box exercises the distributed engine); d_state is NULL on every call."""
        self.lib = capi.load(prec)
        self.prec, self.dtype, self.n, self.batch = prec, DT[prec], n, batch
        self.h = C.c_void_p()
        self._ck("rocsvCreate", self.lib.rocsvCreate(C.byref(self.h)))
        self.d = C.c_void_p()
        if ranks is None:
            self._ck("rocsvAllocateState", self.lib.rocsvAllocateState(self.h, n, C.byref(self.d), batch))
            self._ck("rocsvInitializeState", self.lib.rocsvInitializeState(self.h, self.d, n))
        else:
            assert batch == 1
            self._ck("rocsvxDistSetRanks", self.lib.rocsvxDistSetRanks(self.h, ranks))
            self._ck("rocsvAllocateDistributedState", self.lib.rocsvAllocateDistributedState(self.h, n))
            self._ck("rocsvInitializeDistributedState", self.lib.rocsvInitializeDistributedState(self.h))
        if fusion:
            self.set_fusion(True)
        self.set_seed(seed)           # handles seed themselves from std::random_device; tests and benches want a fixed stream

    @staticmethod
    def _ck(fn, st):
        if st != capi.SUCCESS:
            raise RocsvError(fn, st)

    def close(self):
        if self.h:
            self.lib.rocsvDestroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- control -------------------------------------------------------------------------------
    def init(self):
        self._ck("rocsvInitializeState", self.lib.rocsvInitializeState(self.h, self.d, self.n))

    def set_fusion(self, on: bool):
        self._ck("rocsvxSetFusion", self.lib.rocsvxSetFusion(self.h, int(on)))

    def set_seed(self, seed: int):
        self._ck("rocsvxSetSeed", self.lib.rocsvxSetSeed(self.h, seed))

    def dist_info(self):
        """-> (rank, ranks, local qubits) of the handle's distributed state"""
        r, p, nl = C.c_int(), C.c_int(), C.c_uint()
        self._ck("rocsvxDistGetInfo", self.lib.rocsvxDistGetInfo(self.h, C.byref(r), C.byref(p), C.byref(nl), None))
        return r.value, p.value, nl.value

    def rank_slice(self, rank: int) -> np.ndarray:
        """the raw device slice of one rank, as the reference's multi-GPU test reads it (test_hipStateVec_multi_gpu.cpp:43-81)"""
        import ctypes
        dev, ptr = C.c_int(), C.c_void_p()
        self._ck("rocsvxDistGetRankSlice", self.lib.rocsvxDistGetRankSlice(self.h, rank, C.byref(dev), C.byref(ptr)))
        self.sync()
        _, _, nl = self.dist_info()
        out = np.empty(1 << nl, dtype=self.dtype)
        rt = ctypes.CDLL("libcudart.so.12")
        assert rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), ptr, C.c_size_t(out.nbytes), C.c_int(2)) == 0
        return out

    def sync(self):
        self._ck("rocsvxSynchronize", self.lib.rocsvxSynchronize(self.h))

    def flush(self):
        self._ck("rocsvxFlush", self.lib.rocsvxFlush(self.h))

    def stats(self, reset=False) -> capi.Stats:
        s = capi.Stats()
        self._ck("rocsvxGetStats", self.lib.rocsvxGetStats(self.h, C.byref(s), int(reset)))
        return s

    def timer_start(self):
        self._ck("rocsvxTimerStart", self.lib.rocsvxTimerStart(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._ck("rocsvxTimerStop", self.lib.rocsvxTimerStop(self.h, C.byref(ms)))
        return ms.value

    def set_state(self, v):
        v = np.ascontiguousarray(v, dtype=self.dtype)
        assert v.size == self.batch << self.n
        self._ck("rocsvxSetStateVector", self.lib.rocsvxSetStateVector(self.h, self.d, v.ctypes.data_as(C.c_void_p)))

    def state(self) -> np.ndarray:
        out = np.empty(self.batch << self.n, dtype=self.dtype)
        self._ck("rocsvGetStateVectorFull", self.lib.rocsvGetStateVectorFull(self.h, self.d, out.ctypes.data_as(C.c_void_p)))
        return out

    def state_slice(self, b: int) -> np.ndarray:
        out = np.empty(1 << self.n, dtype=self.dtype)
        self._ck("rocsvGetStateVectorSlice", self.lib.rocsvGetStateVectorSlice(self.h, self.d, out.ctypes.data_as(C.c_void_p), b))
        return out

    # ---- gates, by the reference's names ---------------------------------------------------------
    def gate_status(self, name, *a) -> int:
        L, h, d, n = self.lib, self.h, self.d, self.n
        name = name.lower()
        t1 = {"h": "H", "x": "X", "y": "Y", "z": "Z", "s": "S", "sdg": "Sdg", "t": "T"}
        if name in t1:
            return getattr(L, "rocsvApply" + t1[name])(h, d, n, a[0])
        if name in ("rx", "ry", "rz"):
            return getattr(L, "rocsvApplyR" + name[1])(h, d, n, a[0], a[1])
        if name in ("cnot", "cz", "swap"):
            return getattr(L, "rocsvApply" + name.upper())(h, d, n, a[0], a[1])
        if name in ("crx", "cry", "crz"):
            return getattr(L, "rocsvApply" + name.upper())(h, d, n, a[0], a[1], a[2])
        if name == "mcx":
            return L.rocsvApplyMultiControlledX(h, d, n, capi.uarr(a[0]), len(a[0]), a[1])
        if name == "cswap":
            return L.rocsvApplyCSWAP(h, d, n, a[0], a[1], a[2])
        raise ValueError(name)

    def gate(self, name, *a):
        self._ck(name, self.gate_status(name, *a))

    def _device_matrix(self, M, k):
        """Upload a (2^k,2^k) matrix column-major through the handle's pinned buffer -> returns device ptr owner."""
        import torch  # device memory plumbing only
        D = 1 << k
        Mc = np.ascontiguousarray(np.asarray(M, dtype=self.dtype).reshape(D, D).T).reshape(-1)
        t = torch.from_numpy(Mc.view(np.float32 if self.prec == "c64" else np.float64).copy()).cuda()
        return t

    def apply_matrix(self, targets, M, controls=()):
        k = len(targets)
        t = self._device_matrix(M, k)
        if controls:
            st = self.lib.rocsvApplyControlledMatrix(self.h, self.d, self.n, capi.uarr(controls), len(controls), capi.uarr(targets), k,
                                                     C.c_void_p(t.data_ptr()))
        else:
            st = self.lib.rocsvApplyMatrix(self.h, self.d, self.n, capi.uarr(targets), k, C.c_void_p(t.data_ptr()), 1 << k)
        self.sync()          # the device matrix must outlive the launch
        self._ck("rocsvApplyMatrix", st)

    def apply_fused_1q(self, target, M):
        t = self._device_matrix(M, 1)
        st = self.lib.rocsvApplyFusedSingleQubitMatrix(self.h, target, C.c_void_p(t.data_ptr()))
        self.sync()
        self._ck("rocsvApplyFusedSingleQubitMatrix", st)

    def apply_block6(self, qubits, M):
        """Dense 6-qubit matrix through the tensor-core block sweep (complex64 only).  M[i][j]: row i, column j."""
        Mc = np.ascontiguousarray(np.asarray(M, dtype=np.complex128).reshape(64, 64).T).reshape(-1)   # column-major
        buf = np.empty(2 * Mc.size, dtype=np.float64)
        buf[0::2], buf[1::2] = Mc.real, Mc.imag
        self._ck("rocsvxApplyBlock6", self.lib.rocsvxApplyBlock6(self.h, self.d, self.n, capi.uarr(qubits), buf.ctypes.data_as(C.POINTER(C.c_double))))

    def set_tensor_core_blocks(self, on: bool):
        self._ck("rocsvxSetTensorCoreBlocks", self.lib.rocsvxSetTensorCoreBlocks(self.h, int(on)))

    def set_merge_diagonals(self, on: bool):
        self._ck("rocsvxSetMergeDiagonals", self.lib.rocsvxSetMergeDiagonals(self.h, int(on)))

    def swap_index_bits(self, a, b):
        self._ck("rocsvSwapIndexBits", self.lib.rocsvSwapIndexBits(self.h, a, b))

    def apply_circuit(self, gates):
        arr, keep = capi.make_ops(gates)
        self._ck("rocsvxApplyCircuit", self.lib.rocsvxApplyCircuit(self.h, self.d, self.n, arr, len(list(gates)) if not isinstance(gates, list) else len(gates)))
        del keep

    # ---- reductions ----------------------------------------------------------------------------------
    def norm2(self) -> float:
        r = C.c_double()
        self._ck("rocsvxGetNorm", self.lib.rocsvxGetNorm(self.h, self.d, self.n, C.byref(r)))
        return r.value

    def expect_pauli(self, paulis: str, qubits) -> float:
        r = C.c_double()
        self._ck("rocsvGetExpectationPauliString",
                 self.lib.rocsvGetExpectationPauliString(self.h, self.d, self.n, paulis.encode(), capi.uarr(qubits), len(qubits), C.byref(r)))
        return r.value

    def expect_z(self, q):
        r = C.c_double()
        self._ck("Z", self.lib.rocsvGetExpectationValueSinglePauliZ(self.h, self.d, self.n, q, C.byref(r)))
        return r.value

    def expect_x(self, q):
        r = C.c_double()
        self._ck("X", self.lib.rocsvGetExpectationValueSinglePauliX(self.h, self.d, self.n, q, C.byref(r)))
        return r.value

    def expect_y(self, q):
        r = C.c_double()
        self._ck("Y", self.lib.rocsvGetExpectationValueSinglePauliY(self.h, self.d, self.n, q, C.byref(r)))
        return r.value

    def expect_zprod(self, qubits):
        r = C.c_double()
        self._ck("ZZ", self.lib.rocsvGetExpectationValuePauliProductZ(self.h, self.d, self.n, capi.uarr(qubits), len(qubits), C.byref(r)))
        return r.value

    def measure(self, q):
        o, p = C.c_int(), C.c_double()
        self._ck("rocsvMeasure", self.lib.rocsvMeasure(self.h, self.d, self.n, q, C.byref(o), C.byref(p)))
        return o.value, p.value

    def apply_matrix_and_measure(self, targets, M, q):
        """rocsvApplyMatrixAndMeasure (hipStateVec.h:487-494): the matrix on `targets`, then a projective measurement of q."""
        k = len(targets)
        t = self._device_matrix(M, k)
        o = C.c_int()
        st = self.lib.rocsvApplyMatrixAndMeasure(self.h, self.d, self.n, capi.uarr(targets), k, C.c_void_p(t.data_ptr()), q, C.byref(o))
        self.sync()
        self._ck("rocsvApplyMatrixAndMeasure", st)
        return o.value

    def expect_batch(self, terms, all_states=False) -> np.ndarray:
        """terms: [(pauli string, qubits)] -> rocsvxGetExpectationPauliBatch[AllStates]; shape (terms,) or (batch, terms)."""
        paulis = "".join(t[0] for t in terms).encode()
        qubits = capi.uarr([q for t in terms for q in t[1]])
        offs, acc = [0], 0
        for t in terms:
            acc += len(t[0]); offs.append(acc)
        ns = self.batch if all_states else 1
        res = (C.c_double * max(1, len(terms) * ns))()
        fn = self.lib.rocsvxGetExpectationPauliBatchAllStates if all_states else self.lib.rocsvxGetExpectationPauliBatch
        self._ck("rocsvxGetExpectationPauliBatch", fn(self.h, self.d, self.n, paulis, qubits, capi.uarr(offs), len(terms), res))
        out = np.array(res[:len(terms) * ns])
        return out.reshape(ns, len(terms)) if all_states else out

    def sample(self, qubits, shots) -> np.ndarray:
        out = np.zeros(max(1, shots), dtype=np.uint64)
        self._ck("rocsvSample", self.lib.rocsvSample(self.h, self.d, self.n, capi.uarr(qubits), len(qubits), shots,
                                                     out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out[:shots]
