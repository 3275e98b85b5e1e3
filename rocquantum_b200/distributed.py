"""One-process-per-GPU front end of the distributed state (rocsvAllocateDistributedState & co).

torch.distributed is plumbing only: it carries NCCL's 128-byte unique id from rank 0 to the other ranks.  The
slices, the index-bit exchanges and every collective on the data path live in the C library (csrc/dist.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .statevec import DT, RocsvError


class DistStateVector:
    def __init__(self, n_total: int, prec: str = "c64", seed: int = 0):
        import torch
        import torch.distributed as dist
        self.lib = capi.load(prec)
        self.prec, self.dtype, self.n = prec, DT[prec], n_total
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            self._ck("rocsvxDistGetUniqueId", self.lib.rocsvxDistGetUniqueId(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            ident = ident.cuda()
        dist.broadcast(ident, 0)
        raw = bytes(ident.cpu().tolist())
        self.h = C.c_void_p()
        self._ck("rocsvCreate", self.lib.rocsvCreate(C.byref(self.h)))
        self._ck("rocsvxDistInit", self.lib.rocsvxDistInit(self.h, self.rank, self.world, C.c_char_p(raw)))
        self._ck("rocsvAllocateDistributedState", self.lib.rocsvAllocateDistributedState(self.h, n_total))
        self._ck("rocsvInitializeDistributedState", self.lib.rocsvInitializeDistributedState(self.h))
        nl = C.c_uint()
        self.lib.rocsvxDistGetInfo(self.h, None, None, C.byref(nl), None)
        self.n_local = nl.value
        self._ck("rocsvxSetSeed", self.lib.rocsvxSetSeed(self.h, seed))     # fixed stream (handles seed from std::random_device)

    @staticmethod
    def _ck(fn, st):
        if st != capi.SUCCESS:
            raise RocsvError(fn, st)

    def close(self):
        if self.h:
            self.lib.rocsvDestroy(self.h)
            self.h = C.c_void_p()

    def init(self):
        self._ck("rocsvInitializeDistributedState", self.lib.rocsvInitializeDistributedState(self.h))

    def set_fusion(self, on):
        self._ck("rocsvxSetFusion", self.lib.rocsvxSetFusion(self.h, int(on)))

    def sync(self):
        self._ck("rocsvxSynchronize", self.lib.rocsvxSynchronize(self.h))

    def stats(self, reset=False):
        s = capi.Stats()
        self._ck("rocsvxGetStats", self.lib.rocsvxGetStats(self.h, C.byref(s), int(reset)))
        return s

    def timer_start(self):
        self._ck("rocsvxTimerStart", self.lib.rocsvxTimerStart(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._ck("rocsvxTimerStop", self.lib.rocsvxTimerStop(self.h, C.byref(ms)))
        return ms.value

    def apply_ops(self, arr, count):
        self._ck("rocsvxApplyCircuit", self.lib.rocsvxApplyCircuit(self.h, None, self.n, arr, count))

    def apply_circuit(self, gates):
        gates = list(gates)
        arr, keep = capi.make_ops(gates)
        self.apply_ops(arr, len(gates))
        del keep

    def gate_status(self, name, *a):
        L, h, n = self.lib, self.h, self.n
        name = name.lower()
        t1 = {"h": "H", "x": "X", "y": "Y", "z": "Z", "s": "S", "sdg": "Sdg", "t": "T"}
        if name in t1:
            return getattr(L, "rocsvApply" + t1[name])(h, None, n, a[0])
        if name in ("rx", "ry", "rz"):
            return getattr(L, "rocsvApplyR" + name[1])(h, None, n, a[0], a[1])
        if name in ("cnot", "cz", "swap"):
            return getattr(L, "rocsvApply" + name.upper())(h, None, n, a[0], a[1])
        if name in ("crx", "cry", "crz"):
            return getattr(L, "rocsvApply" + name.upper())(h, None, n, a[0], a[1], a[2])
        if name == "mcx":
            return L.rocsvApplyMultiControlledX(h, None, n, capi.uarr(a[0]), len(a[0]), a[1])
        if name == "cswap":
            return L.rocsvApplyCSWAP(h, None, n, a[0], a[1], a[2])
        raise ValueError(name)

    def gate(self, name, *a):
        self._ck(name, self.gate_status(name, *a))

    def swap_index_bits(self, a, b):
        self._ck("rocsvSwapIndexBits", self.lib.rocsvSwapIndexBits(self.h, a, b))

    def set_local_slice(self, v):
        v = np.ascontiguousarray(v, dtype=self.dtype)
        assert v.size == 1 << self.n_local
        self._ck("rocsvxSetStateVector", self.lib.rocsvxSetStateVector(self.h, None, v.ctypes.data_as(C.c_void_p)))

    def local_slice(self) -> np.ndarray:
        """This rank's 2^n_local amplitudes in the canonical layout global = (rank << n_local) | local."""
        out = np.empty(1 << self.n_local, dtype=self.dtype)
        self._ck("rocsvGetStateVectorFull", self.lib.rocsvGetStateVectorFull(self.h, None, out.ctypes.data_as(C.c_void_p)))
        return out

    def expect_pauli(self, paulis, qubits):
        r = C.c_double()
        self._ck("rocsvGetExpectationPauliString",
                 self.lib.rocsvGetExpectationPauliString(self.h, None, self.n, paulis.encode(), capi.uarr(qubits), len(qubits), C.byref(r)))
        return r.value

    def expect_z(self, q):
        return self.expect_pauli("Z", [q])

    def measure(self, q):
        o, p = C.c_int(), C.c_double()
        self._ck("rocsvMeasure", self.lib.rocsvMeasure(self.h, None, self.n, q, C.byref(o), C.byref(p)))
        return o.value, p.value

    def sample(self, qubits, shots):
        out = np.zeros(max(1, shots), dtype=np.uint64)
        self._ck("rocsvSample", self.lib.rocsvSample(self.h, None, self.n, capi.uarr(qubits), len(qubits), shots,
                                                     out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out[:shots]
