"""ctypes binding of include/hipStateVec.h (the drop-in C ABI).

`load("c64")` / `load("c128")` return the CDLL with argtypes/restypes set for all 42 rocsv* symbols
of the reference header plus the rocsvx* extensions.  SYMBOLS lists every exported name; the CPU test
suite checks each one resolves.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

# rocqStatus_t (hipStateVec.h:22-31 of the reference)
SUCCESS, FAILURE, INVALID_VALUE, ALLOCATION_FAILED, HIP_ERROR, NOT_IMPLEMENTED, RCCL_ERROR = range(7)
STATUS_NAMES = ["SUCCESS", "FAILURE", "INVALID_VALUE", "ALLOCATION_FAILED", "HIP_ERROR", "NOT_IMPLEMENTED", "RCCL_ERROR"]

(ROCSVX_H, ROCSVX_X, ROCSVX_Y, ROCSVX_Z, ROCSVX_S, ROCSVX_SDG, ROCSVX_T, ROCSVX_RX, ROCSVX_RY, ROCSVX_RZ, ROCSVX_CNOT,
 ROCSVX_CZ, ROCSVX_SWAP, ROCSVX_CRX, ROCSVX_CRY, ROCSVX_CRZ, ROCSVX_MCX, ROCSVX_CSWAP, ROCSVX_MATRIX) = range(19)


class GateOp(C.Structure):
    """rocsvxGateOp"""
    _fields_ = [("kind", C.c_int32), ("numTargets", C.c_uint32), ("targets", C.c_uint32 * 8), ("controlMask", C.c_uint64),
                ("theta", C.c_double), ("matrix", C.POINTER(C.c_double))]


class Stats(C.Structure):
    """rocsvxStats"""
    _fields_ = [("kernelLaunches", C.c_uint64), ("sweeps", C.c_uint64), ("gatesSubmitted", C.c_uint64),
                ("opsExecuted", C.c_uint64), ("h2dBytes", C.c_uint64), ("lastSweepMs", C.c_double),
                ("exchanges", C.c_uint64), ("exchangeBytes", C.c_uint64), ("exchangeMs", C.c_double), ("blockSweeps", C.c_uint64),
                ("planCacheHits", C.c_uint64), ("expectationSweeps", C.c_uint64)]


class ExchangeSeg(C.Structure):
    """rocsvxExchangeSeg"""
    _fields_ = [("peer", C.c_int32), ("sendOffset", C.c_uint64), ("recvOffset", C.c_uint64), ("count", C.c_uint64)]


_h, _p, _u, _d, _sz = C.c_void_p, C.c_void_p, C.c_uint, C.c_double, C.c_size_t
_up = C.POINTER(C.c_uint)

_GATE1 = ["rocsvApplyX", "rocsvApplyY", "rocsvApplyZ", "rocsvApplyH", "rocsvApplyS", "rocsvApplyT", "rocsvApplySdg"]
_ROT1 = ["rocsvApplyRx", "rocsvApplyRy", "rocsvApplyRz"]
_GATE2 = ["rocsvApplyCNOT", "rocsvApplyCZ", "rocsvApplySWAP"]
_CROT = ["rocsvApplyCRX", "rocsvApplyCRY", "rocsvApplyCRZ"]

PROTOTYPES = {
    "rocsvCreate": [C.POINTER(_h)],
    "rocsvDestroy": [_h],
    "rocsvAllocateState": [_h, _u, C.POINTER(_p), _sz],
    "rocsvFreeState": [_h],
    "rocsvInitializeState": [_h, _p, _u],
    "rocsvAllocateDistributedState": [_h, _u],
    "rocsvInitializeDistributedState": [_h],
    "rocsvApplyFusedSingleQubitMatrix": [_h, _u, _p],
    "rocsvSwapIndexBits": [_h, _u, _u],
    "rocsvApplyMatrix": [_h, _p, _u, _up, _u, _p, _u],
    "rocsvMeasure": [_h, _p, _u, _u, C.POINTER(C.c_int), C.POINTER(_d)],
    **{g: [_h, _p, _u, _u] for g in _GATE1},
    **{g: [_h, _p, _u, _u, _d] for g in _ROT1},
    **{g: [_h, _p, _u, _u, _u] for g in _GATE2},
    **{g: [_h, _p, _u, _u, _u, _d] for g in _CROT},
    "rocsvApplyMultiControlledX": [_h, _p, _u, _up, _u, _u],
    "rocsvApplyCSWAP": [_h, _p, _u, _u, _u, _u],
    "rocsvGetStateVectorFull": [_h, _p, _p],
    "rocsvGetStateVectorSlice": [_h, _p, _p, _u],
    "rocsvEnsurePinnedBuffer": [_h, _sz],
    "rocsvGetPinnedBufferPointer": [_h],
    "rocsvFreePinnedBuffer": [_h],
    "rocsvGetExpectationValueSinglePauliZ": [_h, _p, _u, _u, C.POINTER(_d)],
    "rocsvGetExpectationValueSinglePauliX": [_h, _p, _u, _u, C.POINTER(_d)],
    "rocsvGetExpectationValueSinglePauliY": [_h, _p, _u, _u, C.POINTER(_d)],
    "rocsvGetExpectationValuePauliProductZ": [_h, _p, _u, _up, _u, C.POINTER(_d)],
    "rocsvGetExpectationPauliString": [_h, _p, _u, C.c_char_p, _up, _u, C.POINTER(_d)],
    "rocsvSample": [_h, _p, _u, _up, _u, _u, C.POINTER(C.c_uint64)],
    "rocsvApplyControlledMatrix": [_h, _p, _u, _up, _u, _up, _u, _p],
    "rocsvApplyMatrixAndMeasure": [_h, _p, _u, _up, _u, _p, _u, C.POINTER(C.c_int)],
    # extensions
    "rocsvxGetPrecisionBytes": [],
    "rocsvxSetSeed": [_h, C.c_uint64],
    "rocsvxSetStateVector": [_h, _p, _p],
    "rocsvxSynchronize": [_h],
    "rocsvxSetFusion": [_h, C.c_int],
    "rocsvxFlush": [_h],
    "rocsvxApplyCircuit": [_h, _p, _u, C.POINTER(GateOp), _sz],
    "rocsvxSetPlanCache": [_h, C.c_int],
    "rocsvxApplyBlock6": [_h, _p, _u, _up, C.POINTER(_d)],
    "rocsvxSetTensorCoreBlocks": [_h, C.c_int],
    "rocsvxSetMergeDiagonals": [_h, C.c_int],
    "rocsvxGetNorm": [_h, _p, _u, C.POINTER(_d)],
    "rocsvxGetExpectationPauliBatch": [_h, _p, _u, C.c_char_p, _up, _up, _u, C.POINTER(_d)],
    "rocsvxGetExpectationPauliBatchAllStates": [_h, _p, _u, C.c_char_p, _up, _up, _u, C.POINTER(_d)],
    "rocsvxGetStats": [_h, C.POINTER(Stats), C.c_int],
    "rocsvxTimerStart": [_h],
    "rocsvxTimerStop": [_h, C.POINTER(_d)],
    "rocsvxPlanCircuit": [_u, _u, C.POINTER(GateOp), _sz, _up, C.c_char_p, _sz],
    "rocsvxPlanCircuitBlocks": [_u, C.POINTER(GateOp), _sz, _d, _up, _up, C.c_char_p, _sz],
    "rocsvxDistGetUniqueId": [_p],
    "rocsvxDistInit": [_h, C.c_int, C.c_int, _p],
    "rocsvxDistGetInfo": [_h, C.POINTER(C.c_int), C.POINTER(C.c_int), _up, C.POINTER(_p)],
    "rocsvxDistSetRanks": [_h, C.c_int],
    "rocsvxDistGetRankSlice": [_h, C.c_int, C.POINTER(C.c_int), C.POINTER(_p)],
    "rocsvxDistPlanCircuit": [_u, C.c_int, C.POINTER(GateOp), _sz, C.c_int, C.c_int, _up, C.c_char_p, _sz],
    "rocsvxDistPlanExchange": [_u, C.c_int, C.c_int, _up, _up, _u, C.POINTER(ExchangeSeg), _sz, C.POINTER(_sz)],
    "rocsvxDistPlanPeerSwap": [_u, C.c_int, C.c_int, _up, _up, _u, C.POINTER(ExchangeSeg), _sz, C.POINTER(_sz)],
}
SYMBOLS = sorted(PROTOTYPES)
REFERENCE_SYMBOLS = sorted(s for s in PROTOTYPES if not s.startswith("rocsvx"))   # the 42 of the reference header


def lib_path(prec: str = "c64") -> str:
    name = {"c64": "libhipStateVec.so", "c128": "libhipStateVec_f64.so"}[prec]
    return os.path.join(os.environ.get("ROCQ_LIB_DIR", os.path.join(_HERE, "lib")), name)


def load(prec: str = "c64") -> C.CDLL:
    """Load the engine.  Raises if the library has not been built: there is no fallback."""
    if prec not in _LIBS:
        path = lib_path(prec)
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python rocquantum_b200/build.py` (nvcc, sm_100a); "
                              "rocquantum_b200 has no CPU or PyTorch fallback")
        lib = C.CDLL(path)
        for name, args in PROTOTYPES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                if "ROCQ_LIB_DIR" in os.environ:      # an older tuning variant built elsewhere: it may lack newer extensions
                    continue
                raise
            fn.argtypes = args
            fn.restype = C.c_int
        lib.rocsvGetPinnedBufferPointer.restype = C.c_void_p
        lib.rocsvxGetPrecisionBytes.restype = C.c_uint
        _LIBS[prec] = lib
    return _LIBS[prec]


def uarr(xs):
    xs = list(xs)
    return (C.c_uint * max(1, len(xs)))(*xs)


def make_ops(gates):
    """gates: iterable of (name, targets, controls, theta[, matrix]) -> (GateOp array, keepalive list).

    name in h,x,y,z,s,sdg,t,rx,ry,rz,cnot,cz,swap,crx,cry,crz,mcx,cswap,matrix; `matrix` is a (2^k,2^k)
    array-like, M[i][j] row i / column j (converted to the ABI's column-major interleaved doubles)."""
    import numpy as np
    kinds = dict(h=ROCSVX_H, x=ROCSVX_X, y=ROCSVX_Y, z=ROCSVX_Z, s=ROCSVX_S, sdg=ROCSVX_SDG, t=ROCSVX_T, rx=ROCSVX_RX,
                 ry=ROCSVX_RY, rz=ROCSVX_RZ, cnot=ROCSVX_CNOT, cz=ROCSVX_CZ, swap=ROCSVX_SWAP, crx=ROCSVX_CRX,
                 cry=ROCSVX_CRY, crz=ROCSVX_CRZ, mcx=ROCSVX_MCX, cswap=ROCSVX_CSWAP, matrix=ROCSVX_MATRIX)
    gates = list(gates)
    arr = (GateOp * max(1, len(gates)))()
    keep = []
    for i, g in enumerate(gates):
        name, targets, controls, theta = g[0], list(g[1]), list(g[2]), float(g[3])
        op = arr[i]
        op.kind = kinds[name]
        op.numTargets = len(targets)
        for j, t in enumerate(targets):
            op.targets[j] = t
        m = 0
        for c in controls:
            m |= 1 << c
        op.controlMask = m
        op.theta = theta
        if name == "matrix":
            M = np.asarray(g[4], dtype=np.complex128).reshape(1 << len(targets), 1 << len(targets))
            flat = np.ascontiguousarray(M.T).reshape(-1)          # column-major
            buf = np.empty(2 * flat.size, dtype=np.float64)
            buf[0::2], buf[1::2] = flat.real, flat.imag
            keep.append(buf)
            op.matrix = buf.ctypes.data_as(C.POINTER(C.c_double))
    return arr, keep
