// pybind11 module `_rocq_hip_backend` -- the state-vector subset of the reference's python/rocq/bindings.cpp:142-494
// (+ GateOp / GateFusion, :685-697), same names and argument order, so python/rocq/api.py can drive this engine.
// Contract kept from the reference: gate functions RETURN the status enum; state/measure/sample/expectation functions
// THROW RuntimeError.  Deliberate fixes: allocate_state_internal returns a NON-owning buffer
// (the handle owns the state; the reference double-frees), rocsvAllocateState is called with its 4 arguments.
#include <cuda_runtime_api.h>
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <complex>

#include "rocquantum/GateFusion.h"
#include "rocquantum/hipStateVec.h"

namespace py = pybind11;
#ifdef ROCQ_PRECISION_DOUBLE
typedef std::complex<double> host_cplx;
#else
typedef std::complex<float> host_cplx;
#endif

class DeviceBuffer {
public:
    void* ptr_ = nullptr;
    size_t size_bytes_ = 0;
    bool owned_ = true;
    DeviceBuffer() = default;
    DeviceBuffer(size_t num_elements, size_t element_size) : size_bytes_(num_elements * element_size) {
        if (cudaMalloc(&ptr_, size_bytes_ ? size_bytes_ : 16) != cudaSuccess) throw std::runtime_error("Failed to allocate device memory in DeviceBuffer constructor");
    }
    DeviceBuffer(void* p, size_t bytes, bool own) : ptr_(p), size_bytes_(bytes), owned_(own) {}
    ~DeviceBuffer() { if (owned_ && ptr_) cudaFree(ptr_); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    DeviceBuffer(DeviceBuffer&& o) noexcept : ptr_(o.ptr_), size_bytes_(o.size_bytes_), owned_(o.owned_) { o.ptr_ = nullptr; o.size_bytes_ = 0; o.owned_ = false; }
    void copy_from_numpy(py::array_t<host_cplx, py::array::c_style | py::array::forcecast> a) {
        if (!ptr_ || (size_t)a.nbytes() > size_bytes_) throw std::runtime_error("Device buffer not allocated, null, or NumPy array too large.");
        if (cudaMemcpy(ptr_, a.data(), a.nbytes(), cudaMemcpyHostToDevice) != cudaSuccess) throw std::runtime_error("Failed to copy NumPy array to device");
    }
    rocComplex* c() const { return static_cast<rocComplex*>(ptr_); }
    size_t nbytes() const { return size_bytes_; }
};

class RocsvHandleWrapper {
public:
    rocsvHandle_t handle_ = nullptr;
    RocsvHandleWrapper() {
        const rocqStatus_t st = rocsvCreate(&handle_);
        if (st != ROCQ_STATUS_SUCCESS) throw std::runtime_error("Failed to create rocsvHandle: " + std::to_string((int)st));
    }
    ~RocsvHandleWrapper() { if (handle_) rocsvDestroy(handle_); }
    RocsvHandleWrapper(const RocsvHandleWrapper&) = delete;
    RocsvHandleWrapper& operator=(const RocsvHandleWrapper&) = delete;
    rocsvHandle_t get() const { return handle_; }
};

static void must(rocqStatus_t st, const char* what) {
    if (st != ROCQ_STATUS_SUCCESS) throw std::runtime_error(std::string(what) + " failed: " + std::to_string((int)st));
}
typedef const RocsvHandleWrapper& H;
typedef DeviceBuffer& D;

PYBIND11_MODULE(_rocq_hip_backend, m) {
    m.doc() = "rocQuantum hipStateVec bindings on the B200-native engine";

    py::enum_<rocqStatus_t>(m, "rocqStatus")
        .value("SUCCESS", ROCQ_STATUS_SUCCESS).value("FAILURE", ROCQ_STATUS_FAILURE).value("INVALID_VALUE", ROCQ_STATUS_INVALID_VALUE)
        .value("ALLOCATION_FAILED", ROCQ_STATUS_ALLOCATION_FAILED).value("HIP_ERROR", ROCQ_STATUS_HIP_ERROR)
        .value("NOT_IMPLEMENTED", ROCQ_STATUS_NOT_IMPLEMENTED).value("RCCL_ERROR", ROCQ_STATUS_RCCL_ERROR)
        .export_values();

    py::class_<DeviceBuffer>(m, "DeviceBuffer")
        .def(py::init<>())
        .def(py::init<size_t, size_t>(), py::arg("num_elements"), py::arg("element_size"))
        .def("copy_from_numpy", &DeviceBuffer::copy_from_numpy)
        .def("nbytes", &DeviceBuffer::nbytes);
    py::class_<RocsvHandleWrapper>(m, "RocsvHandle").def(py::init<>());

    m.def("allocate_state_internal", [](H h, unsigned n) {
        rocComplex* p = nullptr;
        must(rocsvAllocateState(h.get(), n, &p, 1), "rocsvAllocateState");
        return DeviceBuffer(p, ((size_t)1 << n) * sizeof(rocComplex), false);     // the handle owns the state
    }, py::arg("handle"), py::arg("num_qubits"));
    m.def("initialize_state", [](H h, D d, unsigned n) {
        if (d.nbytes() != ((size_t)1 << n) * sizeof(rocComplex)) throw std::runtime_error("DeviceBuffer size mismatch in initialize_state");
        return rocsvInitializeState(h.get(), d.c(), n);
    }, py::arg("handle"), py::arg("d_state_buffer"), py::arg("num_qubits"));
    m.def("allocate_distributed_state", [](H h, unsigned n) { must(rocsvAllocateDistributedState(h.get(), n), "rocsvAllocateDistributedState"); },
          py::arg("handle"), py::arg("total_num_qubits"));
    m.def("initialize_distributed_state", [](H h) { must(rocsvInitializeDistributedState(h.get()), "rocsvInitializeDistributedState"); }, py::arg("handle"));

#define RQ_BIND1(pyname, fn) m.def(pyname, [](H h, D d, unsigned n, unsigned t) { return fn(h.get(), d.c(), n, t); })
    RQ_BIND1("apply_x", rocsvApplyX); RQ_BIND1("apply_y", rocsvApplyY); RQ_BIND1("apply_z", rocsvApplyZ); RQ_BIND1("apply_h", rocsvApplyH);
    RQ_BIND1("apply_s", rocsvApplyS); RQ_BIND1("apply_t", rocsvApplyT); RQ_BIND1("apply_sdg", rocsvApplySdg);
#define RQ_BINDR(pyname, fn) m.def(pyname, [](H h, D d, unsigned n, unsigned t, double a) { return fn(h.get(), d.c(), n, t, a); })
    RQ_BINDR("apply_rx", rocsvApplyRx); RQ_BINDR("apply_ry", rocsvApplyRy); RQ_BINDR("apply_rz", rocsvApplyRz);
#define RQ_BIND2(pyname, fn) m.def(pyname, [](H h, D d, unsigned n, unsigned a, unsigned b) { return fn(h.get(), d.c(), n, a, b); })
    RQ_BIND2("apply_cnot", rocsvApplyCNOT); RQ_BIND2("apply_cz", rocsvApplyCZ); RQ_BIND2("apply_swap", rocsvApplySWAP);
#define RQ_BINDCR(pyname, fn) m.def(pyname, [](H h, D d, unsigned n, unsigned c, unsigned t, double a) { return fn(h.get(), d.c(), n, c, t, a); })
    RQ_BINDCR("apply_crx", rocsvApplyCRX); RQ_BINDCR("apply_cry", rocsvApplyCRY); RQ_BINDCR("apply_crz", rocsvApplyCRZ);
    m.def("apply_mcx", [](H h, D d, unsigned n, const std::vector<unsigned>& c, unsigned t) {
        return rocsvApplyMultiControlledX(h.get(), d.c(), n, c.data(), (unsigned)c.size(), t);
    });
    m.def("apply_cswap", [](H h, D d, unsigned n, unsigned c, unsigned a, unsigned b) { return rocsvApplyCSWAP(h.get(), d.c(), n, c, a, b); });

    m.def("apply_matrix", [](H h, D d, unsigned n, std::vector<unsigned> q, D mat, unsigned dim) {
        if (q.empty()) throw std::runtime_error("qubitIndices must not be empty for apply_matrix");
        return rocsvApplyMatrix(h.get(), d.c(), n, q.data(), (unsigned)q.size(), mat.c(), dim);
    }, py::arg("handle"), py::arg("d_state"), py::arg("num_qubits"), py::arg("qubit_indices"), py::arg("matrix_device"), py::arg("matrix_dim"));
    m.def("apply_controlled_matrix", [](H h, D d, unsigned n, const std::vector<unsigned>& c, const std::vector<unsigned>& t, D mat) {
        if (t.empty()) return ROCQ_STATUS_SUCCESS;
        if (c.empty()) return rocsvApplyMatrix(h.get(), d.c(), n, t.data(), (unsigned)t.size(), mat.c(), 1u << t.size());
        must(rocsvApplyControlledMatrix(h.get(), d.c(), n, c.data(), (unsigned)c.size(), t.data(), (unsigned)t.size(), mat.c()), "rocsvApplyControlledMatrix");
        return ROCQ_STATUS_SUCCESS;
    }, py::arg("handle"), py::arg("d_state"), py::arg("num_qubits"), py::arg("control_qubits"), py::arg("target_qubits"), py::arg("matrix_device"));

    m.def("measure", [](H h, D d, unsigned n, unsigned q) {
        int outcome = 0; double p = 0.0;
        must(rocsvMeasure(h.get(), d.c(), n, q, &outcome, &p), "rocsvMeasure");
        return py::make_tuple(outcome, p);
    }, py::arg("handle"), py::arg("d_state"), py::arg("num_qubits"), py::arg("qubit_to_measure"));
#define RQ_BINDE(pyname, fn) m.def(pyname, [](H h, D d, unsigned n, unsigned t) { double r = 0.0; must(fn(h.get(), d.c(), n, t, &r), #fn); return r; }, \
                                   py::arg("handle"), py::arg("d_state"), py::arg("num_qubits"), py::arg("target_qubit"))
    RQ_BINDE("get_expectation_value_z", rocsvGetExpectationValueSinglePauliZ);
    RQ_BINDE("get_expectation_value_x", rocsvGetExpectationValueSinglePauliX);
    RQ_BINDE("get_expectation_value_y", rocsvGetExpectationValueSinglePauliY);
    m.def("get_expectation_value_pauli_product_z", [](H h, D d, unsigned n, const std::vector<unsigned>& q) {
        if (q.empty()) return 1.0;
        double r = 0.0;
        must(rocsvGetExpectationValuePauliProductZ(h.get(), d.c(), n, q.data(), (unsigned)q.size(), &r), "rocsvGetExpectationValuePauliProductZ");
        return r;
    }, py::arg("handle"), py::arg("d_state"), py::arg("num_qubits"), py::arg("target_qubits"));
    m.def("get_expectation_pauli_string", [](H h, D d, unsigned n, const std::string& s, const std::vector<unsigned>& q) {
        if (s.size() != q.size()) throw std::runtime_error("Pauli string length must match the number of target qubits.");
        if (q.empty()) return 1.0;
        double r = 0.0;
        must(rocsvGetExpectationPauliString(h.get(), d.c(), n, s.c_str(), q.data(), (unsigned)q.size(), &r), "rocsvGetExpectationPauliString");
        return r;
    }, py::arg("handle"), py::arg("d_state"), py::arg("num_qubits"), py::arg("pauli_string"), py::arg("target_qubits"));
    // The whole Hamiltonian in one call (extension; what python/rocq/api.py:520-643 get_expval / grad and
    // rocquantum/solvers/vqe_solver.py:120-136 loop over term by term): terms = [(pauli string, [qubits]), ...] ->
    // ndarray of <psi|P_t|psi>, evaluated by one read sweep per distinct set of X/Y qubits and one device->host copy.
    // all_states=True evaluates every state of the handle's batch (parameter-shift batches): shape (batch, terms).
    m.def("get_expectation_pauli_batch", [](H h, D d, unsigned n, const std::vector<std::pair<std::string, std::vector<unsigned>>>& terms,
                                            bool all_states, size_t batch) {
        std::string paulis;
        std::vector<unsigned> qubits, offsets{0u};
        for (const auto& t : terms) {
            if (t.first.size() != t.second.size()) throw std::runtime_error("Pauli string length must match the number of target qubits.");
            paulis += t.first;
            qubits.insert(qubits.end(), t.second.begin(), t.second.end());
            offsets.push_back((unsigned)paulis.size());
        }
        const size_t states = all_states ? (batch ? batch : 1) : 1;
        py::array_t<double> out(states * terms.size());
        if (terms.empty()) return out;
        if (qubits.empty()) qubits.push_back(0u);
        must((all_states ? rocsvxGetExpectationPauliBatchAllStates : rocsvxGetExpectationPauliBatch)(
                 h.get(), d.c(), n, paulis.c_str(), qubits.data(), offsets.data(), (unsigned)terms.size(), out.mutable_data()),
             "rocsvxGetExpectationPauliBatch");
        if (all_states) out.resize({states, terms.size()});
        return out;
    }, py::arg("handle"), py::arg("d_state"), py::arg("num_qubits"), py::arg("terms"), py::arg("all_states") = false, py::arg("batch_size") = 1);
    m.def("sample", [](H h, D d, unsigned n, const std::vector<unsigned>& q, unsigned shots) {
        py::array_t<uint64_t> out(shots);
        if (shots == 0) return out;
        must(rocsvSample(h.get(), d.c(), n, q.data(), (unsigned)q.size(), shots, out.mutable_data()), "rocsvSample");
        return out;
    }, py::arg("handle"), py::arg("d_state"), py::arg("num_qubits"), py::arg("measured_qubits"), py::arg("num_shots"));
    m.def("get_state_vector_full", [](H h, D d, unsigned n, size_t batch) {
        py::array_t<host_cplx> out(batch * ((size_t)1 << n));
        must(rocsvGetStateVectorFull(h.get(), d.c(), reinterpret_cast<rocComplex*>(out.mutable_data())), "rocsvGetStateVectorFull");
        return out;
    }, py::arg("handle"), py::arg("d_state").noconvert(), py::arg("num_qubits"), py::arg("batch_size"));
    m.def("get_state_vector_slice", [](H h, D d, unsigned n, size_t, unsigned idx) {
        py::array_t<host_cplx> out((size_t)1 << n);
        must(rocsvGetStateVectorSlice(h.get(), d.c(), reinterpret_cast<rocComplex*>(out.mutable_data()), idx), "rocsvGetStateVectorSlice");
        return out;
    }, py::arg("handle"), py::arg("d_state").noconvert(), py::arg("num_qubits"), py::arg("batch_size"), py::arg("batch_index"));
    m.def("create_device_matrix_from_numpy", [](py::array_t<host_cplx, py::array::c_style | py::array::forcecast> a) {
        if (a.ndim() != 2) throw std::runtime_error("NumPy array must be 2D for matrix.");
        DeviceBuffer db((size_t)a.size(), sizeof(rocComplex));
        db.copy_from_numpy(a);
        return db;
    }, py::arg("numpy_array"));

    // extensions used by tests / a fused Circuit.flush()
    m.def("set_seed", [](H h, uint64_t s) { return rocsvxSetSeed(h.get(), s); });
    m.def("set_fusion", [](H h, bool on) { return rocsvxSetFusion(h.get(), on ? 1 : 0); });
    m.def("synchronize", [](H h) { return rocsvxSynchronize(h.get()); });

    py::class_<rocquantum::GateOp>(m, "GateOp")
        .def(py::init<>())
        .def_readwrite("name", &rocquantum::GateOp::name).def_readwrite("targets", &rocquantum::GateOp::targets)
        .def_readwrite("controls", &rocquantum::GateOp::controls).def_readwrite("params", &rocquantum::GateOp::params);
    py::class_<rocquantum::GateFusion>(m, "GateFusion")
        .def(py::init([](H h, D d, unsigned n) { return new rocquantum::GateFusion(h.get(), d.c(), n); }), py::arg("handle"), py::arg("d_state"),
             py::arg("num_qubits"), py::keep_alive<1, 2>(), py::keep_alive<1, 3>())
        .def("process_queue", &rocquantum::GateFusion::processQueue);
}
