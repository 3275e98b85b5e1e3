// pybind11 module `rocquantum_bind` -- same Python surface as the reference's bindings.cpp:14-106:
// class QuantumSimulator (alias QSim) and class MLIRCompiler.  Linked against libhipStateVec_f64.so.
#define ROCQ_PRECISION_DOUBLE 1
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>

#include "rocqCompiler/QuantumBackend.h"
#include "rocquantum/QuantumSimulator.h"

namespace py = pybind11;
using rocquantum::QuantumSimulator;
typedef py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> cmat;

namespace {
// MLIR/LLVM are not part of the hot path: the class exists so that
// `rocquantum_bind.MLIRCompiler(n, "hip_statevec")` constructs (it creates the execution backend exactly like
// bindings.cpp:17-21), but compiling MLIR text is outside this engine.
struct MLIRCompilerStub {
    unsigned num_qubits;
    std::unique_ptr<rocq::QuantumBackend> backend;
    MLIRCompilerStub(unsigned n, const std::string& backend_name) : num_qubits(n), backend(rocq::create_backend(backend_name)) {}
};
std::vector<std::complex<double>> flat(const cmat& m, bool require2x2, const char* who) {
    if (m.ndim() != 2 || m.shape(0) != m.shape(1) || (require2x2 && m.shape(0) != 2))
        throw std::invalid_argument(std::string(who) + " expects a 2x2 complex matrix.");
    std::vector<std::complex<double>> host((size_t)m.size());
    std::memcpy(host.data(), m.data(), host.size() * sizeof(std::complex<double>));
    return host;
}
py::array_t<std::complex<double>> to_numpy(const std::vector<std::complex<double>>& v) {
    py::array_t<std::complex<double>> out(v.size());
    std::memcpy(out.mutable_data(), v.data(), v.size() * sizeof(std::complex<double>));
    return out;
}
}  // namespace

PYBIND11_MODULE(rocquantum_bind, m) {
    m.doc() = "rocQuantum simulator bindings on the B200-native state-vector engine";

    py::class_<MLIRCompilerStub>(m, "MLIRCompiler")
        .def(py::init<unsigned, const std::string&>())
        .def("compile_and_execute", [](MLIRCompilerStub&, const std::string&, py::dict) -> std::vector<std::complex<double>> {
            throw std::runtime_error("MLIRCompiler.compile_and_execute: the MLIR front end is outside the state-vector path");
        })
        .def("emit_qir", [](MLIRCompilerStub&, const std::string&) -> std::string {
            throw std::runtime_error("MLIRCompiler.emit_qir: the MLIR front end is outside the state-vector path");
        });

    py::class_<QuantumSimulator>(m, "QuantumSimulator")
        .def(py::init<unsigned>(), py::arg("num_qubits"))
        .def("reset", &QuantumSimulator::reset)
        .def("apply_gate", [](QuantumSimulator& s, const std::string& name, const std::vector<unsigned>& t, const std::vector<double>& p) { s.apply_gate(name, t, p); },
             py::arg("gate_name"), py::arg("targets"), py::arg("params") = std::vector<double>{})
        .def("apply_matrix", [](QuantumSimulator& s, cmat matrix, const std::vector<unsigned>& targets) {
                 s.apply_matrix(flat(matrix, targets.size() == 1, "apply_matrix"), targets);
             }, py::arg("matrix"), py::arg("targets"))
        .def("get_statevector", [](const QuantumSimulator& s) { return to_numpy(s.get_statevector()); })
        .def("measure", &QuantumSimulator::measure, py::arg("qubits"), py::arg("shots"))
        .def("num_qubits", &QuantumSimulator::num_qubits)
        .def("set_seed", &QuantumSimulator::set_seed, py::arg("seed"))
        .def("ApplyGate", [](QuantumSimulator& s, const std::string& name, int t) { s.ApplyGate(name, t); }, py::arg("gate_name"), py::arg("target_qubit"))
        .def("ApplyGate", [](QuantumSimulator& s, const std::string& name, int c, int t) { s.ApplyGate(name, c, t); }, py::arg("gate_name"),
             py::arg("control_qubit"), py::arg("target_qubit"))
        .def("ApplyGate", [](QuantumSimulator& s, cmat matrix, int t) { s.ApplyGate(flat(matrix, true, "ApplyGate"), t); }, py::arg("gate_matrix"),
             py::arg("target_qubit"))
        .def("Execute", &QuantumSimulator::Execute)
        .def("GetStateVector", [](const QuantumSimulator& s) { return to_numpy(s.GetStateVector()); });

    m.attr("QSim") = m.attr("QuantumSimulator");          // bindings.cpp:105
}
