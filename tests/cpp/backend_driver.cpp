// tests/cpp/backend_driver.cpp -- drives rocq::HipStateVecBackend (the reference's C++ client of the C ABI,
// rocqCompiler/HipStateVecBackend.cpp:153-253) exactly as its MLIR executor would: create_backend("hip_statevec") ->
// initialize -> apply_gate / apply_parametrized_gate by NAME -> get_state_vector -> destroy.  Test infrastructure: the gate
// script comes from stdin, the state goes to stdout; tests/test_gpu_bindings.py compares it with the oracle.
//   n <qubits>            first line
//   g <name> <t0> <t1> ...
//   p <name> <theta> <t0> ...
//   e <name> <t0> ...     the call must throw (prints "threw <type>")
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "rocqCompiler/QuantumBackend.h"

int main() {
    std::unique_ptr<rocq::QuantumBackend> be = rocq::create_backend("hip_statevec");
    std::string line;
    unsigned n = 0;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        std::string op, name;
        in >> op;
        if (op == "n") { in >> n; be->initialize(n); continue; }
        in >> name;
        double theta = 0.0;
        if (op == "p") in >> theta;
        std::vector<unsigned> t;
        for (unsigned q; in >> q;) t.push_back(q);
        if (op == "g") be->apply_gate(name, t);
        else if (op == "p") be->apply_parametrized_gate(name, theta, t);
        else if (op == "e") {
            try { be->apply_gate(name, t); printf("no throw\n"); }
            catch (const std::invalid_argument&) { printf("threw invalid_argument\n"); }
            catch (const std::runtime_error&) { printf("threw runtime_error\n"); }
        }
    }
    const std::vector<std::complex<double>> sv = be->get_state_vector();
    printf("state %zu\n", sv.size());
    for (const auto& a : sv) printf("%.17g %.17g\n", a.real(), a.imag());
    be->destroy();
    try { be->get_state_vector(); printf("no throw\n"); } catch (const std::runtime_error&) { printf("threw runtime_error\n"); }
    return 0;
}
