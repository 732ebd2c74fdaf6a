// Eigen integration of the tracing scalar of ../../cg.hpp: the traits the reference
// gets from CppAD's cppad_eigen.hpp / CppADCodeGen's cppadcg_eigen.hpp (include/ungar/autodiff/data_types.hpp:33-34).
#pragma once

#include <Eigen/Core>

#include "../../cg.hpp"

namespace Eigen {

template <>
struct NumTraits<CppAD::ADCGD> : GenericNumTraits<CppAD::ADCGD> {
    using Real       = CppAD::ADCGD;
    using NonInteger = CppAD::ADCGD;
    using Nested     = CppAD::ADCGD;
    using Literal    = CppAD::ADCGD;
    enum {
        IsComplex = 0, IsInteger = 0, IsSigned = 1, RequireInitialization = 1, ReadCost = 1, AddCost = 2, MulCost = 2
    };
    static Real epsilon() { return std::numeric_limits<double>::epsilon(); }
    static Real dummy_precision() { return 100.0 * std::numeric_limits<double>::epsilon(); }
    static Real highest() { return std::numeric_limits<double>::max(); }
    static Real lowest() { return std::numeric_limits<double>::lowest(); }
    static int digits10() { return std::numeric_limits<double>::digits10; }
};

// AD (x) double mixes, e.g. `-g0 * Vector3r::UnitZ()` (quadruped.example.cpp:168) and
// `Vector3r{0.1, 0.1, 10.0}.cwiseProduct(p - pRef)` (:228).
template <typename BinaryOp>
struct ScalarBinaryOpTraits<CppAD::ADCGD, double, BinaryOp> {
    using ReturnType = CppAD::ADCGD;
};
template <typename BinaryOp>
struct ScalarBinaryOpTraits<double, CppAD::ADCGD, BinaryOp> {
    using ReturnType = CppAD::ADCGD;
};

}  // namespace Eigen
