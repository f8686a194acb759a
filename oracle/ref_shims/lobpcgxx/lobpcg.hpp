// TEST INFRASTRUCTURE (oracle/_ref build glue) -- not part of the product.
// The CI path only needs lobpcgxx::rayleigh_ritz (solvers/davidson.hpp:313);
// the full driver header pulls lapackpp routines we do not shim.
#pragma once
#include <lobpcgxx/rayleigh_ritz.hpp>
