// Trace-gas / aqueous species identifiers.  Aqueous chemistry itself is outside the scope of this
// back-end; the enumeration is kept because it appears in the particle-system method signatures
// (counterpart of reference include/libcloudph++/common/chem.hpp:9-21).
#pragma once
namespace libcloudphxx { namespace common { namespace chem {
  enum chem_species_t
  {
    HNO3, NH3, CO2, SO2, H2O2, O3, S_VI, H,
    chem_gas_n = O3 + 1, chem_rhs_beg = SO2, chem_rhs_fin = S_VI + 1, chem_all = H + 1
  };
}}}
