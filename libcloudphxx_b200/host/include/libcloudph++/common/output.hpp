// Keys of the accumulated-precipitation ("puddle") map returned by diag_puddle()
// (counterpart of reference include/libcloudph++/common/output.hpp:9-57).
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include "chem.hpp"
namespace libcloudphxx { namespace common {
  enum output_t
  {
    outHNO3 = chem::HNO3, outNH3 = chem::NH3, outCO2 = chem::CO2, outSO2 = chem::SO2, outH2O2 = chem::H2O2,
    outO3 = chem::O3, outS_VI = chem::S_VI, outH = chem::H,
    outliq_vol, outdry_vol, outprtcl_num, outice_mass, outliq_num, outice_num
  };
  const std::map<output_t, std::string> output_names = {
    {outHNO3, "HNO3"}, {outNH3, "NH3"}, {outCO2, "CO2"}, {outSO2, "SO2"}, {outH2O2, "H2O2"}, {outO3, "O3"},
    {outS_VI, "S_VI"}, {outH, "H"}, {outliq_vol, "liquid_volume"}, {outdry_vol, "dry_volume"},
    {outprtcl_num, "particle_number"}, {outice_mass, "ice_mass"}, {outliq_num, "liquid_number"},
    {outice_num, "ice_number"}};
  inline output_t get_output_enum(const std::string &name)
  {
    for (const auto &kv : output_names) if (kv.second == name) return kv.first;
    throw std::runtime_error("Incorrect name for puddle: " + name);
  }
}}
