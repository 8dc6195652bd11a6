// forwarding header: the interface lives in lgrngn_b200_api.hpp
#pragma once
#include "lgrngn_b200_api.hpp"
