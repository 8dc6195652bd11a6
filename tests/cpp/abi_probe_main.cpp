// prints the binary layout of the lgrngn API as seen through whatever headers are on the include path (tests/test_cpu_abi.py)
#include <iostream>
#include <lgrngn_abi_probe.hpp>
int main() { std::cout << lgrngn_abi_probe::describe_all(); return 0; }
