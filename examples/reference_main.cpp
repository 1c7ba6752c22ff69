// The reference's main() (main.cpp:181-205) on top of the drop-in layer: same qe, same two dumps.
//   g++ -std=c++17 -Iinclude examples/reference_main.cpp -L<pkg dir> -lsri_cuda -Wl,-rpath,<pkg dir> -o reference_main_gpu
#include <iostream>

#include "sri_reference_api.hpp"

int main() {
    const std::array<double, 9> qe = {0, 0, 0, 1.2877691307032, -1.63807499160786, 0.437406679142598, 0, 0, 0};
    const auto Q_stack = integrateQuaternions<16, 3>(qe);
    std::cout << "Q_stack : \n" << Q_stack << std::endl;
    const auto r_stack = integratePosition<16, 3>(qe);
    std::cout << "r_stack : \n" << r_stack << std::endl;
    return 0;
}
