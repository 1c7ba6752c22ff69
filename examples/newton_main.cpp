// Static shape of tip-loaded rods from C++ host code: the Newton loop runs inside libsri_cuda.so
// (sri_newton_static_shape, include/sri.h).  Known answer used as the check: a pure tip moment about a principal axis
// bends the rod into a circular arc, K = H^-1 M_tip (rod_modeling.pdf eq. 1.25).
//
//   g++ -std=c++17 -O2 -Iinclude examples/newton_main.cpp -L<package dir> -lsri_cuda -o examples/newton_main_gpu
#include <cmath>
#include <cstdio>
#include <vector>

#include "sri.h"

int main() {
    const int N = 16, ne = 3, n = 3 * ne;
    const long long B = 1000;
    const double H[3] = {1.0, 1.0, 0.77};
    sri_handle h = nullptr;
    if (sri_create(N, 0, &h) != SRI_OK) { std::fprintf(stderr, "sri_create: %s\n", sri_last_error_string()); return 2; }
    std::vector<double> F(3 * B, 0.0), Mt(3 * B, 0.0), qe(n * B, 0.0);
    for (long long b = 0; b < B; ++b) Mt[3 * b + 1] = 0.2 + 1.5 * double(b) / double(B);   // moment about y
    sri_newton_report rep;
    if (sri_newton_static_shape(h, B, ne, H, F.data(), Mt.data(), nullptr, qe.data(), 1e-11, 30, 1e-6, 0, nullptr, nullptr, &rep) != SRI_OK) {
        std::fprintf(stderr, "sri_newton_static_shape: %s\n", sri_last_error_string());
        return 2;
    }
    // K = Phi qe with Legendre modes: the constant mode of component c is qe[c*ne + 0]; the others must vanish
    double worst = 0.0;
    for (long long b = 0; b < B; ++b)
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < ne; ++k) {
                const double want = (k == 0) ? Mt[3 * b + c] / H[c] : 0.0;
                worst = std::fmax(worst, std::fabs(qe[b * n + c * ne + k] - want));
            }
    std::printf("converged %d  iterations %d  integrations %lld  rms %.3e  max |qe - H^-1 M| %.3e\n", rep.converged, rep.iterations,
                (long long)rep.integrations, rep.rms, worst);
    sri_destroy(h);
    return (rep.converged && worst < 1e-9) ? 0 : 1;
}
