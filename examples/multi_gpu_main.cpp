// One C++ host process driving every GPU of the box through the C ABI (include/sri.h): sri_create_multi builds one operator
// set per device, sri_integrate_all_sharded splits a host batch into contiguous rod-index blocks (one host thread and one
// stream pipeline per device, no collective), sri_newton_static_shape_sharded runs the static shape solve the same way with
// the residual norms reduced between the device threads.  Checks: the sharded results are bit-identical to a one-device run,
// and the pure-tip-moment Newton solve returns the circular arc K = H^-1 M_tip.
//
//   g++ -std=c++17 -O2 -Iinclude examples/multi_gpu_main.cpp -L<package dir> -lsri_cuda -o examples/multi_gpu_main_gpu
//   examples/multi_gpu_main_gpu [ndev] [rods]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sri.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        if ((call) != SRI_OK) {                                                          \
            std::fprintf(stderr, "%s: %s\n", #call, sri_last_error_string());            \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char** argv) {
    const int N = 16, M = N - 1;
    int ndev = 0;
    CHECK(sri_device_count(&ndev));
    if (argc > 1) ndev = std::atoi(argv[1]);
    const long long B = argc > 2 ? std::atoll(argv[2]) : 200000;
    if (ndev < 1) { std::fprintf(stderr, "no CUDA device\n"); return 2; }

    // synthetic rods: constant + linear curvature, random tip wrench (a plain LCG; the benchmark's Philox stream lives in
    // sri_generate_rods, which fills device buffers)
    std::vector<double> x(N);
    CHECK(sri_chebyshev_points(N, 1.0, x.data()));
    unsigned long long state = 0x5EEDULL;
    auto uni = [&state]() { state = state * 6364136223846793005ULL + 1442695040888963407ULL; return double(state >> 11) / 9007199254740992.0; };
    std::vector<double> K(3 * N * B), F(3 * B), Mt(3 * B);
    for (long long b = 0; b < B; ++b) {
        for (int c = 0; c < 3; ++c) {
            const double alpha = 4 * uni() - 2, beta = 4 * uni() - 2;
            for (int i = 0; i < N; ++i) K[(b * 3 + c) * N + i] = alpha + beta * (2 * x[i] - 1);
            F[3 * b + c] = 2 * uni() - 1;
            Mt[3 * b + c] = 2 * uni() - 1;
        }
    }
    std::vector<double> Q(4 * M * B), r(3 * M * B), n(3 * M * B), m(3 * M * B), Q1(4 * M * B), m1(3 * M * B);
    std::vector<int> info(B, -1);

    sri_multi_handle mh = nullptr;
    CHECK(sri_create_multi(N, nullptr, ndev, &mh));
    sri_rod_batch rb;
    std::memset(&rb, 0, sizeof(rb));
    rb.batch = B; rb.K = K.data(); rb.F_tip = F.data(); rb.M_tip = Mt.data();
    rb.Q = Q.data(); rb.r = r.data(); rb.n = n.data(); rb.m = m.data(); rb.info = info.data();
    CHECK(sri_integrate_all_sharded(mh, &rb));  // warm-up: sizes the per-device staging buffers
    const auto t0 = std::chrono::steady_clock::now();
    CHECK(sri_integrate_all_sharded(mh, &rb));
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // the same batch on device 0 alone
    sri_handle h0 = nullptr;
    CHECK(sri_multi_get_handle(mh, 0, &h0));
    sri_rod_batch one = rb;
    one.Q = Q1.data(); one.m = m1.data(); one.r = nullptr; one.n = n.data(); one.info = nullptr;
    CHECK(sri_integrate_all(h0, &one));
    const bool same = std::memcmp(Q.data(), Q1.data(), sizeof(double) * Q.size()) == 0 &&
                      std::memcmp(m.data(), m1.data(), sizeof(double) * m.size()) == 0;
    std::printf("%d device(s), %lld rods from host memory: %.2f ms (%.3g rods/s end to end), sharded == single device: %s\n", ndev, B,
                dt * 1e3, B / dt, same ? "bit-identical" : "DIFFERENT");

    // static shape solve across the devices: pure tip moment about y -> K = H^-1 M_tip
    const int ne = 3, nq = 3 * ne;
    const long long Bn = 20000;
    const double H[3] = {1.0, 1.0, 0.77};
    std::vector<double> Fn(3 * Bn, 0.0), Mn(3 * Bn, 0.0), qe(nq * Bn, 0.0);
    for (long long b = 0; b < Bn; ++b) Mn[3 * b + 1] = 0.2 + 1.5 * double(b) / double(Bn);
    sri_newton_report rep;
    CHECK(sri_newton_static_shape_sharded(mh, Bn, ne, H, Fn.data(), Mn.data(), nullptr, qe.data(), 1e-11, 30, 0.0, &rep));
    double worst = 0.0;
    for (long long b = 0; b < Bn; ++b)
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < ne; ++k)
                worst = std::fmax(worst, std::fabs(qe[b * nq + c * ne + k] - (k == 0 ? Mn[3 * b + c] / H[c] : 0.0)));
    std::printf("sharded Newton: converged %d  iterations %d  rms %.3e  max |qe - H^-1 M| %.3e\n", rep.converged, rep.iterations, rep.rms, worst);
    CHECK(sri_destroy_multi(mh));
    return (same && rep.converged && worst < 1e-9) ? 0 : 1;
}
