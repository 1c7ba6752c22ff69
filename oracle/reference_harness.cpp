// reference_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Pulls in the reference's OWN translation unit, /root/reference/main.cpp, verbatim (its main() renamed away) and drives
// its own functions over a batch of rods: for every rod the global `qe` (main.cpp:17) is set and
//   integrateQuaternions()  main.cpp:91-118
//   integratePosition()     main.cpp:145-176
// are called exactly as main() does (main.cpp:197,201).  The strain samples the reference itself evaluates inside
// updateA (K = Phi<na,ne>(x[i]) * qe, main.cpp:69) are exported too, so that a checker can feed the oracle and the CUDA
// path the very same nodal K.  Nothing is restated here: every number comes out of the reference's code, compiled against
// oracle/eigen_shim (Eigen/Boost are not in the image; the shim's LU is a plain partial-pivot LU, see its header).
//
// Built by oracle/build_reference.py into oracle/_ref/libreference_harness.so; the include path below is given with -I.
#define main sri_reference_main_unused
#include "main.cpp"  // /root/reference/main.cpp
#undef main

extern "C" {

// qe_in [n][9] (component-major, mode-minor: qe[3c+k] as main.cpp:187-195) ->
//   K [n][3][16]  the reference's own strain samples at its nodes x[0..15] (tip -> base)
//   Q [n][60]     Q_stack of integrateQuaternions()
//   r [n][45]     r_stack of integratePosition(), column-major 15 x 3 = component-major, node-minor
int sri_reference_integrate_rods(long n, const double* qe_in, double* K, double* Q, double* r)
{
    constexpr int N = number_of_Chebyshev_points, M = N - 1;
    for (long b = 0; b < n; ++b) {
        for (int i = 0; i < 9; ++i) qe(i) = qe_in[b * 9 + i];
        for (int i = 0; i < N; ++i) {
            const Eigen::Vector3d Ki = Phi<na, ne>(x[i]) * qe;
            for (int c = 0; c < 3; ++c) K[(b * 3 + c) * N + i] = Ki(c);
        }
        const Eigen::VectorXd Q_stack = integrateQuaternions();
        for (int i = 0; i < 4 * M; ++i) Q[b * 4 * M + i] = Q_stack(i);
        const Eigen::MatrixXd r_stack = integratePosition();
        for (int c = 0; c < 3; ++c)
            for (int i = 0; i < M; ++i) r[(b * 3 + c) * M + i] = r_stack(i, c);
    }
    return N;
}

// the reference's updateA (main.cpp:55-88) on its own D_NN: A_NN [60][60] column-major for one qe
void sri_reference_update_A(const double* qe_in, double* A_colmajor)
{
    constexpr int N = number_of_Chebyshev_points, M = N - 1, n = 4 * M;
    const Eigen::MatrixXd Dn = getDn<number_of_Chebyshev_points>();
    const Eigen::MatrixXd Dn_NN = Dn.block<N - 1, N - 1>(0, 0);
    const MatrixNN D_NN = Eigen::KroneckerProduct(Eigen::MatrixXd::Identity(state_dimension, state_dimension), Dn_NN);
    MatrixNN A_NN = D_NN;
    Eigen::Matrix<double, ne * na, 1> q;
    for (int i = 0; i < 9; ++i) q(i) = qe_in[i];
    updateA(q, A_NN, D_NN);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) A_colmajor[j * n + i] = A_NN(i, j);
}

}  // extern "C"
