// sri_host_math.hpp -- host-side construction of the strain-independent operator set (kernel "K0").
//
// Replaces ComputeChebyshevPoints / GetCoefficients_c / getDn (include/chebyshev_differentiation.h:19-108) and
// the per-call Dn_NN.inverse() of main.cpp:159.  Dn itself is built in FP64 with the reference's operation
// order (so the differentiation matrix is bit-identical to the one Eigen code would produce); its cached
// inverses are computed once in extended precision and rounded to FP64.
#pragma once
#include <cmath>
#include <vector>

namespace sri_host {

inline void chebyshev_points(int N, double L, double* x) {
    // x_j = L/2 (1 + cos(pi j/(N-1))), descending: x_0 = L (tip), x_{N-1} = 0 (base)
    for (int j = 0; j < N; ++j) x[j] = (L / 2) * (1 + std::cos(M_PI * static_cast<double>(j) / static_cast<double>(N - 1)));
}

inline void chebyshev_coefficients(int N, double* c) {
    for (int i = 0; i < N; ++i) c[i] = std::pow(-1, i) * ((i == 0 || i == N - 1) ? 2u : 1u);
}

// column-major N x N
inline void chebyshev_dn(int N, double* Dn) {
    std::vector<double> x(N), c(N);
    chebyshev_points(N, 1.0, x.data());
    chebyshev_coefficients(N, c.data());
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) Dn[j * N + i] = (c[i] / c[j]) / (x[i] - x[j] + (i == j ? 1.0 : 0.0));
    for (int i = 0; i < N; ++i) {
        double rowsum = 0.0;
        for (int j = 0; j < N; ++j) rowsum += Dn[j * N + i];
        Dn[i * N + i] -= rowsum;
    }
}

inline double legendre(int l, double x) {
    if (l == 0) return 1.0;
    double pm = 1.0, p = x;
    for (int k = 1; k < l; ++k) {
        const double pn = ((2 * k + 1) * x * p - k * pm) / (k + 1);
        pm = p;
        p = pn;
    }
    return p;
}

// Inverse of a column-major n x n FP64 matrix by Gauss-Jordan with partial pivoting in long double.
inline bool invert_extended(int n, const double* A, double* Ainv) {
    std::vector<long double> W(static_cast<size_t>(n) * 2 * n, 0.0L);  // row-major [A | I]
    const int ld = 2 * n;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) W[i * ld + j] = A[j * n + i];
        W[i * ld + n + i] = 1.0L;
    }
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i)
            if (fabsl(W[i * ld + k]) > fabsl(W[p * ld + k])) p = i;
        if (W[p * ld + k] == 0.0L) return false;
        if (p != k)
            for (int j = 0; j < ld; ++j) std::swap(W[k * ld + j], W[p * ld + j]);
        const long double inv = 1.0L / W[k * ld + k];
        for (int j = 0; j < ld; ++j) W[k * ld + j] *= inv;
        for (int i = 0; i < n; ++i) {
            if (i == k) continue;
            const long double f = W[i * ld + k];
            if (f == 0.0L) continue;
            for (int j = 0; j < ld; ++j) W[i * ld + j] -= f * W[k * ld + j];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Ainv[j * n + i] = static_cast<double>(W[i * ld + n + j]);
    return true;
}

// Clenshaw-Curtis weights of the N Chebyshev-Gauss-Lobatto nodes x_j = (1 + cos(pi j/(N-1)))/2 on [0,1]
// (exact for polynomials of degree <= N-1).
inline void clenshaw_curtis_weights(int N, double* w) {
    const int n = N - 1;
    for (int j = 0; j <= n; ++j) {
        long double s = 0.0L;
        for (int k = 0; k <= n / 2; ++k) {
            const long double bk = (k == 0 || 2 * k == n) ? 1.0L : 2.0L;
            s += bk / (1.0L - 4.0L * k * k) * cosl(2.0L * k * j * M_PIl / n);
        }
        const long double cj = (j == 0 || j == n) ? 1.0L : 2.0L;
        w[j] = static_cast<double>(0.5L * cj / n * s);  // 0.5: interval [0,1] instead of [-1,1]
    }
}

struct OperatorSet {
    int N = 0, M = 0;
    std::vector<double> x;      // nodes
    std::vector<double> Dn;     // N x N
    std::vector<double> Dn_NN;  // M x M   Dn[0:M,0:M]   (main.cpp:94)
    std::vector<double> Dn_IN;  // M       Dn[0:M,M]     (main.cpp:95)
    std::vector<double> S;      // M x M   Dn_NN^-1      (main.cpp:159)
    std::vector<double> g;      // M       -S Dn_IN
    std::vector<double> D_TT;   // M x M   Dn[1:N,1:N]
    std::vector<double> D_TI;   // M       Dn[1:N,0]
    std::vector<double> ST;     // M x M   D_TT^-1
    std::vector<double> gT;     // M       -ST D_TI

    bool build(int n_nodes) {
        N = n_nodes;
        M = N - 1;
        x.resize(N);
        chebyshev_points(N, 1.0, x.data());
        Dn.resize(static_cast<size_t>(N) * N);
        chebyshev_dn(N, Dn.data());
        Dn_NN.resize(static_cast<size_t>(M) * M);
        D_TT.resize(static_cast<size_t>(M) * M);
        Dn_IN.resize(M);
        D_TI.resize(M);
        for (int j = 0; j < M; ++j)
            for (int i = 0; i < M; ++i) {
                Dn_NN[j * M + i] = Dn[j * N + i];
                D_TT[j * M + i] = Dn[(j + 1) * N + (i + 1)];
            }
        for (int i = 0; i < M; ++i) {
            Dn_IN[i] = Dn[static_cast<size_t>(M) * N + i];
            D_TI[i] = Dn[i + 1];
        }
        S.resize(static_cast<size_t>(M) * M);
        ST.resize(static_cast<size_t>(M) * M);
        if (!invert_extended(M, Dn_NN.data(), S.data())) return false;
        if (!invert_extended(M, D_TT.data(), ST.data())) return false;
        g.resize(M);
        gT.resize(M);
        for (int i = 0; i < M; ++i) {
            long double a = 0.0L, b = 0.0L;
            for (int j = 0; j < M; ++j) {
                a -= static_cast<long double>(S[j * M + i]) * Dn_IN[j];
                b -= static_cast<long double>(ST[j * M + i]) * D_TI[j];
            }
            g[i] = static_cast<double>(a);
            gT[i] = static_cast<double>(b);
        }
        return true;
    }
};

}  // namespace sri_host
