/*
 * sri_oracle.c -- CPU oracle for the spectral rod integration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the reference algorithm
 * (aGotelli/experimental_gpu_programming_for_a_spectral_numerical_integration).  It may be used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker or the
 * timed CPU baseline.  Nothing in the product package links, imports or executes it.
 *
 * Parity pin status: the reference ships no tests, golden vectors or expected output, and its third-party
 * dependencies (Eigen >= 3.4 per CMakeLists.txt:11, Boost.Math unpinned per include/utilities.h:13-14) are not
 * in this image.  The pin is therefore (see DESIGN.md "Oracle"):
 *   1. oracle/_ref/reference_main: the reference's own main.cpp + headers compiled VERBATIM from
 *      /root/reference against oracle/eigen_shim (a minimal stand-in for the Eigen/Boost API subset it uses);
 *      its full-precision output is committed under tests/golden/ and this oracle must reproduce it;
 *   2. known-answer tests (analytic straight rod / circular arc / dead-load couple), numpy/LAPACK and
 *      50-digit mpmath restatements in tests/.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Conventions (main.cpp:80-81,130-133): per-rod stacks are component-major, node-minor; node 0 is the rod tip
 * (X=1), node N-1 the base (X=0) (chebyshev_differentiation.h:26).  All matrices column-major like Eigen.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(i, j, ld) ((size_t)(j) * (size_t)(ld) + (size_t)(i))

/* include/chebyshev_differentiation.h:19-30 -- ComputeChebyshevPoints<N,L>() */
void sri_oracle_chebyshev_points(int N, double L, double* x)
{
    for (int j = 0; j < N; ++j)
        x[j] = (L / 2) * (1 + cos(M_PI * (double)j / (double)(N - 1)));
}

/* include/chebyshev_differentiation.h:37-52 -- GetCoefficients_c<N>() */
void sri_oracle_coefficients_c(int N, double* c)
{
    for (int i = 0; i < N; ++i) {
        const unsigned gain = (i == 0 || i == N - 1) ? 2 : 1;
        c[i] = pow(-1, i) * gain;
    }
}

/* include/chebyshev_differentiation.h:59-108 -- getDn<N>().  Same operation order as the reference:
 * X(i,:) = x_i (:70-71); C = c_i/c_j (:82-86); dX = X - X^T + I (:89); Dn = C/dX (:96-100);
 * diag -= rowwise sum (:104, the sum includes the unit diagonal placeholder). */
void sri_oracle_dn(int N, double* Dn)
{
    double* x = (double*)malloc(sizeof(double) * N);
    double* c = (double*)malloc(sizeof(double) * N);
    sri_oracle_chebyshev_points(N, 1.0, x);
    sri_oracle_coefficients_c(N, c);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            const double C = c[i] / c[j];
            const double dX = x[i] - x[j] + (i == j ? 1.0 : 0.0);
            Dn[IDX(i, j, N)] = C / dX;
        }
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int j = 0; j < N; ++j) s += Dn[IDX(i, j, N)];
        Dn[IDX(i, i, N)] -= s;
    }
    free(x);
    free(c);
}

/* boost::math::legendre_p(l, x) as called at include/utilities.h:59 -- three-term recurrence
 * P_{k+1} = ((2k+1) x P_k - k P_{k-1}) / (k+1) (Boost.Math legendre_next). */
double sri_oracle_legendre_p(int l, double x)
{
    if (l < 0) l = -l - 1;
    double p0 = 1.0, p1 = x;
    if (l == 0) return p0;
    for (int k = 1; k < l; ++k) {
        const double p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1);
        p0 = p1;
        p1 = p2;
    }
    return p1;
}

/* include/utilities.h:49-67 -- Phi<na,ne>(X, begin, end) = I_na (x) [P_0..P_{ne-1}](x)^T, na x (na*ne), col-major. */
void sri_oracle_phi(int na, int ne, double X, double begin, double end, double* out)
{
    const double x = (2 * X - (end + begin)) / (end - begin);
    memset(out, 0, sizeof(double) * na * na * ne);
    for (int a = 0; a < na; ++a)
        for (int k = 0; k < ne; ++k)
            out[IDX(a, a * ne + k, na)] = sri_oracle_legendre_p(k, x);
}

/* main.cpp:69 -- K = Phi<na,ne>(x[i]) * qe at every Chebyshev node; output K[c*N + i], c in 0..na-1. */
void sri_oracle_strain_from_modes(int N, int na, int ne, const double* qe, double* K)
{
    double* x = (double*)malloc(sizeof(double) * N);
    double* phi = (double*)malloc(sizeof(double) * na * na * ne);
    sri_oracle_chebyshev_points(N, 1.0, x);
    for (int i = 0; i < N; ++i) {
        sri_oracle_phi(na, ne, x[i], 0.0, 1.0, phi);
        for (int a = 0; a < na; ++a) {
            double s = 0.0;
            for (int k = 0; k < na * ne; ++k) s += phi[IDX(a, k, na)] * qe[k];
            K[a * N + i] = s;
        }
    }
    free(x);
    free(phi);
}

/* ---- dense helpers standing in for Eigen (not in the image) ------------------------------------------- */

/* Partial-pivot LU, in place, column-major n x n (what Eigen::PartialPivLU does; unblocked).
 * Returns 0, or k+1 if the k-th pivot is exactly zero. */
static int lu_factor(int n, double* A, int* piv)
{
    int info = 0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(A[IDX(k, k, n)]);
        for (int i = k + 1; i < n; ++i) {
            const double v = fabs(A[IDX(i, k, n)]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (best == 0.0) { if (!info) info = k + 1; continue; }
        if (p != k)
            for (int j = 0; j < n; ++j) {
                const double t = A[IDX(k, j, n)];
                A[IDX(k, j, n)] = A[IDX(p, j, n)];
                A[IDX(p, j, n)] = t;
            }
        const double inv = 1.0 / A[IDX(k, k, n)];
        for (int i = k + 1; i < n; ++i) A[IDX(i, k, n)] *= inv;
        for (int j = k + 1; j < n; ++j) {
            const double u = A[IDX(k, j, n)];
            if (u != 0.0)
                for (int i = k + 1; i < n; ++i) A[IDX(i, j, n)] -= A[IDX(i, k, n)] * u;
        }
    }
    return info;
}

/* Solve with the factors for nrhs right-hand sides stored column-major in B (n x nrhs). */
static void lu_solve(int n, const double* LU, const int* piv, double* B, int nrhs)
{
    for (int r = 0; r < nrhs; ++r) {
        double* b = B + (size_t)r * n;
        for (int k = 0; k < n; ++k) {
            const int p = piv[k];
            if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
        }
        for (int k = 0; k < n; ++k) {
            const double v = b[k];
            if (v != 0.0)
                for (int i = k + 1; i < n; ++i) b[i] -= LU[IDX(i, k, n)] * v;
        }
        for (int k = n - 1; k >= 0; --k) {
            b[k] /= LU[IDX(k, k, n)];
            const double v = b[k];
            for (int i = 0; i < k; ++i) b[i] -= LU[IDX(i, k, n)] * v;
        }
    }
}

/* MatrixBase::inverse() for sizes > 4 == PartialPivLU(A).solve(Identity) (main.cpp:113,159). A is destroyed. */
static int dense_inverse(int n, double* A, double* Ainv, int* piv)
{
    const int info = lu_factor(n, A, piv);
    memset(Ainv, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; ++i) Ainv[IDX(i, i, n)] = 1.0;
    lu_solve(n, A, piv, Ainv, n);
    return info;
}

/* Quaterniond(w,x,y,z).toRotationMatrix() (Eigen Geometry, no normalisation), called at main.cpp:136. Row-major R[9]. */
static void quat_to_rot(double w, double x, double y, double z, double* R)
{
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* ---- operator set, once per N -------------------------------------------------------------------------- */

typedef struct {
    int N, M;
    double* Dn;        /* N x N */
    double* Dn_NN;     /* M x M  = Dn[0:M,0:M]   main.cpp:94  */
    double* Dn_IN;     /* M      = Dn[0:M,M]     main.cpp:95  */
    double* Dn_NN_inv; /* M x M                  main.cpp:159 */
    double* D_TT;      /* M x M  = Dn[1:N,1:N]   tip-BC twin */
    double* D_TI;      /* M      = Dn[1:N,0] */
    double* D_TT_inv;  /* M x M */
} sri_oracle_ops;

sri_oracle_ops* sri_oracle_ops_create(int N)
{
    sri_oracle_ops* o = (sri_oracle_ops*)calloc(1, sizeof(*o));
    const int M = N - 1;
    o->N = N; o->M = M;
    o->Dn = (double*)malloc(sizeof(double) * N * N);
    o->Dn_NN = (double*)malloc(sizeof(double) * M * M);
    o->Dn_IN = (double*)malloc(sizeof(double) * M);
    o->Dn_NN_inv = (double*)malloc(sizeof(double) * M * M);
    o->D_TT = (double*)malloc(sizeof(double) * M * M);
    o->D_TI = (double*)malloc(sizeof(double) * M);
    o->D_TT_inv = (double*)malloc(sizeof(double) * M * M);
    sri_oracle_dn(N, o->Dn);
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i) {
            o->Dn_NN[IDX(i, j, M)] = o->Dn[IDX(i, j, N)];
            o->D_TT[IDX(i, j, M)] = o->Dn[IDX(i + 1, j + 1, N)];
        }
    for (int i = 0; i < M; ++i) {
        o->Dn_IN[i] = o->Dn[IDX(i, M, N)];
        o->D_TI[i] = o->Dn[IDX(i + 1, 0, N)];
    }
    double* tmp = (double*)malloc(sizeof(double) * M * M);
    int* piv = (int*)malloc(sizeof(int) * M);
    memcpy(tmp, o->Dn_NN, sizeof(double) * M * M);
    dense_inverse(M, tmp, o->Dn_NN_inv, piv);
    memcpy(tmp, o->D_TT, sizeof(double) * M * M);
    dense_inverse(M, tmp, o->D_TT_inv, piv);
    free(tmp);
    free(piv);
    return o;
}

void sri_oracle_ops_destroy(sri_oracle_ops* o)
{
    if (!o) return;
    free(o->Dn); free(o->Dn_NN); free(o->Dn_IN); free(o->Dn_NN_inv);
    free(o->D_TT); free(o->D_TI); free(o->D_TT_inv);
    free(o);
}

const double* sri_oracle_ops_get(const sri_oracle_ops* o, int which)
{
    switch (which) {
        case 0: return o->Dn;
        case 1: return o->Dn_NN;
        case 2: return o->Dn_IN;
        case 3: return o->Dn_NN_inv;
        case 4: return o->D_TT;
        case 5: return o->D_TI;
        case 6: return o->D_TT_inv;
    }
    return 0;
}

/* ---- stage 1: quaternions ------------------------------------------------------------------------------ */

/* main.cpp:55-88 -- updateA: A_NN starts as D_NN = I4 (x) Dn_NN (main.cpp:98,103) and only the 16*M entries
 * (r*M+i, c*M+i) are overwritten with D_NN - 0.5*A(K_i) (main.cpp:72-82).  K is [3][N] nodal samples (the value
 * Phi(x_i)*qe of main.cpp:69); the base node i=N-1 is not used (loop bound main.cpp:66). */
void sri_oracle_assemble_A(const sri_oracle_ops* o, const double* K, double* A_NN)
{
    const int N = o->N, M = o->M, n = 4 * M;
    memset(A_NN, 0, sizeof(double) * (size_t)n * n);
    for (int blk = 0; blk < 4; ++blk)
        for (int j = 0; j < M; ++j)
            for (int i = 0; i < M; ++i) A_NN[IDX(blk * M + i, blk * M + j, n)] = o->Dn_NN[IDX(i, j, M)];
    for (int i = 0; i < M; ++i) {
        const double k0 = K[0 * N + i], k1 = K[1 * N + i], k2 = K[2 * N + i];
        const double A[4][4] = {{0, -k0, -k1, -k2}, {k0, 0, k2, -k1}, {k1, -k2, 0, k0}, {k2, k1, -k0, 0}};
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                const double d = (r == c) ? o->Dn_NN[IDX(i, i, M)] : 0.0;
                A_NN[IDX(r * M + i, c * M + i, n)] = d - 0.5 * A[r][c];
            }
    }
}

/* main.cpp:91-118 -- integrateQuaternions: ivp = D_IN*q_init (:100,:109), b = 0 (:111),
 * Q_stack = A_NN.inverse() * (b - ivp) (:113).  explicit_inverse=1 reproduces the inverse-then-multiply literally;
 * explicit_inverse=0 is the LU-solve variant reported beside it.  q0 NULL => (1,0,0,0) (:106-107).
 * work: >= 2*n*n + n doubles, iwork: >= n ints.  Returns LU info. */
int sri_oracle_integrate_quaternions(const sri_oracle_ops* o, const double* K, const double* q0, double* Q,
                                     int explicit_inverse, double* work, int* iwork)
{
    const int M = o->M, n = 4 * M;
    static const double q_default[4] = {1.0, 0.0, 0.0, 0.0};
    if (!q0) q0 = q_default;
    double* A = work;
    double* Ainv = work + (size_t)n * n;
    double* rhs = work + 2 * (size_t)n * n;
    sri_oracle_assemble_A(o, K, A);
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < M; ++i) rhs[c * M + i] = 0.0 - o->Dn_IN[i] * q0[c];
    int info;
    if (explicit_inverse) {
        info = dense_inverse(n, A, Ainv, iwork);
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s += Ainv[IDX(i, j, n)] * rhs[j];
            Q[i] = s;
        }
    } else {
        info = lu_factor(n, A, iwork);
        memcpy(Q, rhs, sizeof(double) * n);
        lu_solve(n, A, iwork, Q, 1);
    }
    return info;
}

/* ---- stage 2: positions -------------------------------------------------------------------------------- */

/* main.cpp:121-140 -- updatePositionb: b.row(i) = (R(q_i) * Gamma_i)^T with q = {Q[i],Q[i+M],Q[i+2M],Q[i+3M]}
 * (w,x,y,z; :130-133).  The reference hard-codes Gamma=(1,0,0) (:136); Gamma==NULL reproduces that, otherwise
 * Gamma is [3][N] nodal samples (SURVEY 8f3 generalisation).  b is M x 3 column-major. */
void sri_oracle_update_position_b(const sri_oracle_ops* o, const double* Q, const double* Gamma, double* b)
{
    const int N = o->N, M = o->M;
    double R[9];
    for (int i = 0; i < M; ++i) {
        quat_to_rot(Q[i], Q[i + M], Q[i + 2 * M], Q[i + 3 * M], R);
        const double g0 = Gamma ? Gamma[0 * N + i] : 1.0;
        const double g1 = Gamma ? Gamma[1 * N + i] : 0.0;
        const double g2 = Gamma ? Gamma[2 * N + i] : 0.0;
        for (int c = 0; c < 3; ++c) b[c * M + i] = R[3 * c + 0] * g0 + R[3 * c + 1] * g1 + R[3 * c + 2] * g2;
    }
}

/* main.cpp:145-176 -- integratePosition: ivp.row(i) = Dn_IN(i)*r_init^T (:162-164),
 * r_stack = Dn_NN_inv * (b_NN - ivp) (:172).  r0 NULL => 0 (:151-154).  r is M x 3 column-major. */
void sri_oracle_integrate_position(const sri_oracle_ops* o, const double* Q, const double* Gamma, const double* r0,
                                   double* r)
{
    const int M = o->M;
    double* b = (double*)malloc(sizeof(double) * 3 * M);
    sri_oracle_update_position_b(o, Q, Gamma, b);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < M; ++i) b[c * M + i] -= o->Dn_IN[i] * (r0 ? r0[c] : 0.0);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < M; ++i) {
            double s = 0.0;
            for (int j = 0; j < M; ++j) s += o->Dn_NN_inv[IDX(i, j, M)] * b[c * M + j];
            r[c * M + i] = s;
        }
    free(b);
}

/* ---- stages 3-4: not implemented in the reference; spec = materials/rod_modeling.pdf eqs. 1.17-1.18 with the
 * BC-elimination pattern of main.cpp:94-113 mirrored to the tip node.  The dead helper
 * skew() (include/utilities.h:16-24) defines the cross-product convention. ---------------------------------- */

/* n' = -fbar, n(X=1) = F_tip  =>  n_stack(nodes 1..N-1) = D_TT^{-1} (-fbar[1:N] - D_TI F_tip^T).
 * fbar is [3][N] nodal samples or NULL (=0).  n is M x 3 column-major. */
void sri_oracle_integrate_stress(const sri_oracle_ops* o, const double* fbar, const double* F_tip, double* n)
{
    const int N = o->N, M = o->M;
    double* b = (double*)malloc(sizeof(double) * 3 * M);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < M; ++i) b[c * M + i] = -(fbar ? fbar[c * N + i + 1] : 0.0) - o->D_TI[i] * F_tip[c];
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < M; ++i) {
            double s = 0.0;
            for (int j = 0; j < M; ++j) s += o->D_TT_inv[IDX(i, j, M)] * b[c * M + j];
            n[c * M + i] = s;
        }
    free(b);
}

/* m' = -(r' x n + lbar), m(X=1) = M_tip, r' = R(q) Gamma.  Node 0 uses n_0 = F_tip; node N-1 (base) uses q0.
 * m_stack(nodes 1..N-1) = D_TT^{-1} (-(r' x n + lbar)[1:N] - D_TI M_tip^T). */
void sri_oracle_integrate_couple(const sri_oracle_ops* o, const double* Q, const double* q0, const double* Gamma,
                                 const double* n, const double* lbar, const double* M_tip, double* m)
{
    const int N = o->N, M = o->M;
    static const double q_default[4] = {1.0, 0.0, 0.0, 0.0};
    if (!q0) q0 = q_default;
    double* b = (double*)malloc(sizeof(double) * 3 * M);
    double R[9];
    for (int i = 1; i < N; ++i) { /* global node index; row i-1 of the reduced system */
        if (i < M) quat_to_rot(Q[i], Q[i + M], Q[i + 2 * M], Q[i + 3 * M], R);
        else quat_to_rot(q0[0], q0[1], q0[2], q0[3], R);
        const double g0 = Gamma ? Gamma[0 * N + i] : 1.0;
        const double g1 = Gamma ? Gamma[1 * N + i] : 0.0;
        const double g2 = Gamma ? Gamma[2 * N + i] : 0.0;
        double rp[3], nn[3];
        for (int c = 0; c < 3; ++c) {
            rp[c] = R[3 * c + 0] * g0 + R[3 * c + 1] * g1 + R[3 * c + 2] * g2;
            nn[c] = n[c * M + (i - 1)];
        }
        const double cr[3] = {rp[1] * nn[2] - rp[2] * nn[1], rp[2] * nn[0] - rp[0] * nn[2],
                              rp[0] * nn[1] - rp[1] * nn[0]};
        for (int c = 0; c < 3; ++c)
            b[c * M + (i - 1)] = -(cr[c] + (lbar ? lbar[c * N + i] : 0.0)) - o->D_TI[i - 1] * M_tip[c];
    }
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < M; ++i) {
            double s = 0.0;
            for (int j = 0; j < M; ++j) s += o->D_TT_inv[IDX(i, j, M)] * b[c * M + j];
            m[c * M + i] = s;
        }
    free(b);
}

/* Newton residual of the static shape problem (PDF eq. 1.25):
 * rho_i = H (K_i - K0_i) - R(q_i)^T m_i, i = 0..N-1, with m_0 = M_tip and q_{N-1} = q0.  rho is [3][N]. */
void sri_oracle_shape_residual(const sri_oracle_ops* o, const double* K, const double* K0, const double* H_diag,
                               const double* Q, const double* q0, const double* m, const double* M_tip, double* rho)
{
    const int N = o->N, M = o->M;
    static const double q_default[4] = {1.0, 0.0, 0.0, 0.0};
    if (!q0) q0 = q_default;
    double R[9];
    for (int i = 0; i < N; ++i) {
        if (i < M) quat_to_rot(Q[i], Q[i + M], Q[i + 2 * M], Q[i + 3 * M], R);
        else quat_to_rot(q0[0], q0[1], q0[2], q0[3], R);
        double mi[3];
        for (int c = 0; c < 3; ++c) mi[c] = (i == 0) ? M_tip[c] : m[c * M + (i - 1)];
        for (int c = 0; c < 3; ++c) {
            const double Rtm = R[0 * 3 + c] * mi[0] + R[1 * 3 + c] * mi[1] + R[2 * 3 + c] * mi[2];
            rho[c * N + i] = H_diag[c] * (K[c * N + i] - (K0 ? K0[c * N + i] : 0.0)) - Rtm;
        }
    }
}

/* Local-frame wrench Lambda = [C; N] = [R^T m; R^T n] at all N nodes (rod_modeling.pdf eqs. 1.29, 2.18: the quantity
 * that obeys Lambda' = ad^T_xi Lambda - Fbar; couple first, matching the [k; gamma] ordering of ad(), utilities.h:27-37).
 * Node 0 carries the tip wrench (M_tip, F_tip), node N-1 the base rotation q0.  Lambda is [6][N]. */
void sri_oracle_wrench_local(const sri_oracle_ops* o, const double* Q, const double* q0, const double* n, const double* m,
                             const double* F_tip, const double* M_tip, double* Lambda)
{
    const int N = o->N, M = o->M;
    static const double q_default[4] = {1.0, 0.0, 0.0, 0.0};
    if (!q0) q0 = q_default;
    double R[9];
    for (int i = 0; i < N; ++i) {
        if (i < M) quat_to_rot(Q[i], Q[i + M], Q[i + 2 * M], Q[i + 3 * M], R);
        else quat_to_rot(q0[0], q0[1], q0[2], q0[3], R);
        double mi[3], ni[3];
        for (int c = 0; c < 3; ++c) {
            mi[c] = (i == 0) ? M_tip[c] : m[c * M + (i - 1)];
            ni[c] = (i == 0) ? F_tip[c] : n[c * M + (i - 1)];
        }
        for (int c = 0; c < 3; ++c) {
            Lambda[c * N + i] = R[0 * 3 + c] * mi[0] + R[1 * 3 + c] * mi[1] + R[2 * 3 + c] * mi[2];
            Lambda[(3 + c) * N + i] = R[0 * 3 + c] * ni[0] + R[1 * 3 + c] * ni[1] + R[2 * 3 + c] * ni[2];
        }
    }
}

/* Local-frame statics solved directly (rod_modeling.pdf eqs. 1.29, 2.18, collocated as Ch. 3):
 *   N' = -K^ N - R^T fbar,   C' = -K^ C - Gamma^ N - R^T lbar,   N(1) = R(1)^T F_tip,  C(1) = R(1)^T M_tip
 * i.e. the strain-dependent operator (D_TT (x) I3 + blockdiag K^_i) over the nodes 1..N-1 with the tip node eliminated,
 * one partial-pivot LU, two solves.  Lambda is [6][N], couple first (the [k; gamma] ordering of ad(), utilities.h:27-37).
 * Returns 0, or k+1 if the LU met a zero pivot. */
int sri_oracle_wrench_local_solve(const sri_oracle_ops* o, const double* K, const double* Q, const double* q0,
                                  const double* Gamma, const double* fbar, const double* lbar, const double* F_tip,
                                  const double* M_tip, double* Lambda)
{
    const int N = o->N, M = o->M, n = 3 * M;
    static const double q_default[4] = {1.0, 0.0, 0.0, 0.0};
    if (!q0) q0 = q_default;
    double* A = (double*)calloc((size_t)n * n, sizeof(double));
    double* R = (double*)malloc(sizeof(double) * 9 * N);
    double* b = (double*)malloc(sizeof(double) * n);
    int* piv = (int*)malloc(sizeof(int) * n);
    for (int i = 0; i < N; ++i) {
        if (i < M) quat_to_rot(Q[i], Q[i + M], Q[i + 2 * M], Q[i + 3 * M], R + 9 * i);
        else quat_to_rot(q0[0], q0[1], q0[2], q0[3], R + 9 * i);
    }
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i)
            for (int a = 0; a < 3; ++a) A[IDX(3 * i + a, 3 * j + a, n)] = o->D_TT[IDX(i, j, M)];
    for (int i = 0; i < M; ++i) {
        const double k0 = K[i + 1], k1 = K[N + i + 1], k2 = K[2 * N + i + 1];  /* node i+1 */
        A[IDX(3 * i + 0, 3 * i + 1, n)] += -k2; A[IDX(3 * i + 0, 3 * i + 2, n)] += k1;
        A[IDX(3 * i + 1, 3 * i + 0, n)] += k2;  A[IDX(3 * i + 1, 3 * i + 2, n)] += -k0;
        A[IDX(3 * i + 2, 3 * i + 0, n)] += -k1; A[IDX(3 * i + 2, 3 * i + 1, n)] += k0;
    }
    const int info = lu_factor(n, A, piv);
    double N0[3], C0[3];
    for (int c = 0; c < 3; ++c) {
        N0[c] = R[0 * 3 + c] * F_tip[0] + R[1 * 3 + c] * F_tip[1] + R[2 * 3 + c] * F_tip[2];
        C0[c] = R[0 * 3 + c] * M_tip[0] + R[1 * 3 + c] * M_tip[1] + R[2 * 3 + c] * M_tip[2];
    }
    for (int i = 0; i < M; ++i) {
        const double* Ri = R + 9 * (i + 1);
        for (int c = 0; c < 3; ++c) {
            double rf = 0.0;
            if (fbar) rf = Ri[0 * 3 + c] * fbar[i + 1] + Ri[1 * 3 + c] * fbar[N + i + 1] + Ri[2 * 3 + c] * fbar[2 * N + i + 1];
            b[3 * i + c] = -rf - o->D_TI[i] * N0[c];
        }
    }
    lu_solve(n, A, piv, b, 1);
    for (int c = 0; c < 3; ++c) { Lambda[(3 + c) * N] = N0[c]; Lambda[c * N] = C0[c]; }
    for (int i = 0; i < M; ++i)
        for (int c = 0; c < 3; ++c) Lambda[(3 + c) * N + i + 1] = b[3 * i + c];
    for (int i = 0; i < M; ++i) {
        const double* Ri = R + 9 * (i + 1);
        const double g0 = Gamma ? Gamma[i + 1] : 1.0, g1 = Gamma ? Gamma[N + i + 1] : 0.0, g2 = Gamma ? Gamma[2 * N + i + 1] : 0.0;
        const double n0 = Lambda[3 * N + i + 1], n1 = Lambda[4 * N + i + 1], n2 = Lambda[5 * N + i + 1];
        const double gx[3] = {g1 * n2 - g2 * n1, g2 * n0 - g0 * n2, g0 * n1 - g1 * n0};
        for (int c = 0; c < 3; ++c) {
            double rl = 0.0;
            if (lbar) rl = Ri[0 * 3 + c] * lbar[i + 1] + Ri[1 * 3 + c] * lbar[N + i + 1] + Ri[2 * 3 + c] * lbar[2 * N + i + 1];
            b[3 * i + c] = -gx[c] - rl - o->D_TI[i] * C0[c];
        }
    }
    lu_solve(n, A, piv, b, 1);
    for (int i = 0; i < M; ++i)
        for (int c = 0; c < 3; ++c) Lambda[c * N + i + 1] = b[3 * i + c];
    free(A); free(R); free(b); free(piv);
    return info;
}

/* ---- batched driver (OpenMP over rods) ----------------------------------------------------------------- */

/* All four stages for `batch` rods.  Layouts match include/sri.h: K,Gamma,fbar,lbar [batch][3][N];
 * q0 [batch][4]; r0,F_tip,M_tip [batch][3]; Q [batch][4][M]; r,n,m [batch][3][M].  Optional pointers may be NULL.
 * Returns the number of rods whose LU hit a zero pivot. */
int sri_oracle_integrate_all_batch(int N, long batch, const double* K, const double* q0, const double* r0,
                                   const double* Gamma, const double* fbar, const double* lbar, const double* F_tip,
                                   const double* M_tip, double* Q, double* r, double* n, double* m,
                                   int explicit_inverse, int nthreads)
{
    sri_oracle_ops* o = sri_oracle_ops_create(N);
    const int M = N - 1, nn = 4 * M;
    int bad = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads) reduction(+ : bad)
    {
        double* work = (double*)malloc(sizeof(double) * (2 * (size_t)nn * nn + nn));
        int* iwork = (int*)malloc(sizeof(int) * nn);
        double* ntmp = (double*)malloc(sizeof(double) * 3 * M);
        double* qtmp = (double*)malloc(sizeof(double) * 4 * M);
#pragma omp for schedule(static)
        for (long b = 0; b < batch; ++b) {
            const double* Kb = K + b * 3 * N;
            const double* q0b = q0 ? q0 + b * 4 : 0;
            double* Qb = Q ? Q + b * 4 * M : qtmp;
            const double* Gb = Gamma ? Gamma + b * 3 * N : 0;
            if (sri_oracle_integrate_quaternions(o, Kb, q0b, Qb, explicit_inverse, work, iwork)) bad += 1;
            if (r) sri_oracle_integrate_position(o, Qb, Gb, r0 ? r0 + b * 3 : 0, r + b * 3 * M);
            if (n || m) {
                double* nb = n ? n + b * 3 * M : ntmp;
                sri_oracle_integrate_stress(o, fbar ? fbar + b * 3 * N : 0, F_tip + b * 3, nb);
                if (m)
                    sri_oracle_integrate_couple(o, Qb, q0b, Gb, nb, lbar ? lbar + b * 3 * N : 0, M_tip + b * 3,
                                                m + b * 3 * M);
            }
        }
        free(work); free(iwork); free(ntmp); free(qtmp);
    }
    sri_oracle_ops_destroy(o);
    return bad;
}

int sri_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- synthetic rods of SURVEY 8(d): Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), counter = (rod lo, rod hi,
 * stream, 0), key = seed.  Host twin of the device generator so that CPU baseline and GPU see identical rods. ---- */
#include <stdint.h>

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u01(uint32_t hi, uint32_t lo)
{
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

/* K_c(X) = alpha + beta (2X-1), alpha,beta ~ U(-2,2); F_tip, M_tip ~ U(-1,1)^3; fbar = (0,0,-g), g ~ U(0,1). */
void sri_oracle_generate_rods(int N, uint64_t seed, long first_rod, long batch, double* K, double* F_tip, double* M_tip,
                              double* fbar)
{
    double* x = (double*)malloc(sizeof(double) * N);
    sri_oracle_chebyshev_points(N, 1.0, x);
    for (long b = 0; b < batch; ++b) {
        const uint64_t rod = (uint64_t)(first_rod + b);
        uint32_t w[4];
        double u[14];
        for (int s = 0; s < 7; ++s) {
            philox4x32_10((uint32_t)rod, (uint32_t)(rod >> 32), (uint32_t)s, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
            u[2 * s] = u01(w[0], w[1]);
            u[2 * s + 1] = u01(w[2], w[3]);
        }
        if (K)
            for (int c = 0; c < 3; ++c) {
                const double alpha = 4.0 * u[2 * c] - 2.0, beta = 4.0 * u[2 * c + 1] - 2.0;
                for (int i = 0; i < N; ++i) K[(b * 3 + c) * N + i] = fma(beta, 2 * x[i] - 1, alpha);
            }
        if (F_tip) for (int c = 0; c < 3; ++c) F_tip[b * 3 + c] = 2.0 * u[6 + c] - 1.0;
        if (M_tip) for (int c = 0; c < 3; ++c) M_tip[b * 3 + c] = 2.0 * u[9 + c] - 1.0;
        if (fbar)
            for (int i = 0; i < N; ++i) {
                fbar[(b * 3 + 0) * N + i] = 0.0;
                fbar[(b * 3 + 1) * N + i] = 0.0;
                fbar[(b * 3 + 2) * N + i] = -u[12];
            }
    }
    free(x);
}

/* The modal coordinates behind sri_oracle_generate_rods' curvature: qe[b][3c+0] = alpha, [3c+1] = beta, [3c+2] = 0, i.e.
 * K_c = alpha P_0 + beta P_1(2X-1) in the reference's Phi<3,3> basis (include/utilities.h:49-67, main.cpp:17,69), so the
 * same rods can be handed to the reference's own integrateQuaternions()/integratePosition() (oracle/reference_harness.cpp). */
void sri_oracle_generate_modes(uint64_t seed, long first_rod, long batch, double* qe)
{
    for (long b = 0; b < batch; ++b) {
        const uint64_t rod = (uint64_t)(first_rod + b);
        uint32_t w[4];
        for (int c = 0; c < 3; ++c) {
            philox4x32_10((uint32_t)rod, (uint32_t)(rod >> 32), (uint32_t)c, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
            qe[b * 9 + 3 * c + 0] = 4.0 * u01(w[0], w[1]) - 2.0;
            qe[b * 9 + 3 * c + 1] = 4.0 * u01(w[2], w[3]) - 2.0;
            qe[b * 9 + 3 * c + 2] = 0.0;
        }
    }
}

/* raw Philox block for the known-answer test in tests/ */
void sri_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
