/*
 * CPU oracle, C restatement -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C, per-point statement of the reference's native hot path with OpenMP over points, exactly where the
 * reference has its `#pragma omp parallel for` (Work/python_libs/triangulation_c/triangulation.c:70,109).
 * The OpenCV calls of the reference are restated with a one-sided Jacobi SVD (the algorithm behind
 * cvSolve(DECOMP_SVD) / cv::SVD for small matrices):
 *   orc_linear_ls      triangulation.c:65-83    (cvSolve(A, b, x, DECOMP_SVD), singular values <= 2 eps sum(w) dropped)
 *   orc_iterative_ls   triangulation.c:104-161  (semantics 0) / triangulation.py:100-195 (semantics 1)
 *   orc_linear_eigen   triangulation.py:6-25    (cv2.triangulatePoints: rows 4 = OpenCV >= 3, rows 6 = OpenCV 2.4)
 *   orc_polynomial     triangulation.py:198-232 (cv2.correctMatches: per-point 3x3 SVD epipoles, Durand-Kerner
 *                                                solvePoly(100 iterations), cost scan over real parts; then linear_eigen)
 * Used by tests/ (cross-check against the NumPy oracle and cv2) and by bench.py's cpu_baseline / --impl reference.
 * Parity pinning: agrees with oracle/triangulation_oracle.py, which is pinned to the reference's golden .mat cells
 * and exec'd-reference fixtures (tests/test_host_and_abi.py::test_c_oracle_matches_numpy_oracle, tests/test_oracle_golden.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPS DBL_EPSILON

/* one-sided Jacobi: A is m x n (row-major, leading dim n), V n x n.  Columns of A become U*w. */
static void jacobi(double* A, int m, int n, double* V) {
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 30; ++sweep) {
        int changed = 0;
        for (int i = 0; i < n - 1; ++i) for (int j = i + 1; j < n; ++j) {
            double a = 0, b = 0, p = 0;
            for (int k = 0; k < m; ++k) { a += A[k*n+i]*A[k*n+i]; b += A[k*n+j]*A[k*n+j]; p += A[k*n+i]*A[k*n+j]; }
            if (!(fabs(p) > EPS * sqrt(a * b))) continue;
            changed = 1;
            p *= 2; double beta = a - b, gamma = hypot(p, beta), c, s;
            if (beta < 0) { double delta = (gamma - beta) * 0.5; s = sqrt(delta / gamma); c = p / (gamma * s * 2); }
            else { c = sqrt((gamma + beta) / (gamma * 2)); s = p / (gamma * c * 2); }
            for (int k = 0; k < m; ++k) { double t0 = c*A[k*n+i] + s*A[k*n+j], t1 = -s*A[k*n+i] + c*A[k*n+j]; A[k*n+i] = t0; A[k*n+j] = t1; }
            for (int k = 0; k < n; ++k) { double t0 = c*V[k*n+i] + s*V[k*n+j], t1 = -s*V[k*n+i] + c*V[k*n+j]; V[k*n+i] = t0; V[k*n+j] = t1; }
        }
        if (!changed) break;
    }
}

/* min-norm LS of the 4x3 system, OpenCV back-substitution threshold */
static void solve43(const double A0[12], const double b[4], double x[3]) {
    double A[12], V[9], w[3], utb[3], wsum = 0;
    memcpy(A, A0, sizeof(A));
    jacobi(A, 4, 3, V);
    for (int j = 0; j < 3; ++j) {
        double s = 0, d = 0;
        for (int k = 0; k < 4; ++k) { s += A[k*3+j]*A[k*3+j]; d += A[k*3+j]*b[k]; }
        w[j] = sqrt(s); utb[j] = d; wsum += w[j];
    }
    double thr = 2 * EPS * wsum;
    x[0] = x[1] = x[2] = 0;
    for (int j = 0; j < 3; ++j) {
        double coef = (w[j] > thr) ? utb[j] / (w[j] * w[j]) : ((w[j] == w[j]) ? 0.0 : w[j]);
        for (int k = 0; k < 3; ++k) x[k] += V[k*3+j] * coef;
    }
}

static void build_Ab(const double* u1, const double* u2, const double* P1, const double* P2, double A[12], double b[4]) {
    const double* u[2] = {u1, u2}; const double* P[2] = {P1, P2};
    for (int c = 0; c < 2; ++c) for (int r = 0; r < 2; ++r) {
        int row = 2 * c + r;
        for (int k = 0; k < 3; ++k) A[row*3+k] = u[c][r] * P[c][8+k] - P[c][4*r+k];
        b[row] = -(u[c][r] * P[c][11] - P[c][4*r+3]);
    }
}

void orc_linear_ls(const double* u1, const double* u2, const double* P1, const double* P2, double* x, uint8_t* status, int64_t n) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double A[12], b[4];
        build_Ab(u1 + 2*i, u2 + 2*i, P1, P2, A, b);
        solve43(A, b, x + 3*i);
        status[i] = 1;
    }
}

/* Diagnostics for the parity tests: conditioning s_max / s_min of the unweighted 4x3 system of every point. */
void orc_ls_condition(const double* u1, const double* u2, const double* P1, const double* P2, double* cond, int64_t n) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double A[12], b[4], V[9], lo = DBL_MAX, hi = 0;
        build_Ab(u1 + 2*i, u2 + 2*i, P1, P2, A, b);
        jacobi(A, 4, 3, V);
        for (int j = 0; j < 3; ++j) {
            double sq = 0;
            for (int k = 0; k < 4; ++k) sq += A[k*3+j]*A[k*3+j];
            sq = sqrt(sq);
            if (sq < lo) lo = sq;
            if (sq > hi) hi = sq;
            if (sq != sq) { lo = 0; hi = 1; }
        }
        cond[i] = hi / lo;
    }
}

/* margin (may be NULL): min over the evaluated convergence tests of | |d_new - d| - tol | -- a point whose margin is at
 * rounding level changes its iteration count with the last bit of the solve ("knife edge"). */
void orc_iterative_ls(const double* u1, const double* u2, const double* P1, const double* P2, double* x, int32_t* status,
                      int32_t* nsolves, double* margin, int64_t n, double tol, int py_semantics) {
    #pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t xi = 0; xi < n; ++xi) {
        double A[12], b[4], *xp = x + 3*xi;
        build_Ab(u1 + 2*xi, u2 + 2*xi, P1, P2, A, b);
        double d1 = 1, d2 = 1, d1n = 1, d2n = 1, mg = INFINITY;
        int i;
        for (i = 0; i < 10; ++i) {
            solve43(A, b, xp);
            d1n = P1[8]*xp[0] + P1[9]*xp[1] + P1[10]*xp[2] + P1[11];
            d2n = P2[8]*xp[0] + P2[9]*xp[1] + P2[10]*xp[2] + P2[11];
            { double m1 = fabs(fabs(d1n - d1) - tol), m2 = fabs(fabs(d2n - d2) - tol); if (m1 < mg) mg = m1; if (m2 < mg) mg = m2; }
            if ((fabs(d1n - d1) <= tol && fabs(d2n - d2) <= tol) || (!py_semantics && (d1n == 0 || d2n == 0))) break;
            double s1 = 1. / d1n, s2 = 1. / d2n;
            for (int k = 0; k < 6; ++k) { A[k] *= s1; A[6+k] *= s2; }
            b[0] *= s1; b[1] *= s1; b[2] *= s2; b[3] *= s2;
            d1 = d1n; d2 = d2n;
        }
        if (nsolves) nsolves[xi] = i < 10 ? i + 1 : 10;
        if (margin) margin[xi] = mg;
        if (py_semantics && i == 10) i = 9;
        int st = (i < 10) && (d1n > 0) && (d2n > 0);
        if (d1n <= 0) st -= 1;
        if (d2n <= 0) st -= 2;
        status[xi] = st;
    }
}

/* amp (may be NULL): s_1 / ((s_3 - s_4) |w|) -- how much a relative perturbation eps of the DLT matrix moves the
 * dehomogenised point; points where amp * eps approaches the tolerance are ill-posed for ANY implementation. */
static void eigen_point(const double* u1, const double* u2, const double* P1, const double* P2, int rows, double maxc,
                        double* x, uint8_t* st, double* amp) {
    double B[24], V[16];
    int per = rows / 2;
    const double* u[2] = {u1, u2}; const double* P[2] = {P1, P2};
    for (int c = 0; c < 2; ++c) for (int k = 0; k < 4; ++k) {
        B[(per*c+0)*4+k] = u[c][0] * P[c][8+k] - P[c][k];
        B[(per*c+1)*4+k] = u[c][1] * P[c][8+k] - P[c][4+k];
        if (per == 3) B[(per*c+2)*4+k] = u[c][0] * P[c][4+k] - u[c][1] * P[c][k];
    }
    jacobi(B, rows, 4, V);
    int jb = 0; double best = 0;
    for (int j = 0; j < 4; ++j) {
        double s = 0;
        for (int r = 0; r < rows; ++r) s += B[r*4+j]*B[r*4+j];
        if (j == 0 || s < best || s != s) { best = s; jb = j; }
    }
    double X[4];
    for (int k = 0; k < 4; ++k) X[k] = (best == best) ? V[k*4+jb] : best;
    if (amp) {
        double sv[4];
        for (int j = 0; j < 4; ++j) { double sq = 0; for (int r = 0; r < rows; ++r) sq += B[r*4+j]*B[r*4+j]; sv[j] = sqrt(sq); }
        for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 4; ++j) if (sv[j] > sv[i]) { double t = sv[i]; sv[i] = sv[j]; sv[j] = t; }
        *amp = sv[0] / (sv[2] - sv[3]) / fabs(X[3]);
        if (!(*amp == *amp)) *amp = INFINITY;
    }
    for (int k = 0; k < 3; ++k) x[k] = X[k] / X[3];
    double m = fmax(fmax(fabs(x[0]), fabs(x[1])), fabs(x[2]));
    *st = (x[0] == x[0] && x[1] == x[1] && x[2] == x[2] && m <= maxc) ? 1 : 0;
}

void orc_linear_eigen(const double* u1, const double* u2, const double* P1, const double* P2, double* x, uint8_t* status,
                      double* amp, int64_t n, double maxc, int rows) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) eigen_point(u1 + 2*i, u2 + 2*i, P1, P2, rows, maxc, x + 3*i, status + i, amp ? amp + i : 0);
}

/* smallest right singular vector of a 3x3 (row-major) */
static void null3(const double M[9], double e[3]) {
    double A[9], V[9];
    memcpy(A, M, sizeof(A));
    jacobi(A, 3, 3, V);
    int jb = 0; double best = 0;
    for (int j = 0; j < 3; ++j) { double s = 0; for (int r = 0; r < 3; ++r) s += A[r*3+j]*A[r*3+j]; if (j == 0 || s < best) { best = s; jb = j; } }
    for (int k = 0; k < 3; ++k) e[k] = V[k*3+jb];
}

static double hs_cost(double t, double a, double b, double c, double d, double f1, double f2) {
    return t*t / (1 + f1*f1*t*t) + (c*t+d)*(c*t+d) / ((a*t+b)*(a*t+b) + f2*f2*(c*t+d)*(c*t+d));
}

static void correct_point(const double F[9], const double* p1, const double* p2, double* n1, double* n2) {
    double x1 = p1[0], y1 = p1[1], x2 = p2[0], y2 = p2[1];
    double G[9], TFT[9], TFTt[9], e1[3], e2[3];
    for (int r = 0; r < 3; ++r) { G[r*3+0] = F[r*3+0]; G[r*3+1] = F[r*3+1]; G[r*3+2] = F[r*3+0]*x1 + F[r*3+1]*y1 + F[r*3+2]; }
    for (int k = 0; k < 3; ++k) { TFT[k] = G[k]; TFT[3+k] = G[3+k]; TFT[6+k] = x2*G[k] + y2*G[3+k] + G[6+k]; }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) TFTt[r*3+c] = TFT[c*3+r];
    null3(TFT, e1); null3(TFTt, e2);
    double s1 = sqrt(e1[0]*e1[0] + e1[1]*e1[1]), s2 = sqrt(e2[0]*e2[0] + e2[1]*e2[1]);
    for (int k = 0; k < 3; ++k) { e1[k] /= s1; e2[k] /= s2; }
    if (e1[2] < 0) for (int k = 0; k < 3; ++k) e1[k] = -e1[k];
    if (e2[2] < 0) for (int k = 0; k < 3; ++k) e2[k] = -e2[k];
    double f1 = e1[2], f2 = e2[2];
    /* RTFTR = R2 TFT R1^T, entries (1,1),(1,2),(2,1),(2,2) */
    double h01 = -TFT[0]*e1[1] + TFT[1]*e1[0], h11 = -TFT[3]*e1[1] + TFT[4]*e1[0], h21 = -TFT[6]*e1[1] + TFT[7]*e1[0];
    double a = -e2[1]*h01 + e2[0]*h11, b = -e2[1]*TFT[2] + e2[0]*TFT[5], c = h21, d = TFT[8];
    double f1s = f1*f1, f2s = f2*f2;
    double q2 = a*a + f2s*c*c, q1 = 2*(a*b + f2s*c*d), q0 = b*b + f2s*d*d;
    double e = a*d - b*c, r2 = a*c, r1 = a*d + b*c, r0 = b*d, w2 = 2*f1s, w4 = f1s*f1s;
    double k[7] = { -e*r0, q0*q0 - e*r1, 2*q0*q1 - e*(r2 + w2*r0), q1*q1 + 2*q0*q2 - e*w2*r1,
                    2*q1*q2 - e*(w2*r2 + w4*r0), q2*q2 - e*w4*r1, -e*w4*r2 };
    int finite = 1;
    for (int i = 0; i < 7; ++i) if (!(fabs(k[i]) <= DBL_MAX)) finite = 0;
    double tmin = DBL_MAX;
    if (finite) {
        int n = 6;
        for (; n > 1; --n) if (fabs(k[n]) > EPS) break;
        double zr[6], zi[6], pr = 1, pi = 0;
        for (int i = 0; i < n; ++i) { zr[i] = pr; zi[i] = pi; double nr = pr - pi, ni = pr + pi; pr = nr; pi = ni; }
        for (int iter = 0; iter < 100; ++iter) {
            double maxdiff = 0;
            for (int i = 0; i < n; ++i) {
                double qr = zr[i], qi = zi[i], nr = k[n], ni = 0, dr = k[n], di = 0;
                for (int j = 0; j < n; ++j) {
                    double tr = nr*qr - ni*qi + k[n-j-1], ti = nr*qi + ni*qr; nr = tr; ni = ti;
                    if (j != i) {
                        double er = qr - zr[j], ei = qi - zi[j];
                        if (er != 0 || ei != 0) { double ur = dr*er - di*ei, ui = dr*ei + di*er; dr = ur; di = ui; }
                    }
                }
                double den = dr*dr + di*di, sr = (nr*dr + ni*di) / den, si = (ni*dr - nr*di) / den;
                zr[i] = qr - sr; zi[i] = qi - si;
                double mag = sqrt(sr*sr + si*si);
                if (mag > maxdiff) maxdiff = mag;
            }
            if (!(maxdiff > 0)) break;
        }
        double sval = 1. / (f1*f1) + c*c / (a*a + f2s*c*c);
        for (int i = 0; i < n; ++i) { double s = hs_cost(zr[i], a, b, c, d, f1, f2); if (s < sval) { sval = s; tmin = zr[i]; } }
    }
    if (tmin == DBL_MAX) { n1[0] = n1[1] = n2[0] = n2[1] = NAN; return; }
    double t = tmin;
    double hz = t*t*f1s + 1, hx = t*t*f1 / hz, hy = t / hz;
    n1[0] = e1[0]*hx - e1[1]*hy + x1; n1[1] = e1[1]*hx + e1[0]*hy + y1;
    double ctd = c*t + d, atb = a*t + b;
    hz = f2s*ctd*ctd + atb*atb; hx = f2*ctd*ctd / hz; hy = -atb*ctd / hz;
    n2[0] = e2[0]*hx - e2[1]*hy + x2; n2[1] = e2[1]*hx + e2[0]*hy + y2;
}

void orc_correct_matches(const double* F, const double* u1, const double* u2, double* n1, double* n2, int64_t n) {
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) correct_point(F, u1 + 2*i, u2 + 2*i, n1 + 2*i, n2 + 2*i);
}

void orc_polynomial(const double* F, const double* u1, const double* u2, const double* P1, const double* P2, double* x,
                    uint8_t* status, double* amp, int64_t n, double maxc, int rows) {
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        double c1[2], c2[2];
        correct_point(F, u1 + 2*i, u2 + 2*i, c1, c2);
        eigen_point(c1, c2, P1, P2, rows, maxc, x + 3*i, status + i, amp ? amp + i : 0);   /* amp of the CORRECTED match */
    }
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
