/* CPU oracle for Phi -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/mgrit_oracle.py).
 *
 * Plain-C restatement of the two linear solves the reference delegates to SciPy's SuperLU
 * (scipy.sparse.linalg.spsolve, scipy 1.18.1 in this image; the source is not under /root/reference):
 *   heat/heat_1d.py:213      (I + dt*L) y = b,  L = (a/dx^2) tridiag(-1, 2, -1)
 *   advection/advection_1d.py:140  (I + dt*L) y = b,  L = (c/dx) (I - S_periodic)   (first-order upwind)
 * Both are solved here by the textbook direct method (Thomas elimination / forward substitution with the
 * closed periodic term), which is what a sparse LU without fill reduces to for these matrices.
 * Used only to make large parity cases finish in seconds; pinned against spsolve in tests/test_oracle.py.
 */
#include <stddef.h>

/* r = dt * a / dx^2.  work: n doubles of scratch. */
void oracle_heat1d_be(int n, double r, const double *b, double *out, double *work)
{
    const double d = 1.0 + 2.0 * r;
    double piv = d;
    work[0] = -r / piv;
    out[0] = b[0] / piv;
    for (int i = 1; i < n; ++i) {
        piv = d + r * work[i - 1];
        work[i] = -r / piv;
        out[i] = (b[i] + r * out[i - 1]) / piv;
    }
    for (int i = n - 2; i >= 0; --i)
        out[i] -= work[i] * out[i + 1];
}

/* nu = dt * c / dx.  Row i: (1 + nu) y_i - nu y_{i-1} = b_i, with y_{-1} = y_{n-1}. */
void oracle_advection1d_be(int n, double nu, const double *b, double *out)
{
    const double rho = nu / (1.0 + nu);
    const double sig = 1.0 / (1.0 + nu);
    /* y_i = p_i + q_i * y_{n-1}: first pass accumulates p (in out) and q_{n-1} */
    double p = 0.0, q = 1.0;
    for (int i = 0; i < n; ++i) {
        p = sig * b[i] + rho * p;
        q = rho * q;
        out[i] = p;
    }
    const double last = out[n - 1] / (1.0 - q);
    double w = rho;
    for (int i = 0; i < n; ++i) {
        out[i] += w * last;
        w *= rho;
    }
}
