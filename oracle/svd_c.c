/* ORACLE (test infrastructure, never the product path): C port of oracle/svd.py:fit_epoch -- the sequential SGD pass of
 * /root/reference/src/origin_models/svd/SVD.py:187-221 (float64, rating by rating, item vector updated first and the
 * user vector from the UPDATED item vector, bias rule error * bias) -- for sizes the Python loop cannot reach in
 * seconds.  Checked against oracle/svd.py and tests/golden/svd_golden.npz by tests/test_oracle_svd.py.
 * Built by oracle/svd.py:build_c (gcc -O2 -ffp-contract=off). */
#include <stdint.h>

void svd_fit_epoch(const int32_t* users, const int32_t* items, const double* ratings, int64_t n, double* P, double* Q,
                   double* bu, double* bi, double mu, int32_t d, double lr, double emb_reg, double bias_reg) {
  for (int64_t k = 0; k < n; ++k) {
    double* p = P + (int64_t)users[k] * d;
    double* q = Q + (int64_t)items[k] * d;
    double dot = 0.0;
    for (int c = 0; c < d; ++c) dot += q[c] * p[c];
    const double b_u = bu[users[k]], b_i = bi[items[k]];
    const double e = ratings[k] - (b_u + b_i + mu + dot);
    for (int c = 0; c < d; ++c) {
      const double qn = q[c] + lr * (e * p[c] - emb_reg * q[c]);
      p[c] = p[c] + lr * (e * qn - emb_reg * p[c]);
      q[c] = qn;
    }
    bu[users[k]] = b_u + lr * (e * b_u - bias_reg * b_u);
    bi[items[k]] = b_i + lr * (e * b_i - bias_reg * b_i);
  }
}
