// Test-only: runs the device point functions of csrc/b200_qf.cuh ON THE HOST (they are __host__ __device__) so
// that the CPU test suite can compare the very arithmetic the kernels execute with the reference QFunctions.
// Q-vector layout of the libCEED user QFunctions: in[(row*ncol + col)*Q + i].
#include "b200_qf.cuh"

using namespace b200;

namespace b200 {
cudaStream_t g_stream = 0;
unsigned long long g_launches = 0;
int set_error(cudaError_t, const char *) { return 1; }
int set_error_msg(const char *) { return 1; }
}  // namespace b200

static void load(const double *du, const double *qd, int Q, int i, double (&H)[3][3], double (&A)[3][3], double &w) {
  for (int d = 0; d < 3; d++)
    for (int c = 0; c < 3; c++) H[c][d] = du[(d * 3 + c) * Q + i];
  w = qd[i];
  for (int r = 0; r < 3; r++)
    for (int s = 0; s < 3; s++) A[r][s] = qd[(1 + 3 * r + s) * Q + i];
}

template <int PROB> static void post(const Material &mt, int Q, const double *du, const double *qd, double *energy, double *diag5) {
  for (int i = 0; i < Q; i++) {
    double H[3][3], A[3][3], w, p[5];
    load(du, qd, Q, i, H, A, w);
    post_point<PROB>(mt, A, H, p);
    energy[i] = p[4] * w;
    for (int c = 0; c < 5; c++) diag5[c * Q + i] = p[c];
  }
}

// residual (F): out dv[(k*3+c)*Q+i], gradu[(c*3+k)*Q+i];  Jacobian through the cache: jcache_point + jacobian_point
template <int PROB> static void resjac(const Material &mt, int Q, const double *du, const double *ddu, const double *qd,
                                       double *dv, double *gradu, double *ddv) {
  for (int i = 0; i < Q; i++) {
    double H[3][3], A[3][3], w, g[3][3] = {{0}}, W[3][3];
    load(du, qd, Q, i, H, A, w);
    if (PROB == B200_PROB_LINELAS) linelas_point(mt, w, A, H, W);
    else if (PROB == B200_PROB_HYPERSS) hyperss_f_point(mt, w, A, H, g, W);
    else hyperfs_f_point(mt, w, A, H, g, W);
    for (int k = 0; k < 3; k++)
      for (int c = 0; c < 3; c++) dv[(k * 3 + c) * Q + i] = W[c][k];
    double q10[10], gu[9], jc[JCache<PROB>::N], dH[3][3], dA[3][3], dw;
    for (int n = 0; n < 10; n++) q10[n] = qd[n * Q + i];
    for (int c = 0; c < 3; c++)
      for (int k = 0; k < 3; k++) {
        gu[c * 3 + k] = g[c][k];
        gradu[(c * 3 + k) * Q + i] = g[c][k];
      }
    jcache_point<PROB>(q10, gu, jc);
    load(ddu, qd, Q, i, dH, dA, dw);
    jacobian_point<PROB>(mt, jc, dH, W);
    for (int k = 0; k < 3; k++)
      for (int c = 0; c < 3; c++) ddv[(k * 3 + c) * Q + i] = W[c][k];
  }
}

// closed-form diagonal point blocks (diag_blocks_point, what k_fused_diag executes) next to the blocks obtained by
// pushing the nine unit gradients through jacobian_point: out[((c*3 + dp)*3 + d)*Q + i] = dW[c][d] / dH[c][dp]
template <int PROB> static void diagblocks(const Material &mt, int Q, const double *du, const double *qd, double *closed,
                                           double *probed) {
  const int vj[6] = {0, 1, 2, 1, 0, 0}, vk[6] = {0, 1, 2, 2, 2, 1};
  for (int i = 0; i < Q; i++) {
    double H[3][3], A[3][3], w, g[3][3] = {{0}}, W[3][3];
    load(du, qd, Q, i, H, A, w);
    if (PROB == B200_PROB_HYPERSS) hyperss_f_point(mt, w, A, H, g, W);
    else if (PROB == B200_PROB_HYPERFS) hyperfs_f_point(mt, w, A, H, g, W);
    double q10[10], gu[9], jc[JCache<PROB>::N], M[6], kappa;
    for (int n = 0; n < 10; n++) q10[n] = qd[n * Q + i];
    for (int c = 0; c < 3; c++)
      for (int k = 0; k < 3; k++) gu[c * 3 + k] = g[c][k];
    jcache_point<PROB>(q10, gu, jc);
    diag_blocks_point<PROB>(mt, jc, M, kappa);
    for (int c = 0; c < 3; c++) {
      double Ac[3][3];
      for (int t = 0; t < 6; t++)
        Ac[vj[t]][vk[t]] = Ac[vk[t]][vj[t]] = M[t] + kappa * jc[3 * vj[t] + c] * jc[3 * vk[t] + c];
      for (int dp = 0; dp < 3; dp++) {
        double U[3][3] = {{0}};
        U[c][dp] = 1.0;
        jacobian_point<PROB>(mt, jc, U, W);
        for (int d = 0; d < 3; d++) {
          closed[((c * 3 + dp) * 3 + d) * Q + i] = Ac[dp][d];
          probed[((c * 3 + dp) * 3 + d) * Q + i] = W[c][d];
        }
      }
    }
  }
}

extern "C" int qf_host_diagblocks(int prob, double nu, double E, int Q, const double *du, const double *qd, double *closed,
                                  double *probed) {
  b200_physics ph = {nu, E};
  const Material mt = make_material(&ph);
  if (prob == 0) diagblocks<B200_PROB_LINELAS>(mt, Q, du, qd, closed, probed);
  else if (prob == 1) diagblocks<B200_PROB_HYPERSS>(mt, Q, du, qd, closed, probed);
  else diagblocks<B200_PROB_HYPERFS>(mt, Q, du, qd, closed, probed);
  return 0;
}

extern "C" int qf_host_post(int prob, double nu, double E, int Q, const double *du, const double *qd, double *energy, double *diag5) {
  b200_physics ph = {nu, E};
  const Material mt = make_material(&ph);
  if (prob == 0) post<B200_PROB_LINELAS>(mt, Q, du, qd, energy, diag5);
  else if (prob == 1) post<B200_PROB_HYPERSS>(mt, Q, du, qd, energy, diag5);
  else post<B200_PROB_HYPERFS>(mt, Q, du, qd, energy, diag5);
  return 0;
}

extern "C" int qf_host_resjac(int prob, double nu, double E, int Q, const double *du, const double *ddu, const double *qd,
                              double *dv, double *gradu, double *ddv) {
  b200_physics ph = {nu, E};
  const Material mt = make_material(&ph);
  if (prob == 0) resjac<B200_PROB_LINELAS>(mt, Q, du, ddu, qd, dv, gradu, ddv);
  else if (prob == 1) resjac<B200_PROB_HYPERSS>(mt, Q, du, ddu, qd, dv, gradu, ddv);
  else resjac<B200_PROB_HYPERFS>(mt, Q, du, ddu, qd, dv, gradu, ddv);
  return 0;
}
