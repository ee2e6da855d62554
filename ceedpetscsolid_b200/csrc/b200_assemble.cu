// Coarse-level matrix assembly for the /gpu/b200 backend (sm_100a, FP64).
//
// The reference assembles the coarse (p = 1) Jacobian for PCGAMG by colouring the matrix-free operator
// (FormJacobian -> SNESComputeJacobianDefaultColor, /root/reference/src/misc.c:151-183, 81 operator
// applications per Newton step on a hexahedral mesh).  Every one of those applications re-reads the whole
// Jacobian cache; the kernels here read it ONCE:
//
//   k_assemble_p1        element matrices  K_e = sum_q G_q^T C_q G_q  of a trilinear (P = 2) level straight
//                        from the Jacobian cache -- the values array of CeedOperatorLinearAssemble (COO, one
//                        24 x 24 block per element);
//   k_stencil27_galerkin A_H = P^T A_h P for 27-point block stencils on 2:1 nested node lattices with the
//                        trilinear index-space prolongation: the Galerkin hierarchy of the h-multigrid that
//                        stands in for GAMG's (elasticity.c:569-585).
#include "b200_qf.cuh"

namespace b200 {

template <int Q> struct AsmMats {
  double B[Q * 2];  // interp1d [Q][2]
  double D[Q * 2];  // grad1d   [Q][2]
};

// One thread per (element, column dof); lanes element-fastest inside the group of EB elements that shares a
// q-blocked slab.  The thread walks the Q^3 points once, pushes the column's unit gradient through the point
// Jacobian and accumulates its 24 row entries in registers.
// FP64 work: Q^3 x 24 x (~135 point Jacobian + 72 test-side + 12 shape) ~ 0.66 MFLOP per element at Q = 5;
// bytes: NC x Q^3 x 8 read (16 kB) + 4.6 kB written per element  ->  compute-bound (AI ~ 30 flop/B).
template <int Q, int PROB>
__global__ void __launch_bounds__(24 * elems_per_block(Q))
k_assemble_p1(const __grid_constant__ AsmMats<Q> am, const __grid_constant__ Material mt, int nelem,
              const double *__restrict__ jcp, double *__restrict__ values) {
  constexpr int EB = elems_per_block(Q), T = Q * Q, Q3 = Q * Q * Q, NC = JCache<PROB>::N;
  const int tid = threadIdx.x;
  const int col = tid / EB, eb = tid - col * EB;
  const int blk = blockIdx.x;
  const int rem = nelem - blk * EB;
  const int ebn = rem < EB ? rem : EB;
  if (eb >= ebn) return;
  const int j = col / 3, cb = col - 3 * j;
  const int jx = j & 1, jy = (j >> 1) & 1, jz = j >> 2;
  const size_t ebt = (size_t)ebn * T;
  const double *jce = jcp + (size_t)blk * EB * NC * Q3 + eb;
  double acc[24];
#pragma unroll
  for (int r = 0; r < 24; r++) acc[r] = 0;
#pragma unroll 1
  for (int qz = 0; qz < Q; qz++) {
    const double bz[2] = {am.B[qz * 2], am.B[qz * 2 + 1]}, dz[2] = {am.D[qz * 2], am.D[qz * 2 + 1]};
#pragma unroll 1
    for (int qy = 0; qy < Q; qy++) {
      const double by[2] = {am.B[qy * 2], am.B[qy * 2 + 1]}, dy[2] = {am.D[qy * 2], am.D[qy * 2 + 1]};
      const double *jct = jce + (size_t)(qy + Q * qz) * ebn;
      // shape products that do not depend on qx: [jy'][jz'] for (B_y B_z), (D_y B_z), (B_y D_z)
      double bb[2][2], db[2][2], bd[2][2];
#pragma unroll
      for (int u = 0; u < 2; u++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
          bb[u][v] = by[u] * bz[v];
          db[u][v] = dy[u] * bz[v];
          bd[u][v] = by[u] * dz[v];
        }
#pragma unroll 1
      for (int qx = 0; qx < Q; qx++) {
        double jc[NC];
#pragma unroll
        for (int n = 0; n < NC; n++) jc[n] = __ldg(jct + (size_t)(n * Q + qx) * ebt);
        const double bx[2] = {am.B[qx * 2], am.B[qx * 2 + 1]}, dx[2] = {am.D[qx * 2], am.D[qx * 2 + 1]};
        double g[3];
        g[0] = (jx ? dx[1] : dx[0]) * (jy ? (jz ? bb[1][1] : bb[1][0]) : (jz ? bb[0][1] : bb[0][0]));
        g[1] = (jx ? bx[1] : bx[0]) * (jy ? (jz ? db[1][1] : db[1][0]) : (jz ? db[0][1] : db[0][0]));
        g[2] = (jx ? bx[1] : bx[0]) * (jy ? (jz ? bd[1][1] : bd[1][0]) : (jz ? bd[0][1] : bd[0][0]));
        double H[3][3], W[3][3];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int m = 0; m < 3; m++) H[c][m] = c == cb ? g[m] : 0.;
        jacobian_point<PROB>(mt, jc, H, W);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int ix = i & 1, iy = (i >> 1) & 1, iz = i >> 2;
          const double g0 = dx[ix] * bb[iy][iz], g1 = bx[ix] * db[iy][iz], g2 = bx[ix] * bd[iy][iz];
#pragma unroll
          for (int a = 0; a < 3; a++) acc[i * 3 + a] += g0 * W[a][0] + g1 * W[a][1] + g2 * W[a][2];
        }
      }
    }
  }
  double *out = values + ((size_t)(blk * EB + eb) * 24 + col) * 24;
#pragma unroll
  for (int r = 0; r < 24; r += 2) *reinterpret_cast<double2 *>(out + r) = make_double2(acc[r], acc[r + 1]);
}

template <int Q, int PROB>
static int launch_assemble_p1(const Material &mt, int nelem, const double *hB, const double *hD, const double *jc,
                              double *values) {
  AsmMats<Q> am;
  for (int i = 0; i < Q * 2; i++) {
    am.B[i] = hB[i];
    am.D[i] = hD[i];
  }
  constexpr int EB = elems_per_block(Q);
  const int nblk = (nelem + EB - 1) / EB;
  if (nblk == 0) return 0;
  k_assemble_p1<Q, PROB><<<nblk, 24 * EB, 0, g_stream>>>(am, mt, nelem, jc, values);
  B200_LAUNCH_CHECK("k_assemble_p1");
  return 0;
}

template <int PROB>
static int dispatch_assemble_p1(int Q, const Material &mt, int nelem, const double *hB, const double *hD,
                                const double *jc, double *values) {
  switch (Q) {
    case 2: return launch_assemble_p1<2, PROB>(mt, nelem, hB, hD, jc, values);
    case 3: return launch_assemble_p1<3, PROB>(mt, nelem, hB, hD, jc, values);
    case 4: return launch_assemble_p1<4, PROB>(mt, nelem, hB, hD, jc, values);
    case 5: return launch_assemble_p1<5, PROB>(mt, nelem, hB, hD, jc, values);
  }
  return set_error_msg("element-matrix assembly: Q not instantiated (2..5)");
}

// ---------------------------------------------------------------------------------------------------------
// Galerkin coarse stencil.  Fine lattice Nf = 2 Nc - 1 per axis, coarse node I sits on fine node 2I,
// P[i, I] = prod_d w(i_d - 2 I_d), w(0) = 1, w(+-1) = 1/2.
//   A_H[(I,a),(I+O,b)] = sum_{delta, o in {-1,0,1}^3} P[2I+delta, I] A_h[(2I+delta,a),(2I+delta+o,b)] P[2I+delta+o, I+O]
// One thread per (coarse row dof (I,a), coarse neighbour O): reads 27 x (<= 27) x 3 fine entries, no atomics.
__global__ void k_stencil27_galerkin(int Ncx, int Ncy, int Ncz, const double *__restrict__ fv, double *__restrict__ cv) {
  const int Nfx = 2 * Ncx - 1, Nfy = 2 * Ncy - 1, Nfz = 2 * Ncz - 1;
  const size_t nc = (size_t)3 * Ncx * Ncy * Ncz, nf = (size_t)3 * Nfx * Nfy * Nfz;
  const size_t total = nc * 27;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t row = idx % nc;  // rows fastest: coalesced stores per neighbour slot
    const int O = (int)(idx / nc);
    const int Ox = O % 3 - 1, Oy = (O / 3) % 3 - 1, Oz = O / 9 - 1;
    const int a = (int)(row % 3), node = (int)(row / 3);
    const int I = node % Ncx, J = (node / Ncx) % Ncy, K = node / (Ncx * Ncy);
    double s[3] = {0, 0, 0};
    const bool inside = I + Ox >= 0 && I + Ox < Ncx && J + Oy >= 0 && J + Oy < Ncy && K + Oz >= 0 && K + Oz < Ncz;
    if (inside) {
      for (int dz = -1; dz <= 1; dz++) {
        const int fz = 2 * K + dz;
        if (fz < 0 || fz >= Nfz) continue;
        for (int dy = -1; dy <= 1; dy++) {
          const int fy = 2 * J + dy;
          if (fy < 0 || fy >= Nfy) continue;
          for (int dx = -1; dx <= 1; dx++) {
            const int fx = 2 * I + dx;
            if (fx < 0 || fx >= Nfx) continue;
            const double wl = (dx ? 0.5 : 1.0) * (dy ? 0.5 : 1.0) * (dz ? 0.5 : 1.0);
            const size_t frow = (size_t)3 * (fx + (size_t)Nfx * (fy + (size_t)Nfy * fz)) + a;
            for (int oz = -1; oz <= 1; oz++) {
              const int tz = dz + oz - 2 * Oz;  // fine column relative to coarse column I+O
              if (tz < -1 || tz > 1 || fz + oz < 0 || fz + oz >= Nfz) continue;
              for (int oy = -1; oy <= 1; oy++) {
                const int ty = dy + oy - 2 * Oy;
                if (ty < -1 || ty > 1 || fy + oy < 0 || fy + oy >= Nfy) continue;
                for (int ox = -1; ox <= 1; ox++) {
                  const int tx = dx + ox - 2 * Ox;
                  if (tx < -1 || tx > 1 || fx + ox < 0 || fx + ox >= Nfx) continue;
                  const double w = wl * (tx ? 0.5 : 1.0) * (ty ? 0.5 : 1.0) * (tz ? 0.5 : 1.0);
                  const int o = (ox + 1) + 3 * (oy + 1) + 9 * (oz + 1);
#pragma unroll
                  for (int b = 0; b < 3; b++) s[b] += w * __ldg(fv + (size_t)(o * 3 + b) * nf + frow);
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 3; b++) cv[(size_t)(O * 3 + b) * nc + row] = s[b];
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_assemble_p1(int problem, const b200_physics *phys, int nelem, int Q, const double *hB,
                                const double *hD, const double *d_jcache, double *d_values) {
  const Material mt = make_material(phys);
  switch (problem) {
    case B200_PROB_LINELAS: return dispatch_assemble_p1<B200_PROB_LINELAS>(Q, mt, nelem, hB, hD, d_jcache, d_values);
    case B200_PROB_HYPERSS: return dispatch_assemble_p1<B200_PROB_HYPERSS>(Q, mt, nelem, hB, hD, d_jcache, d_values);
    case B200_PROB_HYPERFS: return dispatch_assemble_p1<B200_PROB_HYPERFS>(Q, mt, nelem, hB, hD, d_jcache, d_values);
  }
  return set_error_msg("b200_assemble_p1: unknown problem");
}

extern "C" int b200_stencil27_galerkin(int Ncx, int Ncy, int Ncz, const double *d_fine, double *d_coarse) {
  const size_t total = (size_t)81 * Ncx * Ncy * Ncz;
  if (total == 0) return 0;
  size_t nb = (total + 127) / 128;
  if (nb > 148 * 64) nb = 148 * 64;
  k_stencil27_galerkin<<<(unsigned)nb, 128, 0, g_stream>>>(Ncx, Ncy, Ncz, d_fine, d_coarse);
  B200_LAUNCH_CHECK("k_stencil27_galerkin");
  return 0;
}
