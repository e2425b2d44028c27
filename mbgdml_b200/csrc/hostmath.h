// Host-side 3x3 helpers for the periodic cell: inverse and Minkowski reduction.
//
// ase.geometry.find_mic (third-party; single call site reference mbgdml/periodic.py:78)
// falls back to general_find_mic on a Minkowski-reduced cell when the naive minimum
// image is not provably minimal.  The reduction runs once per cell on the host; the
// per-vector image search runs on device (geom.cuh: mic_general).  Algorithm: Nguyen &
// Stehle, "Low-dimensional lattice basis reduction revisited" (2009), as used by ASE.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <vector>

namespace mbg {
namespace hostmath {

using V3 = std::array<double, 3>;
using I3 = std::array<long long, 3>;

inline double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline V3 row_combo(const I3& h, const double* B) {  // h @ B
  V3 o;
  for (int i = 0; i < 3; ++i) o[i] = (double)h[0] * B[i] + (double)h[1] * B[3 + i] + (double)h[2] * B[6 + i];
  return o;
}
inline long long round_half_even(double x) { return (long long)std::nearbyint(x); }

inline bool inverse3(const double* A, double* inv) {
  const double a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], f = A[5], g = A[6], h = A[7], i = A[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  if (det == 0.0 || !std::isfinite(det)) return false;
  const bool diagonal = (b == 0 && c == 0 && d == 0 && f == 0 && g == 0 && h == 0);
  if (diagonal) {  // exact reciprocals, as LAPACK returns for a diagonal matrix
    for (int k = 0; k < 9; ++k) inv[k] = 0.0;
    inv[0] = 1.0 / a, inv[4] = 1.0 / e, inv[8] = 1.0 / i;
    return true;
  }
  const double r = 1.0 / det;
  inv[0] = (e * i - f * h) * r, inv[1] = (c * h - b * i) * r, inv[2] = (b * f - c * e) * r;
  inv[3] = (f * g - d * i) * r, inv[4] = (a * i - c * g) * r, inv[5] = (c * d - a * f) * r;
  inv[6] = (d * h - e * g) * r, inv[7] = (b * g - a * h) * r, inv[8] = (a * e - b * d) * r;
  return true;
}

inline void gauss_reduce(const double* B, I3& hu, I3& hv) {
  V3 u = row_combo(hu, B), v = row_combo(hv, B);
  std::vector<std::array<long long, 6>> seen;
  for (int it = 0; it < 5000; ++it) {
    const long long x = round_half_even(dot(u, v) / dot(u, u));
    I3 nhu = {hv[0] - x * hu[0], hv[1] - x * hu[1], hv[2] - x * hu[2]};
    hv = hu, hu = nhu;
    u = row_combo(hu, B), v = row_combo(hv, B);
    std::array<long long, 6> site = {hu[0], hu[1], hu[2], hv[0], hv[1], hv[2]};
    const bool cyc = std::find(seen.begin(), seen.end(), site) != seen.end();
    seen.push_back(site);
    if (dot(u, u) >= dot(v, v) || cyc) {
      std::swap(hu, hv);  // return (hv, hu)
      return;
    }
  }
}

inline void closest_vector_2d(const double t0[2], const double u[2], const double v[2], long long a[2]) {
  struct Cand { double r[2]; int c[2]; double n; };
  std::vector<Cand> cands;
  for (int i = -1; i <= 1; ++i)
    for (int j = -1; j <= 1; ++j) {
      Cand c;
      c.c[0] = i, c.c[1] = j;
      c.r[0] = i * u[0] + j * v[0], c.r[1] = i * u[1] + j * v[1];
      c.n = std::sqrt(c.r[0] * c.r[0] + c.r[1] * c.r[1]);
      cands.push_back(c);
    }
  std::stable_sort(cands.begin(), cands.end(), [](const Cand& x, const Cand& y) { return x.n < y.n; });
  cands.resize(7);
  a[0] = a[1] = 0;
  double t[2] = {t0[0], t0[1]};
  double dprev = INFINITY;
  for (int it = 0; it < 5000; ++it) {
    int best = 0;
    double dbest = INFINITY;
    for (int k = 0; k < 7; ++k) {
      const double dx = cands[k].r[0] + t[0], dy = cands[k].r[1] + t[1];
      const double d = std::sqrt(dx * dx + dy * dy);
      if (d < dbest) dbest = d, best = k;
    }
    if (best == 0 || dbest >= dprev) return;
    dprev = dbest;
    const double* r = cands[best].r;
    const long long kopt = round_half_even(-(t[0] * r[0] + t[1] * r[1]) / (r[0] * r[0] + r[1] * r[1]));
    a[0] += kopt * cands[best].c[0], a[1] += kopt * cands[best].c[1];
    t[0] = t0[0] + a[0] * u[0] + a[1] * v[0];
    t[1] = t0[1] + a[0] * u[1] + a[1] * v[1];
  }
}

// rcell = op @ cell, Minkowski reduced (fully periodic 3-D).  Returns false on failure.
inline bool minkowski_reduce(const double* B, double* rcell) {
  std::array<I3, 3> H = {I3{1, 0, 0}, I3{0, 1, 0}, I3{0, 0, 1}};
  double norms[3];
  for (int i = 0; i < 3; ++i) norms[i] = std::sqrt(B[3 * i] * B[3 * i] + B[3 * i + 1] * B[3 * i + 1] + B[3 * i + 2] * B[3 * i + 2]);
  std::vector<std::array<long long, 9>> seen;
  bool done = false;
  for (int it = 0; it < 5000 && !done; ++it) {
    int order[3] = {0, 1, 2};
    std::stable_sort(order, order + 3, [&](int x, int y) { return norms[x] < norms[y]; });
    std::array<I3, 3> Hs = {H[order[0]], H[order[1]], H[order[2]]};
    I3 hu = Hs[0], hv = Hs[1], hw = Hs[2];
    gauss_reduce(B, hu, hv);
    H = {hu, hv, hw};
    V3 R[3] = {row_combo(H[0], B), row_combo(H[1], B), row_combo(H[2], B)};
    const double nu = std::sqrt(dot(R[0], R[0]));
    V3 X = {R[0][0] / nu, R[0][1] / nu, R[0][2] / nu};
    const double vx = dot(R[1], X);
    V3 Y = {R[1][0] - X[0] * vx, R[1][1] - X[1] * vx, R[1][2] - X[2] * vx};
    const double ny = std::sqrt(dot(Y, Y));
    for (int k = 0; k < 3; ++k) Y[k] /= ny;
    double p[3][2];
    for (int k = 0; k < 3; ++k) p[k][0] = dot(R[k], X), p[k][1] = dot(R[k], Y);
    long long nb[2];
    closest_vector_2d(p[2], p[0], p[1], nb);
    for (int k = 0; k < 3; ++k) H[2][k] = nb[0] * H[0][k] + nb[1] * H[1][k] + H[2][k];
    for (int k = 0; k < 3; ++k) {
      R[k] = row_combo(H[k], B);
      norms[k] = dot(R[k], R[k]);
    }
    std::array<long long, 9> site = {H[0][0], H[0][1], H[0][2], H[1][0], H[1][1], H[1][2], H[2][0], H[2][1], H[2][2]};
    const bool cyc = std::find(seen.begin(), seen.end(), site) != seen.end();
    seen.push_back(site);
    if (norms[2] >= norms[1] || cyc) done = true;
  }
  if (!done) return false;
  for (int k = 0; k < 3; ++k) {
    const V3 r = row_combo(H[k], B);
    rcell[3 * k] = r[0], rcell[3 * k + 1] = r[1], rcell[3 * k + 2] = r[2];
  }
  return true;
}

// ase.geometry.minkowski_reduce(cell, pbc) with a non-periodic direction: two periodic directions are Gauss-reduced
// (periodic vectors permuted to the front, shorter first, reduced, permutation undone, handedness kept -- judged with
// the plane normal in place of the non-periodic vector); one or none: the cell is left alone.
inline bool minkowski_reduce_pbc(const double* B, const int pbc[3], double* rcell) {
  const int dim = (pbc[0] != 0) + (pbc[1] != 0) + (pbc[2] != 0);
  if (dim == 3) return minkowski_reduce(B, rcell);
  long long op[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  if (dim == 2) {
    int asc[3] = {0, 1, 2};
    std::stable_sort(asc, asc + 3, [&](int x, int y) { return (pbc[x] != 0) < (pbc[y] != 0); });
    const int perm[3] = {asc[2], asc[1], asc[0]};  // np.argsort(pbc, kind='merge')[::-1]
    double P[9];
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k) P[3 * i + k] = B[3 * perm[i] + perm[k]];
    const double n0 = std::sqrt(P[0] * P[0] + P[1] * P[1] + P[2] * P[2]), n1 = std::sqrt(P[3] * P[3] + P[4] * P[4] + P[5] * P[5]);
    I3 hu = {1, 0, 0}, hv = {0, 1, 0};
    if (n1 < n0) std::swap(hu, hv);  // op = eye[argsort(norms)], the non-periodic vector last
    gauss_reduce(P, hu, hv);
    const long long opP[3][3] = {{hu[0], hu[1], hu[2]}, {hv[0], hv[1], hv[2]}, {0, 0, 1}};
    int inv[3];
    for (int i = 0; i < 3; ++i) inv[perm[i]] = i;
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k) op[i][k] = opP[inv[i]][inv[k]];
    const int index = pbc[0] ? (pbc[1] ? 2 : 1) : 0;  // np.argmin(pbc)
    const int ia = (index + 1) % 3, ib = (index + 2) % 3;  // cell[index - 2], cell[index - 1]
    auto rowof = [&](const long long h[3], double out[3]) {
      for (int k = 0; k < 3; ++k) out[k] = (double)h[0] * B[k] + (double)h[1] * B[3 + k] + (double)h[2] * B[6 + k];
    };
    const double* a = B + 3 * ia;
    const double* b = B + 3 * ib;
    const double nrm[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    auto det_with_normal = [&](const double M[9]) {
      double C[9];
      for (int k = 0; k < 9; ++k) C[k] = M[k];
      for (int k = 0; k < 3; ++k) C[3 * index + k] = nrm[k];
      return C[0] * (C[4] * C[8] - C[5] * C[7]) - C[1] * (C[3] * C[8] - C[5] * C[6]) + C[2] * (C[3] * C[7] - C[4] * C[6]);
    };
    double R1[9];
    for (int i = 0; i < 3; ++i) rowof(op[i], R1 + 3 * i);
    if ((det_with_normal(B) > 0) != (det_with_normal(R1) > 0))
      for (int k = 0; k < 3; ++k) op[ib][k] = -op[ib][k];
  }
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k)
      rcell[3 * i + k] = (double)op[i][0] * B[k] + (double)op[i][1] * B[3 + k] + (double)op[i][2] * B[6 + k];
  return true;
}

}  // namespace hostmath
}  // namespace mbg
