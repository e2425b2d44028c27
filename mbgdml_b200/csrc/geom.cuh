// Per-fragment geometry on device: gather, minimum image, pair cut-off, acceptance
// criteria, inverse-distance descriptor.  Every discrete decision (cut-offs) is
// evaluated with explicitly rounded, un-fused FP64 operations in the reference's
// operation order so that accept masks reproduce the NumPy path bit for bit.
//
// Reference: mbgdml/periodic.py:61-107 (Cell.r_mic/d_mic -> ase.geometry.find_mic),
//            mbgdml/descriptors.py:65-201 (Criteria.accept, com_distance_sum,
//            max_atom_pair_dist), mbgdml/_gdml/desc.py:79-233 (_from_r),
//            mbgdml/predictors/gdml_predict.py:214-235 (slice -> MIC -> criteria).
#pragma once
#include "common.cuh"

namespace mbg {

// Relative guard band of the exact early-outs: a shortcut may only decide when the quantity is beyond its
// threshold by more than any rounding difference between the shortcut's arithmetic and the full evaluation.
constexpr double kGuard = 1.0 + 1e-9;

__device__ __forceinline__ double norm3_rn(double x, double y, double z) {
  // sqrt((x*x + y*y) + z*z) with each operation rounded (scipy pdist / np.linalg.norm over 3 components)
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// out = v @ M (M row-major 3x3), sequential un-fused accumulation.
__device__ __forceinline__ void vec_mat_rn(const double v[3], const double* M, double out[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
    out[i] = __dadd_rn(__dadd_rn(__dmul_rn(v[0], M[i]), __dmul_rn(v[1], M[3 + i])), __dmul_rn(v[2], M[6 + i]));
}

// ase.geometry.naive_find_mic: f = v cell^-1 ; f -= floor(f + 0.5) ; v' = f cell.  Returns |v'|.
__device__ __forceinline__ double mic_naive(const CellDev& s, const double v[3], double out[3]) {
  double f[3];
  vec_mat_rn(v, s.cell_inv, f);
#pragma unroll
  for (int i = 0; i < 3; ++i) f[i] = __dsub_rn(f[i], floor(__dadd_rn(f[i], 0.5)));
  vec_mat_rn(f, s.cell, out);
  return norm3_rn(out[0], out[1], out[2]);
}

// ase.geometry.general_find_mic on the Minkowski-reduced cell: wrap the periodic fractional coordinates into [0,1),
// then the shortest of (0,0,0) followed by itertools.product of [-1,0,1] over the periodic directions ([0] over the
// others); first minimum wins.  No periodic direction at all: the vector itself (find_mic copies it).
__device__ __noinline__ void mic_general(const CellDev& s, const double v[3], double out[3]) {
  const int l0 = s.pbc[0], l1 = s.pbc[1], l2 = s.pbc[2];
  if (!(l0 | l1 | l2)) {
    out[0] = v[0], out[1] = v[1], out[2] = v[2];
    return;
  }
  double f[3], pos[3];
  vec_mat_rn(v, s.rcell_inv, f);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (s.pbc[i]) f[i] = __dsub_rn(f[i], floor(f[i]));
  vec_mat_rn(f, s.rcell, pos);
  double best = norm3_rn(pos[0], pos[1], pos[2]);
  out[0] = pos[0], out[1] = pos[1], out[2] = pos[2];
#pragma unroll 1
  for (int h = -l0; h <= l0; ++h)
#pragma unroll 1
    for (int k = -l1; k <= l1; ++k)
#pragma unroll 1
      for (int l = -l2; l <= l2; ++l) {
        const double hkl[3] = {(double)h, (double)k, (double)l};
        double sh[3], x[3];
        vec_mat_rn(hkl, s.rcell, sh);
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = __dadd_rn(pos[i], sh[i]);
        const double len = norm3_rn(x[0], x[1], x[2]);
        if (len < best) {
          best = len;
          out[0] = x[0], out[1] = x[1], out[2] = x[2];
        }
      }
}

// _gdml/desc.py:42-76 (_pbc_diff): clamp a pair difference to the lattice a MODEL carries (sGDML periodic descriptors;
// lattice vectors as columns): d -= lat . around(lat_inv . d).  np.around rounds half to even, as rint does.
__device__ __forceinline__ void pbc_diff(const double* __restrict__ lat, const double* __restrict__ lat_inv, double d[3]) {
  double c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = rint(lat_inv[3 * i] * d[0] + lat_inv[3 * i + 1] * d[1] + lat_inv[3 * i + 2] * d[2]);
  const double s0 = lat[0] * c[0] + lat[1] * c[1] + lat[2] * c[2];
  const double s1 = lat[3] * c[0] + lat[4] * c[1] + lat[5] * c[2];
  const double s2 = lat[6] * c[0] + lat[7] * c[1] + lat[8] * c[2];
  d[0] -= s0, d[1] -= s1, d[2] -= s2;
}

// Decode candidate -> entity tuple.  Returns false when the tuple is not strictly ascending
// (utils.gen_combs drops it, utils.py:479-488).  Explicit tuples are taken as given.
__device__ __forceinline__ bool decode_tuple(const JobDev& j, long long cand, int ent[kMaxOrder]) {
  if (j.explicit_combs) {
    for (int s = 0; s < j.order; ++s) ent[s] = j.combs[cand * j.order + s];
    return true;
  }
  if (j.homogeneous) {
    // unrank combination number `cand` (lexicographic): slot s takes the smallest position p > prev with
    // #(combinations whose slot s is in (prev, p]) = C(n-prev-1, r+1) - C(n-p-1, r+1) > rem, r = slots after s
    const int n = j.n_set, w = j.order + 1;
    long long rem = cand;
    int prev = -1;
    for (int s = 0; s < j.order; ++s) {
      const int r = j.order - s - 1;
      int p;
      if (r == 0) {
        p = prev + 1 + (int)rem;
      } else {
        const long long top = j.binom[(long long)(n - prev - 1) * w + (r + 1)];
        const long long target = top - rem;  // largest m with C(m, r+1) < target
        int lo = r, hi = n - prev - 2;       // C(r, r+1) = 0 < target always
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (j.binom[(long long)mid * w + (r + 1)] < target) lo = mid; else hi = mid - 1;
        }
        p = n - 1 - lo;
        rem -= top - j.binom[(long long)(n - p) * w + (r + 1)];
      }
      ent[s] = j.set_ent[p];
      prev = p;
    }
    return true;
  }
  long long idx = cand;
  for (int s = j.order - 1; s >= 0; --s) {
    const int n_s = j.set_off[s + 1] - j.set_off[s];
    const int dig = (int)(idx % n_s);
    idx /= n_s;
    ent[s] = j.set_ent[j.set_off[s] + dig];
  }
  bool asc = true;
  for (int s = 0; s + 1 < j.order; ++s) asc = asc && (ent[s] < ent[s + 1]);
  return asc;
}

// gdml_predict.py:214-216: atoms of the fragment, entity by entity, ascending within an entity.
__device__ __forceinline__ int gather_atoms(const SysDev& s, const JobDev& j, const int ent[kMaxOrder],
                                            int atoms[kMaxFragAtoms]) {
  int a = 0;
  for (int k = 0; k < j.order; ++k) {
    const int lo = s.ent_off[ent[k]], hi = s.ent_off[ent[k] + 1];
    for (int i = lo; i < hi && a < kMaxFragAtoms; ++i) atoms[a++] = s.ent_atoms[i];
  }
  return a;
}

// mbeAlchemyScale.scale with linear_switching for every scaler whose entity is in the tuple
// (gdml_predict.py:256-260): the product of the factors.
__device__ __forceinline__ double alchemy_lambda(const JobDev& j, const int ent[kMaxOrder]) {
  double lam = 1.0;
  for (int a = 0; a < j.n_alch; ++a) {
    bool hit = false;
    for (int s = 0; s < j.order; ++s) hit = hit || (ent[s] == j.alch_ent[a]);
    if (hit) lam *= j.alch_fac[a];
  }
  return lam;
}

// Load the fragment's coordinates; under periodicity re-express them relative to the first atom
// in the minimum image and apply the pair cut-off (periodic.py:61-107).  Returns false if rejected.
template <int NA>
__device__ __forceinline__ bool fragment_coords(const SysDev& sys, long long frame, const double* __restrict__ Rf,
                                                const int* atoms, double (*r)[3]) {
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    const double* p = Rf + 3 * (long long)atoms[a];
    r[a][0] = p[0], r[a][1] = p[1], r[a][2] = p[2];
  }
  if (!sys.periodic) return true;
  const CellDev& s = cell_of(sys, frame);
  const double r0[3] = {r[0][0], r[0][1], r[0][2]};
  bool naive_ok = true, naive_far = false;
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    const double d[3] = {__dsub_rn(r[a][0], r0[0]), __dsub_rn(r[a][1], r0[1]), __dsub_rn(r[a][2], r0[2])};
    double v[3];
    const double len = mic_naive(s, d, v);
    naive_ok = naive_ok && (len < s.half_min_len);
    naive_far = naive_far || (len >= s.half_min_len * kGuard);
    r[a][0] = v[0], r[a][1] = v[1], r[a][2] = v[2];
  }
  if (!naive_ok) {
    // find_mic falls back to the general algorithm for the whole set of vectors.  Shortcut (orthorhombic cell,
    // cutoff <= L_min/2): some |d_0k| is beyond L_min/2 by more than the rounding differences between the naive and
    // the general image (a few ulp; the guard is 1e-9), so the general image of that atom is >= cutoff from atom 0
    // and the pair test below must fail.  Inside the guard band the general branch decides, as in find_mic.
    if (s.fast_reject && naive_far) return false;
    for (int a = 0; a < NA; ++a) {
      const double* p = Rf + 3 * (long long)atoms[a];
      const double d[3] = {__dsub_rn(p[0], r0[0]), __dsub_rn(p[1], r0[1]), __dsub_rn(p[2], r0[2])};
      double v[3];
      mic_general(s, d, v);
      r[a][0] = v[0], r[a][1] = v[1], r[a][2] = v[2];
    }
  }
  // periodic.py:80-83: reject if any pair distance >= cutoff
  bool ok = true;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int k = i + 1; k < NA; ++k) {
      const double dist = norm3_rn(__dsub_rn(r[i][0], r[k][0]), __dsub_rn(r[i][1], r[k][1]),
                                   __dsub_rn(r[i][2], r[k][2]));
      ok = ok && !(dist >= s.cutoff);
    }
  return ok;
}

// Criteria.accept (descriptors.py:65-103) with com_distance_sum / max_atom_pair_dist.
template <int NA>
__device__ __forceinline__ double criteria_value(const SysDev& s, const CritDev& c, const int* atoms,
                                                 const double (*r)[3]) {
  if (c.kind == MBG_CRIT_MAX_ATOM_PAIR_DIST) {
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
      for (int k = i + 1; k < NA; ++k) {
        const double dist = norm3_rn(__dsub_rn(r[i][0], r[k][0]), __dsub_rn(r[i][1], r[k][1]),
                                     __dsub_rn(r[i][2], r[k][2]));
        mx = dist > mx ? dist : mx;
      }
    return mx;
  }
  // com_distance_sum: sum over entity groups of |CM_total - CM_group| (descriptors.py:159-201),
  // centres of mass as np.average(R, weights=masses) = sum(R*m)/sum(m) in atom order (106-128).
  double m[NA];
  double num[3] = {0.0, 0.0, 0.0}, den = 0.0;
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    m[a] = s.masses[atoms[a]];
#pragma unroll
    for (int k = 0; k < 3; ++k) num[k] = __dadd_rn(num[k], __dmul_rn(r[a][k], m[a]));
    den = __dadd_rn(den, m[a]);
  }
  const double cm[3] = {__ddiv_rn(num[0], den), __ddiv_rn(num[1], den), __ddiv_rn(num[2], den)};
  double dsum = 0.0;
  for (int g = 0; g < c.n_groups; ++g) {
    double gn[3] = {0.0, 0.0, 0.0}, gd = 0.0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
      if (c.group[a] == g) {
#pragma unroll
        for (int k = 0; k < 3; ++k) gn[k] = __dadd_rn(gn[k], __dmul_rn(r[a][k], m[a]));
        gd = __dadd_rn(gd, m[a]);
      }
    const double dx = __dsub_rn(cm[0], __ddiv_rn(gn[0], gd));
    const double dy = __dsub_rn(cm[1], __ddiv_rn(gn[1], gd));
    const double dz = __dsub_rn(cm[2], __ddiv_rn(gn[2], gd));
    dsum = __dadd_rn(dsum, norm3_rn(dx, dy, dz));
  }
  return dsum;
}

__device__ __forceinline__ bool criteria_accept(const CritDev& c, double v) {
  if (c.bound == MBG_BOUND_RANGE) return (c.lo < v) && (v < c.hi);
  if (c.bound == MBG_BOUND_UPPER) return v < c.hi;
  return v > c.lo;
}

// Cheap exact pre-filter for the culling kernel: the subset of fragment_coords' rejection tests that only
// involves the FIRST atom of every entity (under an orthorhombic cell with cutoff <= L_min/2, where a naive
// minimum image beyond L_min/2 already implies rejection).  It rejects only when a distance is beyond its
// threshold by the guard band (the fragment's final coordinates may come from the general branch, whose
// rounding differs from the naive image by a few ulp); everything closer goes to the full evaluation, so
// returning false early never changes the accept mask.  Consecutive candidates share their leading entities,
// so warps exit together.
__device__ __forceinline__ bool first_atoms_may_pass(const SysDev& sys, long long frame, const JobDev& j,
                                                     const double* __restrict__ Rf, const int ent[kMaxOrder]) {
  if (!sys.periodic) return true;
  const CellDev& s = cell_of(sys, frame);
  if (!s.fast_reject) return true;
  double v[kMaxOrder][3];
  bool safe[kMaxOrder];  // the naive image is the unique minimum image, whatever branch find_mic ends up taking
  const double* p0 = Rf + 3ll * sys.ent_atoms[sys.ent_off[ent[0]]];
#pragma unroll
  for (int k = 0; k < kMaxOrder; ++k) {
    if (k >= j.order) break;
    const double* p = Rf + 3ll * sys.ent_atoms[sys.ent_off[ent[k]]];
    const double d[3] = {__dsub_rn(p[0], p0[0]), __dsub_rn(p[1], p0[1]), __dsub_rn(p[2], p0[2])};
    const double len = mic_naive(s, d, v[k]);
    if (len >= s.half_min_len * kGuard) return false;
    safe[k] = len * kGuard < s.half_min_len;
#pragma unroll
    for (int i = 0; i < kMaxOrder; ++i) {
      if (i >= k) break;
      const double dist = norm3_rn(__dsub_rn(v[i][0], v[k][0]), __dsub_rn(v[i][1], v[k][1]), __dsub_rn(v[i][2], v[k][2]));
      if (safe[i] && safe[k] && dist >= s.cutoff * kGuard) return false;
    }
  }
  return true;
}

// Full accept decision for one fragment; r receives the coordinates the descriptor is built from.
template <int NA>
__device__ __forceinline__ bool fragment_accept(const SysDev& s, long long frame, const CritDev& c,
                                                const double* __restrict__ Rf, const int* atoms, double (*r)[3]) {
  if (!fragment_coords<NA>(s, frame, Rf, atoms, r)) return false;
  if (c.kind == MBG_CRIT_NONE) return true;
  return criteria_accept(c, criteria_value<NA>(s, c, atoms, r));
}

}  // namespace mbg
