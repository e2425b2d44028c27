// MBE recombination: add / remove lower-order contributions to reference structures.
//
// Reference: mbgdml/mbe.py:73-197 (mbe_worker: for every reference structure, for every combination of its
// entities of the lower order, find the lower-order structure and add its energy and its derivatives onto the
// matching atoms) and mbe.py:381-427 (mbe_contrib: E[r_idxs] +-= E_worker).  The reference finds each
// lower-order row by a linear np.all(...) scan (mbe.py:154-156); here the host does a sort join
// (mbgdml_b200/mbe.py:_mbe_pairs) and the device does the segmented scatter.
//
// One thread per output element (structure i, atom a, component) walks the pairs of structure i IN THE
// REFERENCE'S COMBINATION ORDER, so every sum is accumulated in the same order as the reference's loop
// and the result is bit-identical (no atomics).
#pragma once
#include "common.cuh"

namespace mbg {

struct MbeAccumArgs {
  long long n_ref;          // reference structures selected (rows of ref_rows)
  int n_atoms_ref, n_atoms_lower, n_slots;  // n_slots = entities of a lower-order structure
  const long long* ref_rows;      // [n_ref] row of E / Deriv each selected structure updates
  const long long* pair_offsets;  // [n_ref + 1]
  const long long* lower_idx;     // [n_pairs] row of E_lower / Deriv_lower
  const int* slot_entity;         // [n_pairs][n_slots] entity id (in the reference structure) slot k is added onto
  const int* atom_entity;         // [n_atoms_ref] entity id of each reference atom
  const int* atom_rank;           // [n_atoms_ref] index of the atom within its entity (ascending atom order)
  const int* lower_off;           // [n_slots + 1] CSR over lower_atoms
  const int* lower_atoms;         // atoms of the lower structure with entity_ids_lower == k, ascending
  const double* E_lower;
  const double* D_lower;          // [n_lower][n_atoms_lower][3]
  double sign;                    // +1 add, -1 remove
  double* E;                      // [n_total]
  double* D;                      // [n_total][n_atoms_ref][3]
};

__global__ void __launch_bounds__(256) k_mbe_accumulate(const MbeAccumArgs a) {
  const long long per = (long long)a.n_atoms_ref * 3 + 1;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.n_ref * per) return;
  const long long i = gid / per;
  const int rem = (int)(gid - i * per);
  const long long row = a.ref_rows[i];
  const long long p0 = a.pair_offsets[i], p1 = a.pair_offsets[i + 1];
  if (rem == a.n_atoms_ref * 3) {  // energy: E[i] += E_lower[...] in pair order (mbe.py:165)
    double acc = 0.0;
    for (long long p = p0; p < p1; ++p) acc = __dadd_rn(acc, a.E_lower[a.lower_idx[p]]);
    a.E[row] = __dadd_rn(a.E[row], a.sign * acc);
    return;
  }
  const int atom = rem / 3, c = rem - atom * 3;
  const int ent = a.atom_entity[atom], rank = a.atom_rank[atom];
  double acc = 0.0;
  for (long long p = p0; p < p1; ++p) {
    const int* se = a.slot_entity + p * a.n_slots;
    for (int k = 0; k < a.n_slots; ++k)  // Deriv[i][i_z] += deriv_lower[i_z_lower] per slot (mbe.py:183-193)
      if (se[k] == ent) {
        const int la = a.lower_atoms[a.lower_off[k] + rank];
        acc = __dadd_rn(acc, a.D_lower[(a.lower_idx[p] * a.n_atoms_lower + la) * 3 + c]);
      }
  }
  double* dst = a.D + (row * a.n_atoms_ref + atom) * 3 + c;
  *dst = __dadd_rn(*dst, a.sign * acc);
}

}  // namespace mbg
