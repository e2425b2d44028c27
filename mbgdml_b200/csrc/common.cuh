// Shared definitions for the mbgdml_b200 CUDA translation unit (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/mbgdml_b200.h"

namespace mbg {

constexpr int kMaxOrder = MBG_MAX_ORDER;
constexpr int kMaxFragAtoms = MBG_MAX_FRAG_ATOMS;
constexpr int kMaxAlchemy = MBG_MAX_ALCHEMY;
constexpr int kCullThreads = 256;  // candidates per cull/compact block == sharding granule

// The supersystem on device.
struct SysDev {
  int n_atoms;
  int n_entities;
  const int* ent_off;    // [n_entities+1]
  const int* ent_atoms;  // CSR values
  const double* masses;  // [n_atoms]
  int periodic;          // 0: cluster, 1: minimum-image under `cell`
  int fast_reject;       // orthorhombic cell and cutoff <= half_min_len: a failed naive MIC implies rejection
  double cell[9], cell_inv[9];    // row vectors; f = v @ cell_inv
  double rcell[9], rcell_inv[9];  // Minkowski-reduced cell for the general branch
  double half_min_len;            // 0.5 * min |cell vector|
  double cutoff;                  // Cell.cutoff
};

// Acceptance criteria of one model.
struct CritDev {
  int kind, bound;
  double lo, hi;
  int n_groups;
  int group[kMaxFragAtoms];  // entity label (0..n_groups-1) of each fragment atom, for com_distance_sum
};

// Fragment source + alchemy of one job.
struct JobDev {
  int order;
  int n_frag_atoms;
  int explicit_combs;       // 0: product of sets (ascending filter), 1: explicit tuples
  int homogeneous;          // 1: every slot draws from the same ascending set -> candidates are the C(n, order)
                            //    combinations, unranked in lexicographic order (== product order of the ascending tuples)
  int n_set;                // size of that set
  const long long* binom;   // [(n_set + 1) x (order + 1)] binomial table, binom[m * (order + 1) + r] = C(m, r)
  int set_off[kMaxOrder + 1];
  const int* set_ent;       // concatenated per-slot entity ids (device)
  const int* combs;         // [n][order] (device) when explicit
  long long cand_per_frame; // product size, C(n_set, order), or n_combs
  int n_alch;
  int alch_ent[kMaxAlchemy];
  double alch_fac[kMaxAlchemy];
};

}  // namespace mbg
